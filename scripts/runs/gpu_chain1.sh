mkdir -p gpurun_out
echo "== adjoint min-blocks A/B (c4, 400 steps)"
for v in default mb3 mb4; do
  if [ $v = default ]; then unset CEV_LIB_PATH; else export CEV_LIB_PATH=$PWD/build_variants/libcev_$v.so; fi
  C4_STEPS=400 timeout 600 python scripts/bench_configs.py c4 > gpurun_out/adj_ab_$v.log 2>&1
  echo "-- $v"; python - <<PY
import json
for l in open("gpurun_out/adj_ab_$v.log"):
    if l.startswith("{"):
        r = json.loads(l); print(r["dtype"], r["config"][60:110], "fwd %.1f bwd %.1f" % (r["forward_gcell_per_s"], r["backward_gcell_per_s"]))
PY
done
unset CEV_LIB_PATH
echo "== c4 at 2000 steps + ncu"
bash scripts/gpu_grad2.sh
echo "== fp32 drift"
timeout 900 python scripts/fp32_drift.py > gpurun_out/fp32_drift.json 2> gpurun_out/fp32_drift.err; tail -3 gpurun_out/fp32_drift.err
python - <<PY
import json
for t in json.load(open("gpurun_out/fp32_drift.json")):
    print(t["case"])
    for r in t["rows"]:
        print("  arith %s steps %5d fields %.2e E %.2e worst %.2e series %.2e  energy/peak %.1e" % (r["arith"], r["time_steps"], r["fields_rel_l2"], r["E_rel_l2"], r["worst_field_rel_l2"], r["series_rel_l2_up_to_here"], r["E_energy_vs_peak"]))
PY
echo "== c5 profile"
bash scripts/runs/gpu_c5prof.sh
