mkdir -p gpurun_out
echo "== tests: scaled configs + gradients + variants"
timeout 1200 python -m pytest tests/test_gpu_scaled_configs.py tests/test_gpu_gradients.py tests/test_gpu_variants.py -q > gpurun_out/chain3_pytest.log 2>&1; tail -4 gpurun_out/chain3_pytest.log
echo "== 2-D tangent sweep: batched vs streams"
for d in f32 f64; do TUNE_B=16 timeout 600 python scripts/tune2d.py 2048 $d jvp_batch=0 "" 2>&1 | grep -v Warn | tail -2; done | tee gpurun_out/chain3_tune2d.log
C5_STEPS=300 timeout 600 python scripts/bench_configs.py c5 > gpurun_out/c5_batched.log 2>&1; tail -2 gpurun_out/c5_batched.log
echo "== fp32 drift"
timeout 900 python scripts/fp32_drift.py > gpurun_out/fp32_drift.json 2> gpurun_out/fp32_drift.err; tail -3 gpurun_out/fp32_drift.err
python - <<PY
import json
for t in json.load(open("gpurun_out/fp32_drift.json")):
    print(t["case"])
    for r in t["rows"]:
        print("  arith %s steps %5d fields %.2e series %.2e per-probe %s  energy/peak %.1e" % (r["arith"], r["time_steps"], r["fields_rel_l2"], r["series_rel_l2_up_to_here"], ["%.1e" % v for v in r["series_rel_l2_per_probe"]], r["E_energy_vs_peak"]))
PY
