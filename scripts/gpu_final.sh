mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; cat gpurun_out/smoke.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 900 python bench.py > gpurun_out/bench_f64.json 2> gpurun_out/bench_f64.err
timeout 600 python bench.py --dtype f32 --no-cpu > gpurun_out/bench_f32.json 2> gpurun_out/bench_f32.err
timeout 600 python bench.py --grid 512 512 512 --steps 4 --warmup 3 --chunk 100 --no-cpu > gpurun_out/bench_f64_512.json 2> gpurun_out/bench_f64_512.err
timeout 600 python bench.py --grid 512 512 512 --steps 4 --warmup 3 --chunk 100 --no-cpu --dtype f32 > gpurun_out/bench_f32_512.json 2> gpurun_out/bench_f32_512.err
timeout 900 python scripts/bench_configs.py c1 c4 c5 > gpurun_out/configs.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --chunk 10 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 20 -c 2 -f -o gpurun_out/prof_r1d_f64_256 python bench.py --steps 1 --warmup 3 --chunk 10 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 20 -c 2 -f -o gpurun_out/prof_r1d_f32_256 python bench.py --dtype f32 --steps 1 --warmup 3 --chunk 10 --no-cpu --no-e2e > gpurun_out/ncu_full32.log 2>&1
for f in bench_reference bench_f64 bench_f32 bench_f64_512 bench_f32_512; do python - <<PY
import json
try:
    j = json.load(open("gpurun_out/$f.json"))
    print("$f", round(j["value"], 4), (j.get("e2e") or {}).get("value"), (j.get("roofline") or {}).get("frac"), j.get("clocks"))
except Exception as e:
    print("$f", "FAILED", e)
PY
done
cat gpurun_out/configs.log
