mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; cat gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_f64.json 2> gpurun_out/bench_f64.err
timeout 900 python scripts/bench_configs.py c1 c4 c5 > gpurun_out/configs.log 2>&1
python - <<PY
import json
j = json.load(open("gpurun_out/bench_f64.json"))
print("bench_f64", round(j["value"], 3), j["e2e"]["value"], j["roofline"]["frac"], j["cpu_baseline"], j["clocks"])
PY
cat gpurun_out/configs.log
