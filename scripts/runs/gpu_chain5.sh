mkdir -p gpurun_out
C5_STEPS=20000 timeout 900 python scripts/bench_configs.py c5 > gpurun_out/c5_20000.log 2>&1; tail -2 gpurun_out/c5_20000.log | cut -c1-900
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/suite_pytest.log 2>&1; tail -4 gpurun_out/suite_pytest.log
timeout 900 python bench.py > gpurun_out/suite_bench.json 2> gpurun_out/suite_bench.err; tail -c 1500 gpurun_out/suite_bench.json; tail -3 gpurun_out/suite_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/suite_bench_ref.json 2>&1; tail -c 1200 gpurun_out/suite_bench_ref.json
