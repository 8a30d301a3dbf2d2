mkdir -p gpurun_out
C5_STEPS=12 timeout 900 ncu --set full --import-source on --clock-control none -k regex:_batch -s 8 -c 4 -o gpurun_out/prof_c5_batch -f python scripts/bench_configs.py c5 > gpurun_out/c5_batch_ncu.log 2>&1; tail -2 gpurun_out/c5_batch_ncu.log | cut -c1-200
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/suite_pytest.log 2>&1; tail -3 gpurun_out/suite_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warn | tail -1
timeout 900 python bench.py > gpurun_out/suite_bench.json 2> gpurun_out/suite_bench.err; tail -c 400 gpurun_out/suite_bench.json; tail -3 gpurun_out/suite_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/suite_bench_ref.json 2>&1; tail -c 300 gpurun_out/suite_bench_ref.json
