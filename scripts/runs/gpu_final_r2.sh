# round-2 final evidence run (1 GPU): suite, smoke, bench (both arms)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/suite_pytest.log 2>&1; tail -3 gpurun_out/suite_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warn | tail -1
timeout 900 python bench.py > gpurun_out/suite_bench.json 2> gpurun_out/suite_bench.err; tail -c 300 gpurun_out/suite_bench.json; tail -3 gpurun_out/suite_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/suite_bench_ref.json 2>&1; tail -c 200 gpurun_out/suite_bench_ref.json
