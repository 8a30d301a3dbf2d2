# full 1-GPU suite + default bench + reference arm   (gpurun -- 'bash scripts/gpu_suite.sh')
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/suite_pytest.log 2>&1
tail -6 gpurun_out/suite_pytest.log
timeout 900 python bench.py > gpurun_out/suite_bench.json 2> gpurun_out/suite_bench.err
tail -c 6000 gpurun_out/suite_bench.json; tail -5 gpurun_out/suite_bench.err
