# whole GPU test suite (gpurun -- 'bash scripts/gpu_suite.sh')
mkdir -p gpurun_out; rm -f gpurun_out/suite_*.log
timeout 2000 python -m pytest tests -q -m gpu > gpurun_out/suite_pytest.log 2>&1
tail -15 gpurun_out/suite_pytest.log
