mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -k "config5" > gpurun_out/chain9_pytest.log 2>&1; tail -12 gpurun_out/chain9_pytest.log | cut -c1-300
