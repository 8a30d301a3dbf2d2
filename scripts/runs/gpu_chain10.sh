mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gradients.py -q > gpurun_out/chain10_pytest.log 2>&1; tail -6 gpurun_out/chain10_pytest.log | cut -c1-300
