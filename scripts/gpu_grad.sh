mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_gradients.py tests/test_gpu_scaled_configs.py -q -x > gpurun_out/grad_pytest.log 2>&1
tail -25 gpurun_out/grad_pytest.log
C4_STEPS=${C4_STEPS:-400} timeout 900 python scripts/bench_configs.py c4 > gpurun_out/grad_c4.log 2>&1
tail -4 gpurun_out/grad_c4.log
