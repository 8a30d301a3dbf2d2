mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slab.py -q -k "gradient" > gpurun_out/slabgrad_pytest.log 2>&1
tail -8 gpurun_out/slabgrad_pytest.log | cut -c1-400
