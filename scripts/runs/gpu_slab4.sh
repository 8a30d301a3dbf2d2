mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_slab.py tests/test_gpu_slab_peer.py -q > gpurun_out/pytest_slab_4gpu.log 2>&1; tail -5 gpurun_out/pytest_slab_4gpu.log | cut -c1-300
