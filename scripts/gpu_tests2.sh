# whole GPU suite on a 2-GPU box (the slab tests need >= 2 GPUs)
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/pytest_2gpu.log 2>&1
tail -8 gpurun_out/pytest_2gpu.log
