mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_gradients.py -q > gpurun_out/chain4_pytest.log 2>&1; tail -12 gpurun_out/chain4_pytest.log | cut -c1-300
timeout 600 python scripts/bench_small_adjoint.py 2>&1 | grep -v Warn | tail -6 | tee gpurun_out/chain4_small_adjoint.log
