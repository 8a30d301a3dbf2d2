mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -8 gpurun_out/pytest.log
timeout 600 python scripts/bench_configs.py c1 c5 2>&1 | tee gpurun_out/configs_2d.log
