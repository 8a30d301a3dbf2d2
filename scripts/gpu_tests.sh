mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q "$@" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -40 gpurun_out/pytest.log
