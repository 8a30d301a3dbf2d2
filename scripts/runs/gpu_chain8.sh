mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -k "config4" > gpurun_out/chain8_pytest.log 2>&1; tail -12 gpurun_out/chain8_pytest.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_fields.py -q -k "quirks" 2>&1 | tail -2
