mkdir -p gpurun_out
echo "== tests: scaled configs + gradients"
timeout 900 python -m pytest tests/test_gpu_scaled_configs.py tests/test_gpu_gradients.py -q > gpurun_out/chain2_pytest.log 2>&1; tail -3 gpurun_out/chain2_pytest.log
echo "== 2-D tangent sweep: streams on / off"
for d in f32 f64; do TUNE_B=16 timeout 600 python scripts/tune2d.py 2048 $d jvp_streams=0 jvp_streams=1 "" 2>&1 | grep -v Warn | tail -3; done | tee gpurun_out/chain2_tune2d.log
