mkdir -p gpurun_out; rm -f gpurun_out/tune_jvp.log
timeout 900 python -m pytest tests/test_gpu_gradients.py tests/test_gpu_variants.py -x -q -k "jvp or tangent or forward_mode" 2>&1 | tail -3 | tee -a gpurun_out/tune_jvp.log
for d in f32 f64; do
 TUNE_B=16 timeout 600 python scripts/tune2d.py 2048 $d "jvp_streams=0" "jvp_streams=1" "jvp_streams=0" "jvp_streams=1" >> gpurun_out/tune_jvp.log 2>&1
 TUNE_B=16 timeout 600 python scripts/tune2d.py 640 $d "jvp_streams=0" "jvp_streams=1" >> gpurun_out/tune_jvp.log 2>&1
done
cat gpurun_out/tune_jvp.log
