mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_variants.py -x -q -k "fused" > gpurun_out/pytest_fused.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fused.log
tail -25 gpurun_out/pytest_fused.log
rm -f gpurun_out/tune_fused.log
for d in f64 f32; do for n in 256 512; do
  TUNE_RUN=20 timeout 300 python scripts/tune.py $n $d "kernel_variant=2" "kernel_variant=0" "kernel_variant=4,fused_shape=1604" "fused_shape=804" "fused_shape=3204" "fused_shape=1608" "fused_shape=3208" "fused_shape=1604,xchunk=32" "xchunk=8" >> gpurun_out/tune_fused.log 2>&1
done; done
cat gpurun_out/tune_fused.log
