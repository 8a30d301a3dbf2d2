mkdir -p gpurun_out; rm -f gpurun_out/tune_fused2.log
for L in "$@"; do echo "== lib $L" >> gpurun_out/tune_fused2.log
 CEV_LIB_PATH=$PWD/tuning_libs/lib_$L.so timeout 300 python -m pytest tests/test_gpu_variants.py -x -q -k "fused" 2>&1 | tail -1 >> gpurun_out/tune_fused2.log
 for d in f64 f32; do for n in 256 512; do
  CEV_LIB_PATH=$PWD/tuning_libs/lib_$L.so TUNE_RUN=20 timeout 300 python scripts/tune.py $n $d "kernel_variant=4,fused_shape=1604" "fused_shape=804" "fused_shape=1604,prefetch_planes=2" >> gpurun_out/tune_fused2.log 2>&1
 done; done
done
cat gpurun_out/tune_fused2.log
