mkdir -p gpurun_out; rm -f gpurun_out/tune_pml.log
for L in "$@"; do echo "== lib $L" >> gpurun_out/tune_pml.log
 CEV_LIB_PATH=$PWD/tuning_libs/lib_$L.so timeout 600 python -m pytest tests/test_gpu_variants.py tests/test_gpu_fields.py -x -q 2>&1 | tail -1 >> gpurun_out/tune_pml.log
 for d in f64 f32; do for n in 256 512; do
  CEV_LIB_PATH=$PWD/tuning_libs/lib_$L.so timeout 300 python scripts/tune.py $n $d "kernel_variant=0" "kernel_variant=2" >> gpurun_out/tune_pml.log 2>&1
  CEV_LIB_PATH=$PWD/tuning_libs/lib_$L.so TUNE_RUN=20 timeout 300 python scripts/tune.py $n $d "kernel_variant=0,fused_step=0" "kernel_variant=4" >> gpurun_out/tune_pml.log 2>&1
 done; done
done
cat gpurun_out/tune_pml.log
