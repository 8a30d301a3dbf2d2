mkdir -p gpurun_out; rm -f gpurun_out/tune_v3b.log
for L in b4_s3 b4_s4 b8_s3 b8_s4; do
  echo "== lib $L" >> gpurun_out/tune_v3b.log
  CEV_LIB_PATH=$PWD/tuning_libs/lib_$L.so timeout 200 python -m pytest tests/test_gpu_variants.py -x -q -k "bitwise and f64" 2>&1 | tail -1 >> gpurun_out/tune_v3b.log
  for d in f64 f32; do for n in 256 512; do CEV_LIB_PATH=$PWD/tuning_libs/lib_$L.so timeout 120 python scripts/tune.py $n $d "kernel_variant=3,xchunk=8" "xchunk=16" >> gpurun_out/tune_v3b.log 2>&1; done; done; done
cat gpurun_out/tune_v3b.log
