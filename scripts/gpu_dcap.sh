mkdir -p gpurun_out; rm -f gpurun_out/tune_dcap.log
for L in base d4 d6; do echo "== $L" >> gpurun_out/tune_dcap.log
 for n in 256 512; do
  if [ $L = base ]; then timeout 300 python scripts/tune.py $n f32 "kernel_variant=0" >> gpurun_out/tune_dcap.log 2>&1
  else CEV_LIB_PATH=$PWD/tuning_libs/lib_$L.so timeout 300 python scripts/tune.py $n f32 "kernel_variant=0" >> gpurun_out/tune_dcap.log 2>&1; fi
 done
done
cat gpurun_out/tune_dcap.log
