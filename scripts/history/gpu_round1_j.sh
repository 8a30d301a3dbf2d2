mkdir -p gpurun_out; rm -f gpurun_out/tune7.log
for L in 3_5 4_5 4_6 3_6 4_8; do for np in 0,0,0 20,20,20; do for d in f64 f32; do echo "lib $L npml $np" >> gpurun_out/tune7.log; CEV_LIB_PATH=$PWD/tuning_libs/lib_$L.so TUNE_NPML=$np timeout 300 python scripts/tune.py 256 $d "xchunk=0" >> gpurun_out/tune7.log 2>&1; done; done; done
cat gpurun_out/tune7.log
