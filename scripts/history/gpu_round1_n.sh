mkdir -p gpurun_out; rm -f gpurun_out/tune11.log
for L in d4 d5 d6; do for d in f64 f32; do for n in 256 512; do echo "lib $L" >> gpurun_out/tune11.log; CEV_LIB_PATH=$PWD/tuning_libs/lib_$L.so timeout 300 python scripts/tune.py $n $d "xchunk=0" >> gpurun_out/tune11.log 2>&1; done; done; done
cat gpurun_out/tune11.log
C4_STEPS=200 timeout 600 python scripts/bench_configs.py c4 > gpurun_out/configs_c4.log 2>&1; cat gpurun_out/configs_c4.log
