mkdir -p gpurun_out; rm -f gpurun_out/tune12.log
for L in 3_5 4_5; do for d in f64 f32; do for n in 256 512; do echo "lib $L" >> gpurun_out/tune12.log; CEV_LIB_PATH=$PWD/tuning_libs/lib_$L.so timeout 300 python scripts/tune.py $n $d "xchunk=0" >> gpurun_out/tune12.log 2>&1; done; done; done
cat gpurun_out/tune12.log
