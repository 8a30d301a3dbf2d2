mkdir -p gpurun_out; rm -f gpurun_out/tune9.log
for L in base nopml normw; do for d in f64 f32; do echo "lib $L" >> gpurun_out/tune9.log; CEV_LIB_PATH=$PWD/tuning_libs/lib_$L.so timeout 300 python scripts/tune.py 256 $d "split_launch=0" >> gpurun_out/tune9.log 2>&1; done; done
cat gpurun_out/tune9.log
