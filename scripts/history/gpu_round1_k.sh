mkdir -p gpurun_out; rm -f gpurun_out/tune8.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
for L in i5 i6 i8; do for n in 256 512; do for d in f64 f32; do echo "lib $L" >> gpurun_out/tune8.log; CEV_LIB_PATH=$PWD/tuning_libs/lib_$L.so timeout 300 python scripts/tune.py $n $d "split_launch=1" "split_launch=0" >> gpurun_out/tune8.log 2>&1; done; done; done
cat gpurun_out/tune8.log
