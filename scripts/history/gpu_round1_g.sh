set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
for L in 4_8 4_6 3_6 3_5; do echo "== lib $L" >> gpurun_out/tune4.log; for d in f64 f32; do for n in 256 512; do CEV_LIB_PATH=$PWD/tuning_libs/lib_$L.so timeout 300 python scripts/tune.py $n $d "xchunk=0" "xchunk=8" >> gpurun_out/tune4.log 2>&1; done; done; done
