set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
C="prefetch_planes=0,xchunk=0 xchunk=16 xchunk=8,prefetch_planes=1"
for mb in 4 5 6 8; do echo "== V2_MIN_CTAS=$mb" >> gpurun_out/tune2.log; for d in f64 f32; do for n in 256 512; do CEV_LIB_PATH=$PWD/tuning_libs/gpurun_in_lib_mb$mb.so timeout 300 python scripts/tune.py $n $d $C >> gpurun_out/tune2.log 2>&1; done; done; done
