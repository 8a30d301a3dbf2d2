set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
C="kernel_variant=2,prefetch_planes=0,xchunk=0 prefetch_planes=1 prefetch_planes=2 prefetch_planes=4 prefetch_planes=8 prefetch_planes=0,xchunk=8 xchunk=16 xchunk=32 xchunk=64 xchunk=16,prefetch_planes=4 xchunk=32,prefetch_planes=4 kernel_variant=1"
for d in f64 f32; do for n in 256 512; do timeout 300 python scripts/tune.py $n $d $C >> gpurun_out/tune.log 2>&1; done; done
