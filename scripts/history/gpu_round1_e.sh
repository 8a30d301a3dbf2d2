set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
C="lanes_z=8,prefetch_planes=1,xchunk=8 xchunk=4 xchunk=16 prefetch_planes=2,xchunk=8 prefetch_planes=2,xchunk=16 lanes_z=16,prefetch_planes=1,xchunk=8 prefetch_planes=2,xchunk=8 xchunk=16 lanes_z=32,prefetch_planes=1,xchunk=8 prefetch_planes=2,xchunk=16"
for d in f64 f32; do for n in 256 512; do timeout 300 python scripts/tune.py $n $d $C >> gpurun_out/tune3.log 2>&1; done; done
