mkdir -p gpurun_out; rm -f gpurun_out/tune2d.log
for d in f32 f64; do for n in 2048 640; do
 timeout 600 python scripts/tune2d.py $n $d "xchunk=0" "xchunk=1" "xchunk=2" "xchunk=4" "xchunk=8" "xchunk=16" "xchunk=32" "xchunk=64" "xchunk=8,prefetch_planes=2" "xchunk=8,prefetch_planes=4" "xchunk=16,prefetch_planes=2" "xchunk=16,prefetch_planes=0" >> gpurun_out/tune2d.log 2>&1
done; done
cat gpurun_out/tune2d.log
