mkdir -p gpurun_out; rm -f gpurun_out/tune_v3d.log
for np in 0 20; do for d in f64; do for n in 256; do echo "npml $np" >> gpurun_out/tune_v3d.log; TUNE_NPML=$np timeout 120 python scripts/tune.py $n $d "kernel_variant=3,xchunk=8" "kernel_variant=2,xchunk=0" >> gpurun_out/tune_v3d.log 2>&1; done; done; done
cat gpurun_out/tune_v3d.log
