mkdir -p gpurun_out; rm -f gpurun_out/tune6.log
for np in 0,0,0 20,0,0 0,20,0 0,0,20 20,20,0 20,20,20; do for d in f64 f32; do echo "npml $np" >> gpurun_out/tune6.log; TUNE_NPML=$np timeout 300 python scripts/tune.py 256 $d "xchunk=0" >> gpurun_out/tune6.log 2>&1; done; done
cat gpurun_out/tune6.log
