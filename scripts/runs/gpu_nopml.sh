mkdir -p gpurun_out; rm -f gpurun_out/tune_nopml2.log
timeout 600 python -m pytest tests/test_gpu_variants.py -x -q -k "fused" 2>&1 | tail -1 >> gpurun_out/tune_nopml2.log
for d in f64 f32; do for n in 256 512; do
  TUNE_NPML=0 TUNE_RUN=20 timeout 300 python scripts/tune.py $n $d "kernel_variant=0,fused_step=0" "kernel_variant=4,fused_step=1" "kernel_variant=4,xchunk=32" "kernel_variant=4,xchunk=8" >> gpurun_out/tune_nopml2.log 2>&1
done; done
cat gpurun_out/tune_nopml2.log
