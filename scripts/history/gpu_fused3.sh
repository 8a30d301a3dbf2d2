mkdir -p gpurun_out; rm -f gpurun_out/tune_fused3.log
for d in f64 f32; do for n in 256 512; do
  TUNE_RUN=20 timeout 300 python scripts/tune.py $n $d "kernel_variant=0,fused_step=0" "kernel_variant=2" "kernel_variant=0,fused_step=1" "xchunk=24" "xchunk=12" >> gpurun_out/tune_fused3.log 2>&1
done; done
cat gpurun_out/tune_fused3.log
TUNE_RUN=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step_fused -s 6 -c 1 -o gpurun_out/prof_fused_f64_256 -f python scripts/tune.py 256 f64 "kernel_variant=4" > gpurun_out/ncu_fused.log 2>&1
TUNE_RUN=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step_fused -s 6 -c 1 -o gpurun_out/prof_fused_f32_256 -f python scripts/tune.py 256 f32 "kernel_variant=4" >> gpurun_out/ncu_fused.log 2>&1
tail -5 gpurun_out/ncu_fused.log
