mkdir -p gpurun_out
TUNE_NPML=0 TUNE_RUN=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step_fused -s 6 -c 1 -f -o gpurun_out/prof_fused_nopml_f64_256 python scripts/tune.py 256 f64 "kernel_variant=4" > gpurun_out/ncu_fused_nopml.log 2>&1
TUNE_NPML=0 TUNE_RUN=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step_fused -s 6 -c 1 -f -o gpurun_out/prof_fused_nopml_f32_256 python scripts/tune.py 256 f32 "kernel_variant=4" >> gpurun_out/ncu_fused_nopml.log 2>&1
tail -4 gpurun_out/ncu_fused_nopml.log
