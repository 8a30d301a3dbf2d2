mkdir -p gpurun_out
C5_STEPS=300 timeout 600 python scripts/bench_configs.py c5 > gpurun_out/c5_base.log 2>&1
tail -2 gpurun_out/c5_base.log
C5_STEPS=12 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_step -s 300 -c 8 -o gpurun_out/prof_c5_f64 -f python scripts/bench_configs.py c5 > gpurun_out/c5_ncu.log 2>&1
tail -3 gpurun_out/c5_ncu.log
