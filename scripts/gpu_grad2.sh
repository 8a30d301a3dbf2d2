mkdir -p gpurun_out
C4_STEPS=2000 timeout 900 python scripts/bench_configs.py c4 > gpurun_out/grad_c4_2000.log 2>&1
tail -4 gpurun_out/grad_c4_2000.log
C4_STEPS=60 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_adj_.*_v5 -s 40 -c 4 -o gpurun_out/prof_adj_v5_c4 -f python scripts/bench_configs.py c4 > gpurun_out/grad_c4_ncu.log 2>&1
tail -3 gpurun_out/grad_c4_ncu.log
