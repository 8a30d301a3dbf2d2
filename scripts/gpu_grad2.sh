mkdir -p gpurun_out
C4_STEPS=60 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,smsp__inst_executed.sum --clock-control none -k regex:k_adj -s 40 -c 8 --csv --log-file gpurun_out/grad_launches.csv python scripts/bench_configs.py c4 > gpurun_out/grad_c4_ncu.log 2>&1
tail -30 gpurun_out/grad_launches.csv | cut -c1-300
