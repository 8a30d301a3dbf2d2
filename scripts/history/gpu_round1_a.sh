set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --chunk 200 > gpurun_out/bench_f64.json 2> gpurun_out/bench_f64.err
timeout 300 python bench.py --steps 3 --warmup 3 --chunk 200 --dtype f32 --arith f32 --no-cpu > gpurun_out/bench_f32.json 2> gpurun_out/bench_f32.err
timeout 300 python bench.py --steps 3 --warmup 3 --chunk 200 --dtype f32 --arith f64 --no-cpu > gpurun_out/bench_f32a64.json 2> gpurun_out/bench_f32a64.err
timeout 300 python bench.py --grid 512 512 512 --steps 2 --warmup 3 --chunk 50 --no-cpu > gpurun_out/bench_f64_512.json 2> gpurun_out/bench_f64_512.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/launches_v1.csv python bench.py --steps 1 --warmup 3 --chunk 10 --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 20 -c 4 -o gpurun_out/prof_v1 python bench.py --steps 1 --warmup 3 --chunk 10 --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
