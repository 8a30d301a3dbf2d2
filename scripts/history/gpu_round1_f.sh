set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
for d in f64 f32; do for n in 256 512; do TUNE_NPML=0 timeout 300 python scripts/tune.py $n $d "xchunk=0" >> gpurun_out/tune_nopml.log 2>&1; done; done
timeout 900 python bench.py > gpurun_out/bench_f64.json 2> gpurun_out/bench_f64.err
timeout 600 python bench.py --dtype f32 --no-cpu > gpurun_out/bench_f32.json 2> gpurun_out/bench_f32.err
timeout 600 python bench.py --grid 512 512 512 --steps 4 --warmup 3 --chunk 100 --no-cpu > gpurun_out/bench_f64_512.json 2> gpurun_out/bench_f64_512.err
timeout 600 python bench.py --grid 512 512 512 --steps 4 --warmup 3 --chunk 100 --no-cpu --dtype f32 > gpurun_out/bench_f32_512.json 2> gpurun_out/bench_f32_512.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --chunk 10 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 20 -c 2 -o gpurun_out/prof_r1_f64_256 python bench.py --steps 1 --warmup 3 --chunk 10 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
