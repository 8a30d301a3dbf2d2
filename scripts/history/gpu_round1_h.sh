mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 20 -c 2 -o gpurun_out/prof_r1b_f64_256 python bench.py --steps 1 --warmup 3 --chunk 10 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 20 -c 2 -o gpurun_out/prof_r1b_f32_256 python bench.py --steps 1 --warmup 3 --chunk 10 --no-cpu --no-e2e --dtype f32 > gpurun_out/ncu_full32.log 2>&1
for d in f64 f32; do timeout 300 python scripts/tune.py 256 $d "xchunk=0" >> gpurun_out/tune5.log 2>&1; done
cat gpurun_out/tune5.log
