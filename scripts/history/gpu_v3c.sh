mkdir -p gpurun_out; rm -f gpurun_out/tune_v3c.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
for d in f64 f32; do for n in 256 512; do timeout 120 python scripts/tune.py $n $d "kernel_variant=0" "kernel_variant=2" >> gpurun_out/tune_v3c.log 2>&1; done; done
cat gpurun_out/tune_v3c.log
timeout 600 python bench.py --no-cpu > gpurun_out/bench_f64_v3.json 2> gpurun_out/bench_f64_v3.err
timeout 600 python bench.py --grid 512 512 512 --steps 4 --warmup 3 --chunk 100 --no-cpu > gpurun_out/bench_f64_512_v3.json 2> gpurun_out/bench_f64_512_v3.err
