mkdir -p gpurun_out; rm -f gpurun_out/tune_v3.log
timeout 240 python -m pytest tests/test_gpu_variants.py -x -q > gpurun_out/pytest_v3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_v3.log
tail -25 gpurun_out/pytest_v3.log
for d in f64 f32; do for n in 256 512; do timeout 120 python scripts/tune.py $n $d "kernel_variant=2" "kernel_variant=3,xchunk=0" "xchunk=8" "xchunk=32" "xchunk=4" >> gpurun_out/tune_v3.log 2>&1; done; done
cat gpurun_out/tune_v3.log
