# first run of the tensor-map TMA kernels: bit-identity tests, then an A/B against the round-1 default
mkdir -p gpurun_out; rm -f gpurun_out/v5_*.log
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/v5_gpu.txt
timeout 900 python -m pytest tests/test_gpu_variants.py -x -q -k "tensor_map" > gpurun_out/v5_pytest.log 2>&1
tail -5 gpurun_out/v5_pytest.log
for d in f64 f32; do for n in 256 512; do
  timeout 600 python scripts/tune.py $n $d "kernel_variant=0" "kernel_variant=6,tma_rows=4,tma_stages=4" "tma_rows=4,tma_stages=3" \
     "tma_rows=8,tma_stages=3" "tma_rows=8,tma_stages=4" "tma_rows=4,tma_stages=4,xchunk=8" "xchunk=32" "tma_rows=4,tma_stages=3,xchunk=8" >> gpurun_out/v5_tune.log 2>&1
done; done
cat gpurun_out/v5_tune.log
TUNE_NPML=0 timeout 300 python scripts/tune.py 256 f64 "kernel_variant=0" "kernel_variant=6,tma_rows=4,tma_stages=4" "tma_rows=4,tma_stages=3" >> gpurun_out/v5_tune_nopml.log 2>&1
cat gpurun_out/v5_tune_nopml.log
