# N GPUs: config-3 bench (both dtypes) + the parity-grid slab tests   (gpurun --gpus N -- 'bash scripts/gpu_multi8.sh N')
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/multi_gpus_$N.txt
for d in f64 f32; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 2 --dtype $d \
   > gpurun_out/multi_bench_${N}_$d.json 2> gpurun_out/multi_bench_${N}_$d.err
tail -c 1800 gpurun_out/multi_bench_${N}_$d.json; tail -3 gpurun_out/multi_bench_${N}_$d.err
done
timeout 900 python -m pytest tests/test_gpu_slab.py -x -q -k "256 or gradient" > gpurun_out/multi_pytest_$N.log 2>&1
tail -5 gpurun_out/multi_pytest_$N.log
