# multi-GPU: slab tests (both halo transports) + the config-3 bench on N GPUs   (gpurun --gpus N -- 'bash scripts/gpu_multi2.sh N')
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/multi_gpus_$N.txt
nvidia-smi topo -m > gpurun_out/multi_topo_$N.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_slab.py -x -q > gpurun_out/multi_pytest_$N.log 2>&1
tail -5 gpurun_out/multi_pytest_$N.log
for d in f64 f32; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 2 --dtype $d \
   > gpurun_out/multi_bench_${N}_$d.json 2> gpurun_out/multi_bench_${N}_$d.err
tail -c 2500 gpurun_out/multi_bench_${N}_$d.json; tail -5 gpurun_out/multi_bench_${N}_$d.err
done
