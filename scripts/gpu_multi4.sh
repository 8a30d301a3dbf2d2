N=${1:-4}
mkdir -p gpurun_out
for d in f64 f32; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 2 --dtype $d \
   > gpurun_out/multi_bench_${N}_$d.json 2> gpurun_out/multi_bench_${N}_$d.err
tail -c 600 gpurun_out/multi_bench_${N}_$d.json; tail -2 gpurun_out/multi_bench_${N}_$d.err
done
