N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m pytest tests/test_gpu_slab.py -x -q > gpurun_out/pytest_slab.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_slab.log
tail -25 gpurun_out/pytest_slab.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 --slab-chunk 50 > gpurun_out/bench_slab_${N}_f64.json 2> gpurun_out/bench_slab_${N}_f64.err
tail -3 gpurun_out/bench_slab_${N}_f64.err; cat gpurun_out/bench_slab_${N}_f64.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 3 --slab-chunk 50 --dtype f32 > gpurun_out/bench_slab_${N}_f32.json 2> gpurun_out/bench_slab_${N}_f32.err
tail -3 gpurun_out/bench_slab_${N}_f32.err; cat gpurun_out/bench_slab_${N}_f32.json
