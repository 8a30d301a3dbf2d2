mkdir -p gpurun_out; rm -f gpurun_out/slab_prio.log
for prio in 0 1; do for d in f64 f32; do
 CEV_SLAB_PRIO=$prio timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$prio bench.py --gpus 2 --steps 3 --warmup 3 --slab-chunk 50 --slab-grid 256 1024 512 --dtype $d 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('prio=$prio $d', 'value %.2f per-GPU %.2f ms/step %.3f' % (j['value'], j['value']/2, j['ms_per_step']/50), 'H-kernel-alone GB/s %.0f' % j['roofline']['achieved'])" >> gpurun_out/slab_prio.log
done; done
timeout 600 python -m pytest tests/test_gpu_slab.py -x -q 2>&1 | tail -2 >> gpurun_out/slab_prio.log
cat gpurun_out/slab_prio.log
