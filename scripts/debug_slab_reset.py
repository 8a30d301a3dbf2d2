"""2-GPU debug: where does a re-run after initialize_fields() differ from the first (split) run?"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def worker(rank, world, store, path, shape, npml):
    import torch.distributed as dist
    from test_gpu_slab import _case
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="file://" + store, rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import ceviche_b200
    steps = 30
    case = _case(shape, npml, steps, 11)
    sim = ceviche_b200.fdtd(case["eps"], case["dL"], case["npml"], devices=list(range(world)), path=path)
    wf = np.stack([w for _, _, w in case["sources"]], 1)
    half = steps // 2
    srcs = [(c, p) for c, p, _ in case["sources"]]
    s1 = sim.run(half, srcs, case["probes"], waveforms=wf[:half])
    s2 = sim.run(steps - half, waveforms=wf[half:])
    series = torch.cat([s1, s2]).cpu().numpy()
    sim.initialize_fields()
    s3 = sim.run(steps, waveforms=wf).cpu().numpy()
    sim.initialize_fields()
    s4 = sim.run(steps, waveforms=wf).cpu().numpy()
    sim.initialize_fields()
    s5 = torch.cat([sim.run(half, waveforms=wf[:half]), sim.run(steps - half, waveforms=wf[half:])]).cpu().numpy()
    if rank == 0:
        for name, s in (("s3 (one run)", s3), ("s4 (one run again)", s4), ("s5 (split again)", s5)):
            d = np.abs(s - series)
            bad = np.argwhere(d > 0)
            print(path, shape, name, "max diff", d.max(), "n_bad", len(bad), "first", bad[:6].tolist(), flush=True)
        print("s3 vs s4 equal:", np.array_equal(s3, s4), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    import tempfile
    import torch.multiprocessing as mp
    for path, shape, npml in (("nccl", (24, 20, 72), (4, 3, 6)), ("peer", (64, 32, 136), (6, 5, 7))):
        with tempfile.TemporaryDirectory() as d:
            mp.spawn(worker, args=(2, os.path.join(d, "rv"), path, shape, npml), nprocs=2, join=True)
