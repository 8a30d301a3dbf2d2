"""x-slab decomposition on real GPUs (needs >= 2): NCCL halo exchange, bit-identical to the 1-GPU run."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case(shape, npml, steps, seed):
    rng = np.random.default_rng(seed)
    eps = 1 + 2 * rng.random(shape)
    src = [("z", rng.random(shape) * (rng.random(shape) < 0.02), cases.modulated(steps, steps / 3, steps / 8, 9.0, 2.0)),
           ("y", cases.one_hot(shape, (0, 1, min(2, shape[2] - 1))), cases.gaussian(steps, steps / 4, steps / 10)),
           ("x", cases.one_hot(shape, (shape[0] // 2, 0, 0)), cases.gaussian(steps, steps / 5, steps / 10))]
    probes = [("Ez", rng.random(shape)), ("Hy", cases.one_hot(shape, (shape[0] - 1, 2, min(1, shape[2] - 1)))), ("Dx", rng.random(shape))]
    return dict(eps=eps, dL=cases.DL, npml=list(npml), steps=steps, sources=src, probes=probes)


def _worker(rank, world, port, shape, npml, steps, seed, dtype_name, out, path):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"       # (NCCL bootstrap); the store is a file: no port to collide on
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="file://" + port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        import ceviche_b200
        case = _case(shape, npml, steps, seed)
        # the reference's constructor with the one extra keyword: every rank passes the GLOBAL permittivity
        sim = ceviche_b200.fdtd(case["eps"], case["dL"], case["npml"], dtype=getattr(torch, dtype_name),
                                devices=list(range(world)), path=path)
        assert sim.path == (path or sim.path) and sim.grid_shape == tuple(shape) and sim.t_index == 0
        wf = np.stack([w for _, _, w in case["sources"]], 1)
        half = steps // 2
        s1 = sim.run(half, [(c, p) for c, p, _ in case["sources"]], case["probes"], waveforms=wf[:half])
        s2 = sim.run(steps - half, waveforms=wf[half:])          # prepared sources / probes are kept
        series = torch.cat([s1, s2]).cpu().numpy()
        assert sim.t_index == steps
        fields = {k: sim.fields[k].cpu().numpy() for k in ("Ex", "Ey", "Ez", "Dx", "Dy", "Dz", "Hx", "Hy", "Hz")}
        # a reset and the same run again on the same object
        sim.initialize_fields()
        s3 = sim.run(steps, waveforms=wf).cpu().numpy()
        assert np.array_equal(s3, series), "second run after initialize_fields() differs"
        if rank == 0:
            np.savez(out, series=series, **fields)
    finally:
        dist.destroy_process_group()


def _rendezvous(tmp_path):
    """A fresh file for torch.distributed's FileStore (TCP ports picked in advance can be taken by the time the
    workers bind them: seen once on an 8-GPU box)."""
    return str(tmp_path / "rendezvous")


@pytest.mark.parametrize("dtype_name", ["float64", "float32"])
@pytest.mark.parametrize("shape,npml,path", [((24, 20, 72), (4, 3, 6), "nccl"), ((14, 9, 40), (0, 2, 3), "nccl"),
                                             ((40, 64, 1), (5, 6, 0), "nccl"),        # a 2-D grid (relabelled x, z, y)
                                             ((256, 128, 64), (20, 20, 20), "nccl"),  # the parity grid of BASELINE config 3
                                             ((256, 128, 64), (20, 20, 20), None),    # ... on the default path (peer memory in fp64)
                                             ((64, 32, 136), (6, 5, 7), "peer"), ((256, 128, 128), (20, 20, 20), "peer")])
def test_slabs_bit_identical_to_single_gpu(shape, npml, path, dtype_name, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    import ceviche_b200
    world = torch.cuda.device_count()            # every GPU of the box: 2, 4 or 8 slabs
    while shape[0] // world < 2:
        world //= 2
    steps, seed = 30, 11
    out = str(tmp_path / "slab.npz")
    mp.spawn(_worker, args=(world, _rendezvous(tmp_path), shape, npml, steps, seed, dtype_name, out, path), nprocs=world, join=True)
    got = np.load(out)
    case = _case(shape, npml, steps, seed)
    F = ceviche_b200.fdtd(case["eps"], case["dL"], case["npml"], dtype=getattr(torch, dtype_name))
    series = F.run(steps, case["sources"], case["probes"]).cpu().numpy()
    for k in F.fields:
        assert np.array_equal(got[k], F.fields[k].cpu().numpy()), k
    np.testing.assert_allclose(got["series"], series, rtol=1e-11, atol=1e-12 * np.abs(series).max())
    if shape == (256, 128, 64) and dtype_name == "float64":     # ... and that grid against the CPU oracle
        from oracle.fdtd_c import OracleFDTDC
        from oracle.fdtd_numpy import rel_l2
        O = OracleFDTDC(case["eps"], case["dL"], case["npml"])
        o_series, _ = O.run(steps, case["sources"], case["probes"])
        for k in F.fields:
            assert rel_l2(got[k], O.fields()[k]) <= 1e-10, k
        for p in range(series.shape[1]):
            assert rel_l2(got["series"][:, p], o_series[:, p]) <= 1e-10, p


def _grad_worker(rank, world, store, shape, npml, steps, seed, dtype_name, out, path, region):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="file://" + store, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        import ceviche_b200
        case = _case(shape, npml, steps, seed)
        eps = torch.as_tensor(case["eps"]).cuda().requires_grad_(True)       # the GLOBAL permittivity on every rank
        sim = ceviche_b200.fdtd(eps, case["dL"], case["npml"], dtype=getattr(torch, dtype_name), devices=list(range(world)), path=path)
        sim.design_region = region
        wf = np.stack([w for _, _, w in case["sources"]], 1)
        series = sim.run(steps, [(c, p) for c, p, _ in case["sources"]], case["probes"], waveforms=wf, checkpoint_every=7)
        w = torch.as_tensor(cases.objective_weights(steps, len(case["probes"]))).cuda()
        (g,) = torch.autograd.grad((series ** 2 * w).sum(), eps)
        # the sweep leaves the end-of-run state in place: a plain run continues from it
        with torch.no_grad():
            more = sim.run(4, waveforms=np.zeros((4, wf.shape[1])))
        if rank == 0:
            np.savez(out, grad=g.cpu().numpy(), series=series.detach().cpu().numpy(), more=more.cpu().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dtype_name,shape,npml,path,region", [
    ("float64", (24, 20, 72), (4, 3, 6), "nccl", None),
    ("float64", (64, 32, 136), (6, 5, 7), "peer", None),
    ("float64", (24, 20, 72), (4, 3, 6), "nccl", ((9, 15), (4, 16), (20, 50))),     # design box across the slab boundary
    ("float64", (40, 64, 1), (5, 6, 0), "nccl", None),                               # 2-D
    ("float32", (64, 32, 136), (6, 5, 7), "peer", None)])
def test_slab_gradient_equals_single_gpu_gradient(dtype_name, shape, npml, path, region, tmp_path):
    """run() on x-slabs is differentiable w.r.t. eps_r (adjoint FDTD per slab, cotangent halo planes exchanged between the
    parts of the transposed step, gradient summed over the ranks): equal to the single-GPU gradient to 1e-12."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    import ceviche_b200
    from oracle.fdtd_numpy import rel_l2
    world = torch.cuda.device_count()
    while shape[0] // world < 2:
        world //= 2
    steps, seed = 30, 12
    out = str(tmp_path / "slab_grad.npz")
    mp.spawn(_grad_worker, args=(world, _rendezvous(tmp_path), shape, npml, steps, seed, dtype_name, out, path, region),
             nprocs=world, join=True)
    got = np.load(out)
    case = _case(shape, npml, steps, seed)
    eps = torch.as_tensor(case["eps"]).cuda().requires_grad_(True)
    F = ceviche_b200.fdtd(eps, case["dL"], case["npml"], dtype=getattr(torch, dtype_name))
    series = F.run(steps, case["sources"], case["probes"], checkpoint_every=7)
    w = torch.as_tensor(cases.objective_weights(steps, len(case["probes"]))).cuda()
    (g,) = torch.autograd.grad((series ** 2 * w).sum(), eps)
    g = g.cpu().numpy()
    sl = tuple(slice(lo, hi) for lo, hi in region) if region else slice(None)
    tol = 1e-12 if dtype_name == "float64" else 2e-5
    assert np.abs(g[sl]).max() > 0
    assert rel_l2(got["grad"][sl], g[sl]) <= tol
    with torch.no_grad():
        more = F.run(4, waveforms=np.zeros((4, len(case["sources"]))))
    np.testing.assert_allclose(got["more"], more.cpu().numpy(), rtol=1e-9, atol=1e-12 * np.abs(got["series"]).max())
