"""The multi-GPU slab driver's host logic on CPU: world_size-2 (and 3) gloo groups, numpy compute
backend, against the single-domain oracle.  Covers the ring wrap (npml[0] = 0), uneven partitions,
sources/probes on slab boundaries and two consecutive run() calls."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cases
from oracle.fdtd_numpy import OracleFDTD

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case(shape, npml, steps, seed):
    rng = np.random.default_rng(seed)
    eps = 1 + 2 * rng.random(shape)
    src = [("z", rng.random(shape) * (rng.random(shape) < 0.05), cases.modulated(steps, steps / 3, steps / 8, 9.0, 2.0)),
           ("y", cases.one_hot(shape, (0, 1, min(2, shape[2] - 1))), cases.gaussian(steps, steps / 4, steps / 10)),
           ("x", cases.one_hot(shape, (shape[0] // 2, 0, 0)), cases.gaussian(steps, steps / 5, steps / 10))]
    probes = [("Ez", rng.random(shape)), ("Hy", cases.one_hot(shape, (shape[0] - 1, 2, min(1, shape[2] - 1)))), ("Dx", rng.random(shape))]
    return dict(eps=eps, dL=cases.DL, npml=list(npml), steps=steps, sources=src, probes=probes)


def _worker(rank, world, port, shape, npml, steps, seed, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    dist.init_process_group("gloo", init_method="file://" + port, rank=rank, world_size=world)   # file store: no port races
    try:
        from ceviche_b200.slab import SlabFDTD, partition
        from slab_backend_cpu import NumpySlabBackend
        case = _case(shape, npml, steps, seed)
        lo, hi = partition(shape[0], world)[rank]
        eps = case["eps"]
        eps_local = np.concatenate([eps[(lo - 1) % shape[0]][None], eps[lo:hi]], 0)
        sim = SlabFDTD(shape, eps_local, case["dL"], case["npml"], backend_factory=NumpySlabBackend)
        sim.prepare([(c, p) for c, p, _ in case["sources"]], case["probes"])
        wf = np.stack([w for _, _, w in case["sources"]], 1)
        half = steps // 2
        s1 = sim.run(half, wf[:half])
        s2 = sim.run(steps - half, wf[half:])
        series = torch.cat([s1, s2]).numpy()
        fields = {k: sim.gather(k).numpy() for k in ("Ex", "Ey", "Ez", "Dx", "Dy", "Dz", "Hx", "Hy", "Hz")}
        if rank == 0:
            np.savez(out, series=series, **fields)
    finally:
        dist.destroy_process_group()


def _rendezvous(tmp_path):
    """A fresh file for torch.distributed's FileStore (TCP ports picked in advance can be taken by the time the
    workers bind them: seen once on an 8-GPU box)."""
    return str(tmp_path / "rendezvous")


@pytest.mark.parametrize("world,shape,npml", [(2, (12, 7, 6), (3, 2, 2)), (2, (9, 6, 5), (0, 2, 0)), (3, (11, 5, 6), (2, 0, 2)),
                                              (2, (10, 9, 1), (2, 3, 0))])      # 2-D grid
def test_slab_driver_matches_single_domain_oracle(world, shape, npml, tmp_path):
    steps, seed = 40, 5
    out = str(tmp_path / "slab.npz")
    mp.spawn(_worker, args=(world, _rendezvous(tmp_path), shape, npml, steps, seed, out), nprocs=world, join=True)
    got = np.load(out)
    case = _case(shape, npml, steps, seed)
    O = OracleFDTD(case["eps"], case["dL"], case["npml"])
    o_series, _ = O.run(steps, case["sources"], case["probes"])
    for k, v in O.fields().items():
        assert np.array_equal(got[k], v), k          # no arithmetic is reordered by the decomposition
    np.testing.assert_allclose(got["series"], o_series, rtol=1e-12, atol=1e-12 * np.abs(o_series).max())


def test_partition_and_localize():
    from ceviche_b200.slab import localize_points, partition
    assert partition(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert partition(8, 8) == [(i, i + 1) for i in range(8)]
    a = np.zeros((6, 2, 3))
    a[2, 1, 2], a[3, 0, 0], a[5, 1, 1] = 1.5, -2.0, 3.0
    idx, w = localize_points(a, 2, 4, 6)
    assert idx.tolist() == [5, 6] and w.tolist() == [1.5, -2.0]
    idx, w = localize_points(a, 4, 6, 6)
    assert idx.tolist() == [6 + 4] and w.tolist() == [3.0]


def _grad_chain_worker(rank, world, port, shape, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    dist.init_process_group("gloo", init_method="file://" + port, rank=rank, world_size=world)
    try:
        from ceviche_b200.slab import SlabFDTD, _AllReduceGrad, partition
        from slab_backend_cpu import NumpySlabBackend
        rng = np.random.default_rng(3)
        eps = torch.as_tensor(1 + 2 * rng.random(shape)).requires_grad_(True)
        G = torch.as_tensor(rng.standard_normal((3,) + shape))           # stands for dL/d(1/eps_yee) of the adjoint sweep
        lo, hi = partition(shape[0], world)[rank]
        idx = torch.arange(lo - 1, hi) % shape[0]
        sim = SlabFDTD(shape, _AllReduceGrad.apply(eps, None)[idx], cases.DL, [2, 0, 2], backend_factory=NumpySlabBackend)
        loss = sum((sim._mE64[c] * G[c][lo:hi]).sum() for c in range(3))
        (g,) = torch.autograd.grad(loss, eps)
        np.save(out + ".%d.npy" % rank, g.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape", [(2, (8, 5, 4)), (3, (10, 4, 3))])
def test_slab_gradient_chain_to_global_eps(world, shape, tmp_path):
    """The host side of gradients on x-slabs: every rank differentiates through its slab's eps slice (with the extra plane
    lo - 1 of the Yee average along x, utils.py:167, periodic) and the all-reduce in the backward of the slicing hands every
    rank the complete gradient w.r.t. the global eps_r -- equal to differentiating the single-domain expression."""
    out = str(tmp_path / "g")
    mp.spawn(_grad_chain_worker, args=(world, _rendezvous(tmp_path), shape, out), nprocs=world, join=True)
    rng = np.random.default_rng(3)
    eps = torch.as_tensor(1 + 2 * rng.random(shape)).requires_grad_(True)
    G = torch.as_tensor(rng.standard_normal((3,) + shape))
    inv = [1 / ((eps + torch.roll(eps, 1, a)) / 2) for a in range(3)]
    (want,) = torch.autograd.grad(sum((inv[c] * G[c]).sum() for c in range(3)), eps)
    for r in range(world):
        got = np.load(out + ".%d.npy" % r)
        np.testing.assert_allclose(got, want.numpy(), rtol=1e-13, atol=1e-13)
