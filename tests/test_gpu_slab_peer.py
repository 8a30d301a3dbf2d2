"""The peer-memory halo path of the x-slab decomposition on ONE GPU: P slabs of a grid stepped in lockstep in one
process (tests/slab_ring.py) against the same grid as a single domain -- all nine fields and the probe series
bit for bit.  The multi-process version of the same check (CUDA IPC, one process per GPU) is tests/test_gpu_slab.py
and the `parity_check` of bench.py's N > 1 arm."""
import numpy as np
import pytest
import torch

from oracle import cases

pytestmark = pytest.mark.gpu
KEYS = ("Ex", "Ey", "Ez", "Dx", "Dy", "Dz", "Hx", "Hy", "Hz")


def _case(shape, npml, steps, seed):
    rng = np.random.default_rng(seed)
    eps = 1 + 2 * rng.random(shape)
    src = [("z", rng.random(shape) * (rng.random(shape) < 0.02), cases.modulated(steps, steps / 3, steps / 8, 9.0, 2.0)),
           ("y", cases.one_hot(shape, (0, 1, 2)), cases.gaussian(steps, steps / 4, steps / 10)),       # in a boundary plane
           ("z", cases.one_hot(shape, (shape[0] - 1, 3, 5)), cases.gaussian(steps, steps / 5, steps / 10)),
           ("x", cases.one_hot(shape, (shape[0] // 2, 0, 0)), cases.gaussian(steps, steps / 5, steps / 10))]
    probes = [("Ez", rng.random(shape)), ("Hy", cases.one_hot(shape, (shape[0] - 1, 2, 1))), ("Dx", rng.random(shape))]
    return dict(eps=eps, dL=cases.DL, npml=list(npml), steps=steps, sources=src, probes=probes)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("shape,npml,P", [((24, 16, 136), (4, 3, 6), 2), ((25, 8, 256), (3, 2, 9), 3), ((16, 12, 128), (0, 2, 3), 4),
                                          ((40, 16, 128), (5, 0, 0), 8), ((256, 128, 64), (20, 20, 20), 2)])
def test_peer_halo_slabs_bit_identical_to_single_domain(shape, npml, P, dtype):
    import ceviche_b200
    from slab_ring import LocalRing
    if dtype == torch.float32 and shape[2] < 128:
        pytest.skip("the fp32 tensor-map kernels need Nz >= 128")
    steps = 30
    case = _case(shape, npml, steps, 11)
    wf = np.stack([w for _, _, w in case["sources"]], 1)
    F = ceviche_b200.fdtd(case["eps"], case["dL"], case["npml"], dtype=dtype)
    series = F.run(steps, case["sources"], case["probes"])
    for opts in ((), (("xchunk", 3), ("tma_rows", 4)), (("xchunk", 64),)):
        ring = LocalRing(case["eps"], case["dL"], case["npml"], P, dtype, options=opts)
        ring.prepare([(c, p) for c, p, _ in case["sources"]], case["probes"])
        half = steps // 2
        got = torch.cat([ring.run(half, wf[:half]), ring.run(steps - half, wf[half:])])
        for k in KEYS:
            assert torch.equal(ring.field(k), F.fields[k]), (k, opts)
        assert float((got - series).abs().max()) <= 1e-11 * float(series.abs().max()), opts
        # a reset and a second run on the same slabs: the counters keep counting, the halos start from zero again
        ring.initialize_fields()
        got2 = ring.run(steps, wf)
        assert float((got2 - series).abs().max()) <= 1e-11 * float(series.abs().max()), opts
        for k in KEYS:
            assert torch.equal(ring.field(k), F.fields[k]), (k, opts, "second run")
        ring.close()
