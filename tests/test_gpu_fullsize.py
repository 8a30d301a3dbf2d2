"""Parity at BASELINE.json's full size (config 2: 256^3, npml 20, the splitter geometry of bench.py):
the first time steps against the numpy oracle (the reference costs ~3 s per step at this size), and
size-independent exact properties over more steps: the tuned / TMA-staged / fused kernels bit-identical to the
one-thread-per-cell kernel at the real launch geometry, and exact linearity under a power-of-two source scale."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle.fdtd_numpy import FIELD_KEYS, OracleFDTD, rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPE = (256, 256, 256)
NPML = [20, 20, 20]
DL = 5e-8


def _workload(steps, amp=1.0):
    sys.path.insert(0, ROOT)
    import bench
    wl = bench.workload(SHAPE, steps)
    prof = wl["sources"][0][1]
    t = np.arange(steps)
    wave = amp * np.cos(0.3 * t) * (1 - np.exp(-(t + 1) / 3.0))          # strong from the first step on
    near = np.zeros(SHAPE); near[30:32, 121:135, 123:133] = 1.0           # probes the wave reaches within a few steps
    pml_probe = np.zeros(SHAPE); pml_probe[29:32, 118:138, 238:256] = 1.0  # ... one of them inside the z-PML,
    prof_pml = np.roll(prof, 120, axis=2)                                  # next to a second source inside the z-PML
    return wl["eps"], [("z", prof, wave), ("y", prof_pml, 0.5 * wave)], \
        [("Ez", near), ("Hy", near), ("Ey", pml_probe)] + wl["probes"]


def _run(eps, sources, probes, steps, dtype, **options):
    import ceviche_b200
    F = ceviche_b200.fdtd(eps, DL, NPML, dtype=dtype)
    for k, v in options.items():
        F.set_option(k, v)
    series = F.run(steps, sources, probes)
    return F, series.cpu().numpy()


def test_config2_full_size_first_steps_against_oracle():
    steps = 6
    eps, sources, probes = _workload(steps)
    F, series = _run(eps, sources, probes, steps, torch.float64)
    O = OracleFDTD(eps, DL, NPML)
    o_series, _ = O.run(steps, sources, probes)
    of = O.fields()
    for k in FIELD_KEYS:
        assert rel_l2(F.fields[k].cpu().numpy(), of[k]) <= 1e-10, k
    for p in range(3):
        assert np.abs(o_series[:, p]).max() > 0
        assert rel_l2(series[:, p], o_series[:, p]) <= 1e-10, p
    F32, s32 = _run(eps, sources, probes, steps, torch.float32)
    for k in FIELD_KEYS:
        assert rel_l2(F32.fields[k].cpu().numpy(), of[k]) <= 1e-5, k


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_full_size_kernel_variants_bitwise(dtype):
    steps = 24
    eps, sources, probes = _workload(steps)
    ref, s_ref = _run(eps, sources, probes, steps, dtype, kernel_variant=1)
    ref_f = {k: ref.fields[k].clone() for k in FIELD_KEYS}
    ref_p = [t.clone() for fam in ("ICE", "IH", "ICH", "ID") for t in ref._pml[fam]]
    del ref
    for opts in (dict(kernel_variant=0), dict(kernel_variant=2, split_launch=1), dict(kernel_variant=3), dict(kernel_variant=4),
                 dict(kernel_variant=5)):
        F, s = _run(eps, sources, probes, steps, dtype, **opts)
        assert np.array_equal(s, s_ref), opts
        for k in FIELD_KEYS:
            assert torch.equal(F.fields[k], ref_f[k]), (opts, k)
        for a, b in zip([t for fam in ("ICE", "IH", "ICH", "ID") for t in F._pml[fam]], ref_p):
            assert torch.equal(a, b), opts
        del F
        torch.cuda.empty_cache()


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_full_size_linearity_is_exact_for_power_of_two_scales(dtype):
    """Every operation of the step is linear, and scaling by 2^k is exact in binary floating point: a source
    4 x stronger must give fields and series exactly 4 x larger (bit for bit), PML and all."""
    steps = 40
    eps, sources, probes = _workload(steps)
    F1, s1 = _run(eps, sources, probes, steps, dtype)
    f1 = {k: F1.fields[k].clone() for k in FIELD_KEYS}
    del F1
    F4, s4 = _run(eps, [(c, p, 4.0 * w) for c, p, w in sources], probes, steps, dtype)
    assert np.abs(s1).max() > 0
    if dtype == torch.float64:
        assert np.array_equal(s4, 4.0 * s1)
    else:
        np.testing.assert_allclose(s4, 4.0 * s1, rtol=1e-12, atol=1e-30)   # (fp64 sums of fp32 values incl. subnormals)
    for k in FIELD_KEYS:
        a, b = F4.fields[k], 4.0 * f1[k]
        if dtype == torch.float64:
            assert torch.equal(a, b), k
        else:
            # fp32 subnormals (intermediates at the leading edge of the wave front, < 1.2e-38) do not scale exactly,
            # and their 1e-45-sized errors can reach the last bit of values just above them: exactness is asserted
            # well clear of that edge, agreement to ~subnormal spacing below it
            clear = b.abs() >= 1e-25
            assert int(clear.sum()) > 1000, k
            assert torch.equal(a[clear], b[clear]), k
            assert float((a - b)[~clear].abs().max()) <= 1e-32, k
