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


def _bounding_box(arr):
    nz = np.nonzero(arr)
    return tuple(slice(int(i.min()), int(i.max()) + 1) for i in nz)


def test_config2_full_size_first_100_steps_against_c_oracle():
    """config 2's parity at its full size over the first 100 time steps.  The reference itself needs ~3 s per
    step here, so the CPU side is its C restatement (oracle/fdtd_c.c: bit-identical to the numpy port, which is bit-identical
    to ceviche/fdtd.py -- tests/test_oracle_c.py, test_oracle_vs_reference.py) on all host cores; sources are written and
    probes summed on their bounding boxes only."""
    from oracle.fdtd_c import OracleFDTDC, set_threads
    steps = 100
    eps, sources, probes = _workload(steps)
    F, series = _run(eps, sources, probes, steps, torch.float64)
    set_threads(os.cpu_count() or 1)
    O = OracleFDTDC(eps, DL, NPML)
    J = {comp: np.zeros(SHAPE) for comp, _, _ in sources}
    src = [(comp, prof, wave, _bounding_box(prof)) for comp, prof, wave in sources]
    prb = [(key, mask, _bounding_box(mask)) for key, mask in probes]
    o_series = np.zeros((steps, len(probes)))
    for t in range(steps):
        for comp, prof, wave, box in src:
            J[comp][box] = prof[box] * wave[t]
        f = O.step(**{"J" + comp: a for comp, a in J.items()})
        for p, (key, mask, box) in enumerate(prb):
            o_series[t, p] = np.sum(f[key][box] * mask[box])
    of = O.fields()
    for k in FIELD_KEYS:
        assert np.linalg.norm(of[k]) > 0, k
        assert rel_l2(F.fields[k].cpu().numpy(), of[k]) <= 1e-10, k
    for p in range(len(probes)):
        if np.abs(o_series[:, p]).max() > 0:          # (the arm probes at x = 226 are not reached within 100 steps)
            assert rel_l2(series[:, p], o_series[:, p]) <= 1e-10, p
        else:
            assert np.abs(series[:, p]).max() == 0, p
    assert sum(np.abs(o_series[:, p]).max() > 0 for p in range(len(probes))) >= 3
    del F
    F32, s32 = _run(eps, sources, probes, steps, torch.float32)
    for k in FIELD_KEYS:
        assert rel_l2(F32.fields[k].cpu().numpy(), of[k]) <= 1e-5, k
    for p in range(3):
        assert rel_l2(s32[:, p], o_series[:, p]) <= 1e-5, p


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


def test_config4_full_size_gradient_properties():
    """Config 4 at its stated size (256 x 256 x 128, npml 20, 64 x 64 x 32 design box): size-independent checks of the
    reverse sweep.  (i) the design-box gradient from the D-box record (tensor-map kernels, no recomputation)
    equals the every-cell gradient (checkpoints + recomputation) inside the box; (ii) <grad, v> equals the directional
    derivative of the forward-mode tangent sweep for a random v in the box; (iii) ... and a central finite difference of
    the forward run."""
    import ceviche_b200
    shape, steps = (256, 256, 128), 320
    box = ((96, 160), (96, 160), (48, 80))
    sl = tuple(slice(lo, hi) for lo, hi in box)
    rng = np.random.default_rng(1)
    eps_np = np.ones(shape)
    eps_np[sl] = 1 + 4.95 * rng.random((64, 64, 32))
    prof = np.zeros(shape); prof[84, 123:133, 61:67] = 1.0                 # sheet source just before the box
    mask = np.zeros(shape); mask[170, 118:138, 58:70] = rng.random((20, 12))    # weighted probe just behind it
    t = np.arange(steps)
    wave = 5 * np.exp(-(t - 60) ** 2 / (2 * 20.0 ** 2)) * np.cos(0.25 * t)
    srcs, probes = [("z", prof, wave)], [("Ez", mask)]
    w = torch.as_tensor(rng.random((steps, 1))).cuda()

    def loss_of(eps_t, region):
        F = ceviche_b200.fdtd(eps_t, DL, NPML)
        F.design_region = region
        return (F.run(steps, srcs, probes) ** 2 * w).sum()

    grads = {}
    for region in (None, box):
        eps = torch.as_tensor(eps_np).cuda().requires_grad_(True)
        L = loss_of(eps, region)
        (g,) = torch.autograd.grad(L, eps)
        grads[region is not None] = g[sl].clone()
        del g, eps
        torch.cuda.empty_cache()
    assert float(grads[True].abs().max()) > 0
    assert rel_l2(grads[True].cpu().numpy(), grads[False].cpu().numpy()) <= 1e-12
    # (ii) the tangent sweep
    v = torch.zeros(shape, dtype=torch.float64, device="cuda")
    v[sl] = torch.as_tensor(rng.standard_normal((64, 64, 32))).cuda()
    F = ceviche_b200.fdtd(eps_np, DL, NPML)
    series, dseries = F.jvp_run(steps, v[None], srcs, probes)
    d_fwd = float((2 * series * w * dseries[0]).sum())
    d_rev = float((grads[True] * v[sl]).sum())
    assert abs(d_fwd - d_rev) <= 1e-10 * abs(d_fwd), (d_fwd, d_rev)
    # (iii) central finite difference of the forward run
    h = 1e-5
    with torch.no_grad():
        eps0 = torch.as_tensor(eps_np).cuda()
        fd = (float(loss_of(eps0 + h * v, None)) - float(loss_of(eps0 - h * v, None))) / (2 * h)
    assert abs(fd - d_rev) <= 1e-6 * abs(fd), (fd, d_rev)


def test_config5_full_size_batched_tangents():
    """Config 5 at its stated grid (2-D 2048 x 2048 grating coupler, 16 fill-factor directions through the sigmoid
    projection, examples/forwardmode_grating_coupler.py:138-162): the batched tangent half-steps and the fused tangent
    step are bit-identical to one launch per tangent, and a tangent equals a central finite difference of the forward run."""
    import ceviche_b200
    from ceviche_b200.parametrization import grating_coupler
    steps, B = 260, 16
    G = grating_coupler(2048, 2048, DL, 20, groups=B)
    ff = torch.full((B,), 0.5, dtype=torch.float64, device="cuda")
    eps, V = G.eps_r(ff), G.fill_factor_directions(ff)
    shape = (2048, 2048, 1)
    prof = np.zeros(shape); prof[1000, G.y_base[0]:G.y_teeth[1], 0] = 1.0          # sheet across the slab, mid-grating
    mask = np.zeros(shape); mask[900:1100, G.y_teeth[1] + 6, 0] = 1.0                # line probe just above the teeth
    t = np.arange(steps)
    wave = np.exp(-(t - 60) ** 2 / (2 * 20.0 ** 2)) * np.cos(0.3 * t)
    srcs, probes = [("z", prof, wave)], [("Ez", mask)]
    out = {}
    for name, opts in (("one launch per tangent", {"jvp_fused": 0, "jvp_batch": 0}), ("batched half-steps", {"jvp_fused": 0}),
                       ("fused tangent step", {})):
        F = ceviche_b200.fdtd(eps, DL, [20, 20, 0])
        for k, v in opts.items():
            F.set_option(k, v)
        out[name] = F.jvp_run(steps, V, srcs, probes)
        assert F._active == 0b011100
    ref = out["one launch per tangent"]
    for name in ("batched half-steps", "fused tangent step"):
        assert torch.equal(out[name][0], ref[0]) and torch.equal(out[name][1], ref[1]), name
    series, dseries = out["fused tangent step"]
    norms = dseries.flatten(1).norm(dim=1)
    assert float(norms.max()) > 0
    b = int(norms.argmax())                                   # the tooth group under the source / probe
    h = 1e-5
    runs = []
    for sgn in (+1, -1):
        F = ceviche_b200.fdtd(eps + sgn * h * V[b], DL, [20, 20, 0])
        runs.append(F.run(steps, srcs, probes))
    fd = (runs[0] - runs[1]) / (2 * h)
    assert rel_l2(dseries[b].cpu().numpy(), fd.cpu().numpy()) <= 1e-6
