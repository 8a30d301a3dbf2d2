"""The tuned marching kernels (step_v2.cuh) against the baseline one-thread-per-cell kernels
(step_v1.cuh): bit-identical fields, PML integrals and probe series, for every x-chunking."""
import numpy as np
import pytest
import torch

from oracle import cases
from oracle.fdtd_numpy import FIELD_KEYS, OracleFDTD, rel_l2

pytestmark = pytest.mark.gpu


def _case(shape, npml, steps, seed):
    rng = np.random.default_rng(seed)
    eps = 1 + 3 * rng.random(shape)
    src = [("z", rng.random(shape) * (rng.random(shape) < 0.02), cases.modulated(steps, steps / 3, steps / 8, 11.0, 2.0)),
           ("x", cases.one_hot(shape, (shape[0] // 2, shape[1] // 3, shape[2] // 2)), cases.gaussian(steps, steps / 4, steps / 10)),
           ("y", cases.one_hot(shape, (0, 0, shape[2] - 1)), cases.gaussian(steps, steps / 5, steps / 10))]
    probes = [("Ez", rng.random(shape)), ("Hx", rng.random(shape) * (rng.random(shape) < 0.2)),
              ("Dy", cases.one_hot(shape, (shape[0] - 1, shape[1] - 1, 0)))]
    return dict(eps=eps, dL=cases.DL, npml=list(npml), steps=steps, sources=src, probes=probes)


def _run(case, dtype, variant, xchunk=0, per_step=False, lanes_z=8, prefetch=1, split=0, fused_shape=0, steps=None,
         tma_rows=4, tma_stages=4):
    import ceviche_b200
    if steps is not None:
        case = dict(case, steps=steps, sources=[(c, p, w[:steps]) for c, p, w in case["sources"]])
    F = ceviche_b200.fdtd(case["eps"], case["dL"], case["npml"], dtype=dtype)
    F.set_option("kernel_variant", variant)
    F.set_option("fused_shape", fused_shape)
    F.set_option("xchunk", xchunk)
    F.set_option("lanes_z", lanes_z)
    F.set_option("prefetch_planes", prefetch)
    F.set_option("split_launch", split)
    F.set_option("tma_rows", tma_rows)
    F.set_option("tma_stages", tma_stages)
    if per_step:
        profs = [(c, torch.as_tensor(p).cuda()) for c, p, _ in case["sources"]]
        for t in range(case["steps"]):
            J = {c: p * float(w[t]) for (c, p), (_, _, w) in zip(profs, case["sources"])}
            F.forward(Jx=J.get("x"), Jy=J.get("y"), Jz=J.get("z"))
        series = None
    else:
        series = F.run(case["steps"], case["sources"], case["probes"]).cpu().numpy()
    fields = {k: F.fields[k].cpu().numpy() for k in FIELD_KEYS}
    pml = [t.cpu().numpy() for fam in ("ICE", "IH", "ICH", "ID") for t in F._pml[fam]]
    return series, fields, pml


SHAPES = [((20, 18, 136), (4, 3, 6)), ((9, 7, 64), (2, 0, 5)), ((3, 5, 24), (0, 2, 2)), ((33, 40, 1), (5, 6, 0)),
          ((6, 10, 260), (2, 3, 20)), ((44, 38, 136), (6, 5, 8)), ((40, 48, 100), (5, 0, 7))]   # the last two split interior / shell


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("shape,npml", SHAPES)
def test_marching_equals_baseline_bitwise(shape, npml, dtype):
    case = _case(shape, npml, 40, 7)
    s1, f1, p1 = _run(case, dtype, 1)
    s0, f0, p0 = _run(case, dtype, 0)          # the default (auto) kernel choice
    for k in FIELD_KEYS:
        assert np.array_equal(f1[k], f0[k]), (k, 'auto')
    assert np.array_equal(s1, s0)
    for xchunk, lanes_z, pf, split in ((0, 8, 1, 1), (1, 16, 0, 1), (3, 32, 2, 0), (1000, 8, 5, 1), (5, 32, 1, 1), (0, 8, 1, 0)):
        s2, f2, p2 = _run(case, dtype, 2, xchunk, lanes_z=lanes_z, prefetch=pf, split=split)
        for k in FIELD_KEYS:
            assert np.array_equal(f1[k], f2[k]), (k, xchunk)
        for a, b in zip(p1, p2):
            assert np.array_equal(a, b), xchunk
        assert np.array_equal(s1, s2)
    for xchunk in (0, 1, 5):       # TMA-staged kernels (kernel_variant = 3)
        s3, f3, p3 = _run(case, dtype, 3, xchunk)
        for k in FIELD_KEYS:
            assert np.array_equal(f1[k], f3[k]), (k, "v3", xchunk)
        for a, b in zip(p1, p3):
            assert np.array_equal(a, b), ("v3", xchunk)
        assert np.array_equal(s1, s3)


# grids the tensor-map TMA kernels serve (Ny a multiple of the tile rows, Nz >= one tile row of 32 vectors), incl. a
# partial last z-tile, wide / absent PML on single axes and no PML at all; the last two are not eligible (fallback)
TMA_SHAPES = [((20, 16, 136), (4, 3, 6)), ((6, 8, 260), (2, 3, 20)), ((9, 24, 128), (2, 0, 5)), ((40, 48, 192), (5, 7, 0)),
              ((18, 16, 128), (0, 0, 0)), ((3, 8, 256), (0, 2, 9)), ((35, 32, 384), (6, 5, 8)),
              ((12, 10, 136), (2, 3, 4)), ((12, 16, 40), (2, 3, 4))]


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("shape,npml", TMA_SHAPES)
def test_tensor_map_kernels_equal_baseline_bitwise(shape, npml, dtype):
    """step_v5.cuh (cp.async.bulk.tensor box copies, lean PML update) against the one-thread-per-cell kernels:
    fields, the twelve PML integral arrays and the probe series, for both tile heights, both ring depths and
    several x-chunkings (chunk ends exercise the x+1 / x-1 neighbour-only stages and the periodic wrap planes)."""
    case = _case(shape, npml, 40, 7)
    s1, f1, p1 = _run(case, dtype, 1)
    for rows, stages, xchunk in ((4, 4, 0), (4, 3, 1), (8, 4, 3), (8, 3, 5), (4, 4, 2), (4, 3, 1000), (8, 4, 7)):
        s6, f6, p6 = _run(case, dtype, 6, xchunk, tma_rows=rows, tma_stages=stages)
        for k in FIELD_KEYS:
            assert np.array_equal(f1[k], f6[k]), (k, rows, stages, xchunk)
        for q, (a, b) in enumerate(zip(p1, p6)):
            assert np.array_equal(a, b), (q, rows, stages, xchunk)
        assert np.array_equal(s1, s6), (rows, stages, xchunk)


FUSED_SHAPES = SHAPES + [((5, 3, 8), (1, 1, 2)), ((1, 6, 16), (0, 2, 3)), ((17, 1, 36), (3, 0, 4)), ((36, 31, 124), (5, 4, 6)),
                         ((20, 9, 4), (3, 2, 0)), ((18, 14, 72), (0, 0, 0)),    # no PML at all = the lean instantiation
                         ((44, 40, 96), (6, 5, 8)), ((40, 30, 72), (0, 4, 6)), ((52, 24, 64), (7, 0, 0))]   # hybrid boxes


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("shape,npml", FUSED_SHAPES)
def test_fused_step_equals_baseline_bitwise(shape, npml, dtype):
    """The fused full-step kernel (step_v4.cuh, one launch per time step, ping-ponged state) against the
    one-thread-per-cell kernels: fields, PML integrals and probe series, even and odd step counts, every
    tile shape and several x-chunkings (the pre-roll plane, halo rows / lanes and periodic wraps all differ)."""
    case = _case(shape, npml, 41, 7)
    ref = {n: _run(case, dtype, 1, steps=n) for n in (41, 40, 2)}
    for fs, xchunk, n in ((0, 0, 41), (1604, 3, 40), (804, 1, 41), (1608, 1000, 40), (3204, 5, 41), (3208, 2, 2), (0, 7, 2),
                          (-5, 0, 41), (-5, 6, 40)):       # fs = -5: the hybrid path (kernel_variant 5)
        s1, f1, p1 = ref[n]
        s4, f4, p4 = _run(case, dtype, 5 if fs < 0 else 4, xchunk, fused_shape=max(fs, 0), steps=n)
        for k in FIELD_KEYS:
            assert np.array_equal(f1[k], f4[k]), (k, fs, xchunk, n)
        for q, (a, b) in enumerate(zip(p1, p4)):
            assert np.array_equal(a, b), (q, fs, xchunk, n)
        assert np.array_equal(s1, s4), (fs, xchunk, n)


def test_fused_step_keeps_running_across_calls_and_against_oracle():
    case = _case((24, 22, 72), (4, 3, 5), 60, 5)
    import ceviche_b200
    F = ceviche_b200.fdtd(case["eps"], case["dL"], case["npml"])
    F.set_option("kernel_variant", 4)
    parts = []
    for t0, t1 in ((0, 7), (7, 8), (8, 30), (30, 60)):          # odd / single / even legs on the same object
        parts.append(F.run(t1 - t0, [(c, p, w[t0:t1]) for c, p, w in case["sources"]], case["probes"]).cpu().numpy())
    series = np.concatenate(parts)
    O = OracleFDTD(case["eps"], case["dL"], case["npml"])
    o_series, _ = O.run(case["steps"], case["sources"], case["probes"])
    for k in FIELD_KEYS:
        assert rel_l2(F.fields[k].cpu().numpy(), O.fields()[k]) <= 1e-10, k
    for p in range(series.shape[1]):
        assert rel_l2(series[:, p], o_series[:, p]) <= 1e-10


def test_marching_forward_api_with_dense_J_and_E_output():
    case = _case((10, 12, 40), (2, 3, 4), 25, 9)
    _, f1, _ = _run(case, torch.float64, 1, per_step=True)
    _, f2, _ = _run(case, torch.float64, 2, xchunk=4, per_step=True)
    _, f3, _ = _run(case, torch.float64, 2)
    for k in FIELD_KEYS:
        assert np.array_equal(f1[k], f2[k]), k
        assert np.array_equal(f1[k], f3[k]), k


def test_marching_against_oracle_midsize():
    """A grid large enough for several CTAs per axis and several x-chunks, against the numpy oracle."""
    case = _case((40, 36, 136), (6, 5, 8), 60, 3)
    O = OracleFDTD(case["eps"], case["dL"], case["npml"])
    o_series, _ = O.run(case["steps"], case["sources"], case["probes"])
    series, fields, _ = _run(case, torch.float64, 2, xchunk=7)
    for k in FIELD_KEYS:
        assert rel_l2(fields[k], O.fields()[k]) <= 1e-10, k
    for p in range(series.shape[1]):
        assert rel_l2(series[:, p], o_series[:, p]) <= 1e-10
    series32, fields32, _ = _run(case, torch.float32, 2)
    for k in FIELD_KEYS:
        assert rel_l2(fields32[k], O.fields()[k]) <= 1e-5, k


# ---- provably-zero components skipped on 2-D / 1-D grids (active_components)
def _polarised_case(shape, npml, comp, steps=120, seed=4):
    rng = np.random.default_rng(seed)
    eps = 1 + 3 * rng.random(shape)
    mid = tuple(n // 2 for n in shape)
    src = [(comp, cases.one_hot(shape, mid, 3.0), cases.modulated(steps, steps / 3, steps / 8, 9.0, 2.0)),
           (comp, rng.random(shape) * (rng.random(shape) < 0.05), cases.gaussian(steps, steps / 4, steps / 10))]
    probes = [(k, rng.random(shape)) for k in ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz", "Dz")]
    return dict(eps=eps, dL=cases.DL, npml=list(npml), steps=steps, sources=src, probes=probes)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("shape,npml,comp,expect", [
    ((48, 64, 1), (5, 6, 0), "z", 0b011100),      # TM: Ez, Hx, Hy
    ((48, 64, 1), (5, 6, 0), "x", 0b100011),      # TE: Ex, Ey, Hz
    ((48, 64, 1), (5, 6, 0), "y", 0b100011),
    ((33, 17, 1), (4, 3, 0), "z", 0b011100),      # odd contiguous extent: baseline kernels
    ((96, 1, 1), (8, 0, 0), "z", 0b010100),       # 1-D: Ez, Hy
    ((1, 24, 40), (0, 3, 4), "x", 0b110001),      # Nx = 1: Ex, Hy, Hz
    ((12, 10, 16), (2, 2, 3), "z", 0b111111),     # 3-D: everything couples
])
def test_component_specialisation_is_exact(shape, npml, comp, expect, dtype):
    import ceviche_b200
    case = _polarised_case(shape, npml, comp)
    out = []
    for spec in (False, True):
        F = ceviche_b200.fdtd(case["eps"], case["dL"], case["npml"], dtype=dtype)
        F.specialise_components = spec
        series = F.run(case["steps"], case["sources"], case["probes"]).cpu().numpy()
        if spec:
            assert F._active == expect and F._options.get("active_components", 63) == expect
        # a second leg through the same object: the mask must persist / widen correctly
        series2 = F.run(30, [(c, p, w[:30]) for c, p, w in case["sources"]], case["probes"]).cpu().numpy()
        out.append((series, series2, {k: F.fields[k].cpu().numpy() for k in FIELD_KEYS},
                    [t.cpu().numpy() for fam in ("ICE", "IH", "ICH", "ID") for t in F._pml[fam]]))
    (s0, t0, f0, p0), (s1, t1, f1, p1) = out
    assert np.array_equal(s0, s1) and np.array_equal(t0, t1)
    for k in FIELD_KEYS:
        assert np.array_equal(f0[k], f1[k]), k
    for a, b in zip(p0, p1):
        assert np.array_equal(a, b)
    if dtype == torch.float64:
        O = OracleFDTD(case["eps"], case["dL"], case["npml"])
        o_series, _ = O.run(case["steps"], case["sources"], case["probes"])
        for p in range(s1.shape[1]):
            assert rel_l2(s1[:, p], o_series[:, p]) <= 1e-10


def test_component_mask_widens_when_a_new_polarisation_is_driven():
    """TM run, then a TE source on the same object, then the per-step API: the skipped set must shrink."""
    import ceviche_b200
    shape, npml = (40, 48, 1), (4, 5, 0)
    rng = np.random.default_rng(8)
    eps = 1 + rng.random(shape)
    pz, px = cases.one_hot(shape, (20, 24, 0)), cases.one_hot(shape, (11, 30, 0))
    w = cases.gaussian(50, 20, 6)
    res = []
    for spec in (False, True):
        F = ceviche_b200.fdtd(eps, cases.DL, npml)
        F.specialise_components = spec
        F.run(50, [("z", pz, w)], [])
        a1 = F._active
        F.run(50, [("x", px, w)], [])
        a2 = F._active
        F.forward(Jy=torch.as_tensor(px).cuda() * 0.3)
        res.append({k: F.fields[k].cpu().numpy() for k in FIELD_KEYS})
        if spec:
            assert (a1, a2) == (0b011100, 0b111111)
    for k in FIELD_KEYS:
        assert np.array_equal(res[0][k], res[1][k]), k
    assert all(np.abs(res[1][k]).max() > 0 for k in FIELD_KEYS)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("shape,npml,comps", [
    ((20, 18, 136), (4, 3, 6), "xyz"),       # 3-D, all components
    ((12, 1, 72), (3, 0, 5), "y"),           # Ny = 1 (identity labelling, single-row planes)
    ((48, 64, 1), (5, 6, 0), "z"),           # 2-D TM: relabelled (x, z, y), compile-time mask
    ((48, 64, 1), (5, 6, 0), "x"),           # 2-D TE
    ((48, 64, 1), (5, 6, 0), "xz"),          # 2-D, both polarisations (all six components)
    ((33, 17, 1), (4, 3, 0), "z"),           # odd contiguous extent: baseline kernels either way
])
def test_tangent_marching_equals_baseline_bitwise(shape, npml, comps, dtype):
    """Forward mode: the marching tangent H kernel (E = mE*dD + dmE*D) against the one-thread-per-cell one."""
    import ceviche_b200
    rng = np.random.default_rng(12)
    steps = 40
    eps = 1 + 3 * rng.random(shape)
    mid = tuple(n // 2 for n in shape)
    src = [(c, cases.one_hot(shape, mid, 2.0) + rng.random(shape) * (rng.random(shape) < 0.03),
            cases.modulated(steps, steps / 3, steps / 8, 9.0, 2.0)) for c in comps]
    probes = [(k, rng.random(shape)) for k in ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz")]
    V = torch.as_tensor(rng.standard_normal((3,) + shape))
    out = []
    for variant, streams in ((1, 0), (0, 0), (0, 1)):      # streams: the tangent states on side streams (fork / join)
        F = ceviche_b200.fdtd(eps, cases.DL, list(npml), dtype=dtype)
        F.set_option("kernel_variant", variant)
        F.set_option("jvp_streams", streams)
        s, ds = F.jvp_run(steps, V, src, probes)
        tangents = [[t.cpu().numpy() for t in tH + tD] for _, tH, tD, _ in F._tangent_states]
        out.append((s.cpu().numpy(), ds.cpu().numpy(), tangents, {k: F.fields[k].cpu().numpy() for k in FIELD_KEYS}))
    s1, d1, t1, f1 = out[0]
    assert np.abs(d1).max() > 0
    for s0, d0, t0, f0 in out[1:]:
        assert np.array_equal(s1, s0) and np.array_equal(d1, d0)
        for a, b in zip(t1, t0):
            for x, y in zip(a, b):
                assert np.array_equal(x, y)
        for k in FIELD_KEYS:
            assert np.array_equal(f1[k], f0[k]), k


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("shape,npml,comps", [((24, 20, 40), (4, 3, 5), "xyz"), ((60, 48, 1), (6, 5, 0), "z"),
                                              ((31, 17, 1), (5, 4, 0), "x")])
def test_graph_replay_equals_plain_launches(shape, npml, comps, dtype):
    """CUDA-graph replay of the caller loop (blocks of 50 steps through staging buffers) against the plain
    launch loop: series (both probe families), fields and PML integrals, over several run() legs on one object
    (graph cache hits, new state pointers, changed sources)."""
    import ceviche_b200
    rng = np.random.default_rng(3)
    steps = 173
    eps = 1 + 3 * rng.random(shape)
    mid = tuple(n // 2 for n in shape)
    src = [(c, cases.one_hot(shape, mid, 2.0) + rng.random(shape) * (rng.random(shape) < 0.03),
            cases.modulated(steps, steps / 3, steps / 8, 9.0, 2.0)) for c in comps]
    probes = [(k, rng.random(shape)) for k in ("Ez", "Hx", "Dx", "Hy", "Ey", "Hz")]
    res = []
    for graph in (0, 1):
        F = ceviche_b200.fdtd(eps, cases.DL, list(npml), dtype=dtype)
        F.set_option("use_graph", graph)
        legs = [F.run(steps, src, probes).cpu().numpy()]
        F.prepare(src, probes)
        wf = torch.as_tensor(np.stack([w for _, _, w in src], 1)).cuda()
        legs.append(F.run(steps, waveforms=wf).cpu().numpy())          # prepared sources: no epoch change
        legs.append(F.run(120, waveforms=wf[:120]).cpu().numpy())
        legs.append(F.run(101, [(c, p * 0.5, w[:101]) for c, p, w in src], probes[:3]).cpu().numpy())   # new tables
        legs.append(F.run(60, waveforms=wf[:60]).cpu().numpy())         # too short for a graph
        res.append((legs, {k: F.fields[k].cpu().numpy() for k in FIELD_KEYS},
                    [t.cpu().numpy() for fam in ("ICE", "IH", "ICH", "ID") for t in F._pml[fam]]))
    (l0, f0, p0), (l1, f1, p1) = res
    for a, b in zip(l0, l1):
        assert np.array_equal(a, b)
    assert np.abs(l1[0]).max() > 0
    for k in FIELD_KEYS:
        assert np.array_equal(f0[k], f1[k]), k
    for a, b in zip(p0, p1):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("dtype,arith", [(torch.float64, None), (torch.float32, None), (torch.float32, "f64")])
@pytest.mark.parametrize("shape,npml", [((96, 128, 1), (5, 6, 0)), ((70, 264, 1), (8, 9, 0)), ((33, 40, 1), (0, 0, 0))])
def test_fused_tangent_step_equals_the_two_kernel_path(shape, npml, dtype, arith):
    """2-D TM forward-mode sweeps: the fused tangent step (csrc/tan2d_fused.cuh: both half-steps of all B tangent states in
    one launch, states ping-ponged) against the batched half-step launches: probe series, tangent series and the final
    tangent states (fields and PML integrals) bit for bit, for odd and even step counts."""
    import ceviche_b200
    for steps in (1, 2, 7, 30):
        rng = np.random.default_rng(5)
        eps = 1 + 2 * rng.random(shape)
        V = torch.as_tensor(rng.standard_normal((3,) + shape))
        prof = np.zeros(shape); prof[shape[0] // 2, shape[1] // 3, 0] = 1.0
        prof2 = rng.random(shape) * (rng.random(shape) < 0.02)
        probes = [("Ez", rng.random(shape)), ("Hx", rng.random(shape)), ("Hy", (rng.random(shape) < 0.1) * 1.0), ("Dz", rng.random(shape))]
        t = np.arange(steps)
        srcs = [("z", prof, np.cos(0.3 * t) + 1), ("z", prof2, np.sin(0.2 * t + 0.3))]
        out = {}
        for fused in (0, 1):
            F = ceviche_b200.fdtd(eps, 5e-8, list(npml), dtype=dtype, arith=arith)
            F.set_option("jvp_fused", fused)
            s_, ds = F.jvp_run(steps, V, srcs, probes)
            out[fused] = (s_, ds, [x.clone() for st in F._tangent_states for grp in st[1:] for x in grp])
        assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1]), steps
        assert all(torch.equal(a, b) for a, b in zip(out[0][2], out[1][2])), steps

