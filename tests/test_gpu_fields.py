"""Parity of the CUDA path (through the Python boundary -> C ABI) against the numpy oracle and
the committed golden vectors.  Tolerances: rel-L2 <= 1e-10 (fp64), <= 1e-5 (fp32), as north_star."""
import os

import numpy as np
import pytest
import torch

from oracle import cases
from oracle.fdtd_numpy import FIELD_KEYS, OracleFDTD, pad_to_3d, rel_l2

pytestmark = pytest.mark.gpu

TOL = {torch.float64: 1e-10, torch.float32: 1e-5}


def _run_cuda(case, dtype, fused, arith=None):
    import ceviche_b200
    F = ceviche_b200.fdtd(case["eps"], case["dL"], case["npml"], dtype=dtype, arith=arith)
    steps = case["steps"]
    snaps = {}
    if fused:
        series = []
        t0 = 0
        for t1 in sorted(set(case["snapshots"]) | {steps}):
            srcs = [(c, p, w[t0:t1]) for c, p, w in case["sources"]]
            series.append(F.run(t1 - t0, srcs, case["probes"]))
            if t1 in case["snapshots"]:
                snaps[t1] = {k: F.fields[k].cpu().numpy() for k in FIELD_KEYS}
            t0 = t1
        series = torch.cat(series).cpu().numpy()
    else:
        series = np.zeros((steps, len(case["probes"])))
        masks = [torch.as_tensor(pad_to_3d(m)).cuda() for _, m in case["probes"]]
        profs = [(c, torch.as_tensor(pad_to_3d(p)).cuda()) for c, p, _ in case["sources"]]
        for t in range(steps):
            J = {"x": None, "y": None, "z": None}
            for (c, p), (_, _, w) in zip(profs, case["sources"]):
                term = p * float(w[t])
                J[c] = term if J[c] is None else J[c] + term
            f = F.forward(Jx=J["x"], Jy=J["y"], Jz=J["z"])
            for q, (key, _) in enumerate(case["probes"]):
                series[t, q] = float(torch.sum(f[key].double() * masks[q]))
            if (t + 1) in case["snapshots"]:
                snaps[t + 1] = {k: f[k].cpu().numpy() for k in FIELD_KEYS}
    return F, series, snaps


def _check(case, series, snaps, o_series, o_snaps, tol):
    worst = 0.0
    for t in case["snapshots"]:
        for k in FIELD_KEYS:
            e = rel_l2(snaps[t][k], o_snaps[t][k])
            assert e <= tol, (t, k, e)
            worst = max(worst, e)
        allk = np.concatenate([snaps[t][k].ravel() for k in FIELD_KEYS])
        allo = np.concatenate([o_snaps[t][k].ravel() for k in FIELD_KEYS])
        assert rel_l2(allk, allo) <= tol
    for p in range(series.shape[1]):
        e = rel_l2(series[:, p], o_series[:, p])
        assert e <= tol, ("series", p, e)
    return worst


@pytest.mark.parametrize("fused", [True, False], ids=["run", "forward"])
@pytest.mark.parametrize("name", cases.SMALL_FIELD_CASES)
def test_small_cases_fp64(name, fused):
    case = cases.field_case(name)
    O = OracleFDTD(case["eps"], case["dL"], case["npml"])
    o_series, o_snaps = O.run(case["steps"], case["sources"], case["probes"], case["snapshots"])
    F, series, snaps = _run_cuda(case, torch.float64, fused)
    assert F.dt == O.dt
    _check(case, series, snaps, o_series, o_snaps, TOL[torch.float64])


@pytest.mark.parametrize("arith", ["f32", "f64"])
@pytest.mark.parametrize("name", ["pml3d", "odd2d", "tall_z"])
def test_small_cases_fp32(name, arith):
    case = cases.field_case(name)
    O = OracleFDTD(case["eps"], case["dL"], case["npml"])
    o_series, o_snaps = O.run(case["steps"], case["sources"], case["probes"], case["snapshots"])
    _, series, snaps = _run_cuda(case, torch.float32, True, arith=arith)
    _check(case, series, snaps, o_series, o_snaps, TOL[torch.float32])


@pytest.mark.parametrize("name", ["c1_tm", "c1_te"])
def test_config1_against_golden(name, golden_dir):
    """BASELINE config 1 (2-D 200x200, npml 20, dipole, 1000 steps, fp64) against vectors produced
    by the reference itself; zero fields must stay exactly zero."""
    case = cases.field_case(name)
    gold = np.load(os.path.join(golden_dir, "fields_%s.npz" % name))
    F, series, snaps = _run_cuda(case, torch.float64, True)
    assert F.dt == float(gold["dt"])
    s = int(gold["stride"])
    for p in range(series.shape[1]):
        assert rel_l2(series[:, p], gold["series"][:, p]) <= 1e-10
    for t in case["snapshots"]:
        for k in FIELD_KEYS:
            g = gold["t%d_%s" % (t, k)]
            assert rel_l2(snaps[t][k][::s, ::s, :], g) <= 1e-10, (t, k)
            gn = float(gold["t%d_%s_norm" % (t, k)])
            n = float(np.linalg.norm(snaps[t][k]))
            assert (n == 0.0) if gn == 0.0 else abs(n - gn) <= 1e-10 * gn, (t, k)


def test_config1_fp32_within_tolerance(golden_dir):
    case = cases.field_case("c1_tm")
    gold = np.load(os.path.join(golden_dir, "fields_c1_tm.npz"))
    for arith in ("f32", "f64"):
        _, series, snaps = _run_cuda(case, torch.float32, True, arith=arith)
        s = int(gold["stride"])
        for k in ("Ez", "Hx", "Hy"):
            e = rel_l2(snaps[1000][k][::s, ::s, :], gold["t1000_%s" % k])
            assert e <= 1e-5, (arith, k, e)
        assert rel_l2(series[:, 0], gold["series"][:, 0]) <= 1e-5


def test_forward_and_run_agree_bitwise():
    case = cases.field_case("pml3d")
    _, s_run, f_run = _run_cuda(case, torch.float64, True)
    _, s_fwd, f_fwd = _run_cuda(case, torch.float64, False)
    t = case["steps"]
    for k in FIELD_KEYS:
        assert np.array_equal(f_run[t][k], f_fwd[t][k]), k


def test_reference_api_quirks():
    import ceviche_b200
    eps = 1 + np.random.default_rng(3).random((10, 9))
    F = ceviche_b200.fdtd(eps, 5e-8, [2, 2, 0])
    assert F.grid_shape == (10, 9, 1) and F.N == 90 and F.t_index == 0
    d0 = F.fields
    f1 = F.forward(Jz=2.0)                      # scalar J adds to every cell
    assert f1 is d0 and F.t_index == 1
    assert torch.all(f1["Dz"] == 2.0)
    held = f1["Ez"]
    held_copy = held.clone()
    F.forward(Jz=np.ones((10, 9, 1)))
    assert torch.equal(held, held_copy)         # previously returned arrays stay valid
    assert F.fields["Ez"] is not held
    F.run(3)
    assert torch.equal(held, held_copy) and F.t_index == 5
    F.eps_r = torch.as_tensor(eps.reshape(10, 9, 1) * 2)   # setter resets fields and t_index
    assert F.t_index == 0 and float(F.fields["Ez"].abs().max()) == 0.0
    assert F.fields is not d0
    with pytest.raises(ValueError):
        F.eps_r = eps                            # 2-D through the setter is an error in the reference too
    with pytest.raises(ValueError):
        ceviche_b200.fdtd(np.ones((2, 2, 2, 2)), 5e-8, [0, 0, 0])
    with pytest.raises(TypeError):               # complex J: the reference's in-place `D += J` raises numpy's UFuncTypeError
        F.forward(Jz=np.ones((10, 9, 1)) * (1 + 2j))     # (a TypeError; tests/test_oracle_vs_reference.py shows it)
    O = OracleFDTD(eps, 5e-8, [2, 2, 0])
    assert np.array_equal(F.eps_xx.cpu().numpy() / 2, O.eps_yee[0])
    assert repr(F) == "FDTD(eps_r.shape=(10, 9, 1), dL=5e-08, NPML=[2, 2, 0])"


def test_edge_cases_empty_inputs():
    """Zero steps, probes / sources with no points, no probes at all, a run continued after forward()."""
    import ceviche_b200
    shape = (9, 8, 16)
    eps = 1 + np.random.default_rng(4).random(shape)
    F = ceviche_b200.fdtd(eps, 5e-8, [2, 0, 3])
    zero = np.zeros(shape)
    hot = np.zeros(shape); hot[4, 4, 8] = 1.0
    s = F.run(0, [("z", hot, np.zeros(0))], [("Ez", hot)])
    assert tuple(s.shape) == (0, 1) and F.t_index == 0
    s = F.run(5, [("z", zero, np.ones(5)), ("x", hot, np.ones(5))], [("Ez", zero), ("Hy", hot)])
    assert tuple(s.shape) == (5, 2) and float(s[:, 0].abs().max()) == 0.0 and float(s[:, 1].abs().max()) > 0.0
    s = F.run(3, sources=(), probes=())            # no sources, no probes: free evolution
    assert tuple(s.shape) == (3, 0) and F.t_index == 8
    O = OracleFDTD(eps, 5e-8, [2, 0, 3])
    for t in range(5):
        O.step(Jx=hot * 1.0)
    for t in range(3):
        O.step()
    f = F.forward(Jy=hot)                          # per-step call continues the fused run's state
    g = O.step(Jy=hot)
    for k in FIELD_KEYS:
        assert rel_l2(f[k].cpu().numpy(), g[k]) <= 1e-12, k


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("shape,npml", [((18, 16, 24), (3, 3, 4)), ((40, 36, 1), (5, 4, 0))])
def test_running_dft_monitors_equal_dft_of_the_time_series(shape, npml, dtype):
    """Running-DFT monitors against (i) the DFT of the numpy oracle's per-point time series and (ii) np.fft.fft
    bins (the convention of ceviche/utils.py:383-386), over two run() legs."""
    import ceviche_b200
    from oracle import cases
    from oracle.fdtd_numpy import OracleFDTD, rel_l2
    rng = np.random.default_rng(2)
    steps = 96
    eps = 1 + 2 * rng.random(shape)
    mid = tuple(n // 2 for n in shape)
    src = [("z", cases.one_hot(shape, mid, 3.0), cases.modulated(steps, 30, 10, 8.0, 2.0))]
    # two monitors: a few scattered Ez points and a small Hy patch
    pts_e = [tuple(int(rng.integers(0, n)) for n in shape) for _ in range(5)] + [mid]
    m_e = np.zeros(shape); m_h = np.zeros(shape)
    for q in pts_e:
        m_e[q] = 1.0
    m_h[mid[0] - 2:mid[0] + 2, mid[1] - 1:mid[1] + 2, mid[2]] = 1.0
    F = ceviche_b200.fdtd(eps, cases.DL, list(npml), dtype=dtype)
    dt = F.dt
    freqs = np.array([3, 7, 12]) / (steps * dt)          # FFT bins 3, 7, 12 of a 96-sample series
    F.run(40, [(c, p, w[:40]) for c, p, w in src], [], monitors=[("Ez", m_e), ("Hy", m_h)], freqs=freqs)
    F.run(steps - 40, [(c, p, w[40:]) for c, p, w in src], [])
    vals = [v.cpu().numpy() for v in F.monitor_values()]
    idx = [p.cpu().numpy() for p in F.monitor_points]
    assert vals[0].shape == (3, len(set(pts_e))) and vals[1].shape == (3, int(m_h.sum()))
    # oracle: every monitored point as a one-hot probe
    probes = [("Ez", cases.one_hot(shape, np.unravel_index(q, shape))) for q in idx[0]] + \
             [("Hy", cases.one_hot(shape, np.unravel_index(q, shape))) for q in idx[1]]
    O = OracleFDTD(eps, cases.DL, list(npml))
    series, _ = O.run(steps, src, probes)
    n = np.arange(steps)
    dft = np.exp(-2j * np.pi * freqs[:, None] * n[None, :] * dt) @ series          # (freq, point)
    fft = np.fft.fft(series, axis=0)[[3, 7, 12]]
    got = np.concatenate(vals, axis=1)
    tol = 1e-10 if dtype == torch.float64 else 1e-5
    assert rel_l2(got.real, dft.real) <= tol and rel_l2(got.imag, dft.imag) <= tol
    assert rel_l2(got.real, fft.real) <= tol * 10 and rel_l2(got.imag, fft.imag) <= tol * 10
    F.initialize_fields()
    assert all(float(v.abs().max()) == 0.0 for v in F.monitor_values())


def test_measure_fields_and_aniplot_against_the_oracle_loop(capsys):
    """ceviche/utils.py:316-332 (measure_fields) and :279-313 (aniplot) on the drop-in object: both source forms of
    measure_fields -- the reference's callable t -> J array (one forward() per step) and (profile, waveform) (the fused
    device loop) -- against the oracle's caller loop; aniplot's panels against the oracle's snapshots."""
    import ceviche_b200
    from ceviche_b200.utils import aniplot, measure_fields
    from oracle.fdtd_numpy import OracleFDTD
    shape, npml, steps = (40, 30, 1), [6, 5, 0], 60
    rng = np.random.default_rng(3)
    eps = 1 + 2 * rng.random(shape)
    prof = cases.one_hot(shape, (20, 15, 0), 3.0)
    wave = cases.gaussian(steps, 20, 6, 2.0)
    probes = [cases.one_hot(shape, (25, 12, 0)), rng.random(shape)]
    for comp in ("Ez", "Hy"):
        O = OracleFDTD(eps, cases.DL, npml)
        want, _ = O.run(steps, [("z", prof, wave)], [(comp, p) for p in probes])
        F = ceviche_b200.fdtd(eps, cases.DL, npml)
        got_loop = measure_fields(F, lambda t: prof * wave[t], steps, probes, component=comp, verbose=True)
        assert "% done" in capsys.readouterr().out                     # the reference's progress lines
        got_run = measure_fields(F, (prof, wave), steps, probes, component=comp)
        assert got_loop.shape == got_run.shape == (steps, 2)
        for p in range(2):
            assert rel_l2(got_loop[:, p], want[:, p]) <= 1e-10 and rel_l2(got_run[:, p], want[:, p]) <= 1e-10
        single = measure_fields(F, (prof, wave), steps, probes[0], component=comp)     # a bare probe, not a list
        assert np.array_equal(single[:, 0], got_run[:, 0])
    O = OracleFDTD(eps, cases.DL, npml)
    _, snaps = O.run(steps, [("z", prof, wave)], [], snapshots=tuple(range(1, steps + 1)))
    panels = aniplot(ceviche_b200.fdtd(eps, cases.DL, npml), lambda t: prof * wave[t], steps, num_panels=5, show=False)
    assert [t for t, _ in panels] == [0, 12, 24, 36, 48]
    for t, arr in panels:
        assert rel_l2(arr, snaps[t + 1]["Ez"][:, :, 0]) <= 1e-10


def test_splitter_example_against_the_reference(golden_dir):
    """examples/simulate_splitter_fdtd.py (the time-domain workflow of the reference's simulate_splitter_fdtd.ipynb) at
    half size: the probe series of the straight guide and of the two splitter arms against what the REFERENCE's own
    measure_fields returned on the same inputs (tests/golden/example_splitter.npz, oracle/make_golden.py splitter), and
    the derived transmission / peak frequency."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("simulate_splitter_fdtd", os.path.join(root, "examples", "simulate_splitter_fdtd.py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    gold = np.load(os.path.join(golden_dir, "example_splitter.npz"))
    kw = dict(Nx=int(gold["Nx"]), Ny=int(gold["Ny"]), steps=int(gold["steps"]), t0=int(gold["t0"]), sigma=int(gold["sigma"]),
              npml=int(gold["npml"]), dL=float(gold["dL"]))
    R = ex.simulate(**kw)
    assert R["F"].dt == float(gold["dt"])
    assert rel_l2(R["measured_wg"][:, 0], gold["measured_wg"][:, 0]) <= 1e-10
    for arm in range(2):
        assert rel_l2(R["measured"][:, arm], gold["measured"][:, arm]) <= 1e-10, arm
    assert np.allclose(R["T"], gold["T"], rtol=1e-9, atol=0) and R["f_max"] == float(gold["f_max"])
    assert 0.9 < R["T"].sum() < 1.0
    R32 = ex.simulate(dtype=torch.float32, **kw)
    assert rel_l2(R32["measured_wg"][:, 0], gold["measured_wg"][:, 0]) <= 1e-5
    for arm in range(2):
        assert rel_l2(R32["measured"][:, arm], gold["measured"][:, arm]) <= 1e-5, arm
    assert np.allclose(R32["T"], gold["T"], rtol=1e-3, atol=0)
    # the reference's own call form (a callable source, one forward() per step) on the first 700 steps
    from ceviche_b200.utils import measure_fields
    early = measure_fields(R["F_wg"], lambda t: R["J_in"] * R["wave"][t], 700, R["J_in"])
    again = measure_fields(R["F_wg"], (R["J_in"], R["wave"][:700]), 700, R["J_in"])
    assert np.abs(again).max() > 0 and rel_l2(early[:, 0], again[:, 0]) <= 1e-12
