"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (TEST INFRASTRUCTURE).

Run in the build container (needs /root/reference):   python -m oracle.make_golden

Field cases: the reference's own ``ceviche/fdtd.py`` (loaded unmodified by
oracle/ref_loader.py) is stepped in the reference's caller loop
(ceviche/utils.py:325-331) on the seeded inputs of oracle/cases.py; we keep the probe
series, every field at the snapshot steps (sub-sampled for the two 200x200 cases so the
fixtures stay small) and the full-grid L2 norm of every field.

Gradient cases: d(objective)/d(eps_r) by (i) torch.autograd / torch.func.jvp over the
bit-identical torch restatement and (ii) finite differences through the reference numpy
code -- one-sided with step 1e-6 exactly as tests/test_gradients_fdtd.py:19-20 does, and
central.  HIPS autograd is not installed here, so (i)+(ii) stand in for the reference's AD.
"""
import os
import sys

import numpy as np
import torch

from . import cases, ref_loader
from .fdtd_numpy import FIELD_KEYS, pad_to_3d
from .fdtd_torch import series_fn

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def run_reference(case, eps=None):
    """The reference's caller loop; returns (series, {t: fields})."""
    ref = ref_loader.load()
    F = ref.fdtd(case["eps"] if eps is None else eps, case["dL"], case["npml"])
    steps = case["steps"]
    series = np.zeros((steps, len(case["probes"])))
    snaps = {}
    for t in range(steps):
        J = {"x": None, "y": None, "z": None}
        for comp, profile, wave in case["sources"]:
            term = pad_to_3d(profile) * wave[t]
            J[comp] = term if J[comp] is None else J[comp] + term
        f = F.forward(Jx=J["x"], Jy=J["y"], Jz=J["z"])
        for p, (key, mask) in enumerate(case["probes"]):
            series[t, p] = np.sum(f[key] * pad_to_3d(mask))
        if (t + 1) in case.get("snapshots", ()):
            snaps[t + 1] = {k: np.array(f[k], copy=True) for k in FIELD_KEYS}
    return series, snaps, F


def make_field_golden(name):
    case = cases.field_case(name)
    series, snaps, F = run_reference(case)
    out = {"series": series, "dt": np.float64(F.dt)}
    stride = 5 if name.startswith("c1_") else 1
    out["stride"] = np.int64(stride)
    for t, fields in snaps.items():
        for k, v in fields.items():
            out["t%d_%s" % (t, k)] = v[::stride, ::stride, :]
            out["t%d_%s_norm" % (t, k)] = np.float64(np.linalg.norm(v))
    np.savez_compressed(os.path.join(OUT, "fields_%s.npz" % name), **out)
    print("fields", name, "series norm", np.linalg.norm(series))


def _scalar_objective(case):
    w = cases.objective_weights(case["steps"], len(case["probes"]))
    if case.get("ref_style", False):
        return lambda series: series.sum()
    return lambda series: (series ** 2 * (torch.as_tensor(w) if torch.is_tensor(series) else w)).sum()


def make_grad_golden(name):
    case = cases.grad_case(name)
    shape = case["eps"].shape
    fn = series_fn(shape, case["dL"], case["npml"], case["steps"], case["sources"], case["probes"])
    eps0 = torch.as_tensor(case["eps"].copy())
    out = {}
    if name.startswith("ref_"):
        case["ref_style"] = True
    if case["mode"] == "rev":
        obj = _scalar_objective(case)
        x = eps0.clone().requires_grad_(True)
        L = obj(fn(x))
        (g,) = torch.autograd.grad(L, x)
        out["value"] = np.float64(L.item())
        out["grad_ad"] = g.numpy()
        ref_L = lambda e: obj(run_reference(case, eps=e)[0])
        base = ref_L(case["eps"])
        out["value_ref"] = np.float64(base)
        if name.startswith("ref_"):
            cells = list(np.ndindex(shape))
        else:
            rng = np.random.default_rng(3)
            cells = [tuple(int(rng.integers(0, n)) for n in shape) for _ in range(6)]
        h = 1e-6
        fd1, fdc = [], []
        for idx in cells:
            e = case["eps"].copy(); e[idx] += h
            up = ref_L(e)
            e = case["eps"].copy(); e[idx] -= h
            dn = ref_L(e)
            fd1.append((up - base) / h)
            fdc.append((up - dn) / (2 * h))
        out["fd_cells"] = np.array(cells)
        out["fd_one_sided"] = np.array(fd1)
        out["fd_central"] = np.array(fdc)
    else:
        # tests/test_gradients_fdtd.py:91-116: objective(c) = sum_t (F_x+F_y+F_z) with eps = c*eps0, c0 = 2
        c0 = 2.0
        arr_obj = lambda series_like: series_like
        def of_c_torch(c):
            sim_fn = _fields_sum_fn(case)
            return sim_fn(c * eps0)
        c = torch.tensor(c0, dtype=torch.float64)
        val, tan = torch.func.jvp(of_c_torch, (c,), (torch.ones_like(c),))
        out["value"] = val.numpy()
        out["jvp_ad"] = tan.numpy()
        ref_S = lambda cc: _fields_sum_reference(case, cc * case["eps"])
        h = 1e-6
        base = ref_S(c0)
        out["value_ref"] = base
        out["fd_one_sided"] = (ref_S(c0 + h) - base) / h
        out["fd_central"] = (ref_S(c0 + h) - ref_S(c0 - h)) / (2 * h)
    np.savez_compressed(os.path.join(OUT, "grad_%s.npz" % name), **out)
    print("grad", name, {k: (np.linalg.norm(v) if np.ndim(v) else float(v)) for k, v in out.items() if k != "fd_cells"})


def _fields_sum_fn(case):
    """eps (torch) -> sum_t (Fx+Fy+Fz) as an array, the forward-mode objective of the reference tests."""
    from .fdtd_torch import TorchFDTD
    keys = [k for k, _ in case["probes"]]
    def fn(eps):
        sim = TorchFDTD(eps, case["dL"], case["npml"])
        S = 0.0
        for t in range(case["steps"]):
            J = {"x": None, "y": None, "z": None}
            for comp, profile, wave in case["sources"]:
                J[comp] = torch.as_tensor(profile) * float(wave[t])
            f = sim.step(Jx=J["x"], Jy=J["y"], Jz=J["z"])
            S = S + f[keys[0]] + f[keys[1]] + f[keys[2]]
        return S
    return fn


def _fields_sum_reference(case, eps):
    ref = ref_loader.load()
    F = ref.fdtd(eps, case["dL"], case["npml"])
    keys = [k for k, _ in case["probes"]]
    S = 0.0
    for t in range(case["steps"]):
        J = {"x": None, "y": None, "z": None}
        for comp, profile, wave in case["sources"]:
            J[comp] = profile * wave[t]
        f = F.forward(Jx=J["x"], Jy=J["y"], Jz=J["z"])
        S = S + f[keys[0]] + f[keys[1]] + f[keys[2]]
    return S


def make_modes_golden():
    """Modes of the reference's own self-test ridge (ceviche/modes.py:140-165) from the REFERENCE's get_modes."""
    modes = ref_loader.load_modes()
    lambda0 = 1.550e-6
    dL = lambda0 / 100
    omega = 2 * np.pi * modes.C_0 / lambda0
    Nx = int(lambda0 * 10 / dL)
    eps = np.ones((Nx,))
    w = int(lambda0 / dL / 2)
    eps[Nx // 2 - w:Nx // 2 + w] = 4.0
    vals, vecs = modes.get_modes(eps, omega, dL, 10, m=6)
    np.savez_compressed(os.path.join(OUT, "modes_ridge.npz"), vals=vals, vecs=vecs, eps=eps, omega=omega, dL=dL)
    print("modes_ridge", vals)


def make_c2_scaled_golden():
    """Config 2 at 96^3 for 2000 steps: the first 150 steps by the REFERENCE ITSELF, all 2000 by the C restatement
    (oracle/fdtd_c.c), asserted bit-identical on the overlap (and in tests/test_oracle_c.py in general)."""
    from .fdtd_c import OracleFDTDC
    case = cases.scaled_case("c2_96")
    head = dict(case, steps=150, sources=[(c, p, w[:150]) for c, p, w in case["sources"]], snapshots=())
    ref_series, _, F = run_reference(head)
    O = OracleFDTDC(case["eps"], case["dL"], case["npml"])
    series, snaps = O.run(case["steps"], case["sources"], case["probes"], case["snapshots"])
    assert np.array_equal(series[:150], ref_series), "C restatement differs from the reference"
    out = {"series": series, "dt": np.float64(F.dt), "stride": np.int64(6), "reference_steps": np.int64(150)}
    for k, v in snaps[case["steps"]].items():
        out["end_" + k] = v[::6, ::6, ::6]
        out["end_%s_norm" % k] = np.float64(np.linalg.norm(v))
    np.savez_compressed(os.path.join(OUT, "fields_c2_96.npz"), **out)
    print("fields c2_96 series norm", np.linalg.norm(series, axis=0))


def make_c5_scaled_golden():
    """Config 5 at 256x256: d(series)/d(eps_r) along four directions by torch.func.jvp over the torch restatement, and
    by a central finite difference through the REFERENCE's numpy code for the first direction."""
    case = cases.scaled_case("c5_small")
    shape = case["eps"].shape
    fn = series_fn(shape, case["dL"], case["npml"], case["steps"], case["sources"], case["probes"])
    eps0 = torch.as_tensor(case["eps"].copy())
    series, ds = None, []
    for b in range(case["directions"].shape[0]):
        series, tan = torch.func.jvp(fn, (eps0,), (torch.as_tensor(case["directions"][b]),))
        ds.append(tan.numpy())
    h = 1e-4
    up = run_reference(case, eps=case["eps"] + h * case["directions"][0])[0]
    dn = run_reference(case, eps=case["eps"] - h * case["directions"][0])[0]
    np.savez_compressed(os.path.join(OUT, "jvp_c5_small.npz"), series=series.numpy(), dseries=np.stack(ds),
                        fd_central_dir0=(up - dn) / (2 * h))
    print("jvp c5_small", np.linalg.norm(np.stack(ds), axis=(1, 2)), "fd/ad dir0",
          np.linalg.norm((up - dn) / (2 * h) - ds[0]) / np.linalg.norm(ds[0]))


SPLITTER_EXAMPLE = dict(Nx=320, Ny=180, steps=4500, t0=600, sigma=100, npml=20, dL=5e-8)


def make_splitter_example_golden():
    """examples/simulate_splitter_fdtd.py at half size: the REFERENCE's own `measure_fields` (ceviche/utils.py:316-332)
    over the REFERENCE's own fdtd object, called as the notebook calls it (`measure_fields(F, source, steps, J_outs)` with
    source = t -> J_in * amp * gaussian(t)), on the example's geometry, for the straight guide and for the splitter."""
    import contextlib
    import importlib.util
    import io
    root = os.path.dirname(OUT.rstrip(os.sep).rsplit(os.sep, 1)[0])
    spec = importlib.util.spec_from_file_location("simulate_splitter_fdtd", os.path.join(root, "examples", "simulate_splitter_fdtd.py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    ref = ref_loader.load()
    ref_utils = sys.modules["ceviche.utils"]
    c = SPLITTER_EXAMPLE
    eps_wg, eps_r, J_in, J_wg, J_outs = ex.geometry(c["Nx"], c["Ny"], c["npml"])
    F = ref.fdtd(eps_r, dL=c["dL"], npml=[c["npml"], c["npml"], 0])
    F_wg = ref.fdtd(eps_wg, dL=c["dL"], npml=[c["npml"], c["npml"], 0])
    wave = ex.pulse(c["steps"], F.dt, c["t0"], c["sigma"])
    source = lambda t: J_in * wave[t]
    with contextlib.redirect_stdout(io.StringIO()):          # (the reference prints a progress line every 5 %)
        measured_wg = ref_utils.measure_fields(F_wg, source, c["steps"], J_wg)
        measured = ref_utils.measure_fields(F, source, c["steps"], J_outs)
    T, f_max = ex.transmission(measured, measured_wg, F.dt)
    np.savez_compressed(os.path.join(OUT, "example_splitter.npz"), measured_wg=measured_wg, measured=measured, T=T,
                        f_max=np.float64(f_max), dt=np.float64(F.dt), **{k: np.float64(v) for k, v in c.items()})
    print("example_splitter T", T, "f_max", f_max, "peak", np.abs(measured_wg).max(), np.abs(measured).max(0))


def main(argv):
    os.makedirs(OUT, exist_ok=True)
    which = argv[1:] or ["fields", "grads", "modes", "scaled", "splitter"]
    if "splitter" in which:
        make_splitter_example_golden()
    if "scaled" in which:
        make_c2_scaled_golden()
        make_c5_scaled_golden()
    if "modes" in which:
        make_modes_golden()
    if "fields" in which:
        for name in cases.FIELD_CASES:
            make_field_golden(name)
    if "grads" in which:
        for name in cases.GRAD_CASES:
            make_grad_golden(name)


if __name__ == "__main__":
    main(sys.argv)
