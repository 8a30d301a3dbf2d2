"""Seeded parity cases shared by oracle/make_golden.py and tests/ (TEST INFRASTRUCTURE).

Every case is a plain dict: eps, dL, npml, steps, sources [(comp, profile, waveform)],
probes [(field key, mask)], snapshots (1-based step counts at which all nine fields are compared).
Inputs mirror the reference's own tests: point dipole with the amplitude applied twice and a
Gaussian pulse (tests/test_fields_fdtd.py:16-34, tests/test_gradients_fdtd.py:31-50), dL = 5e-8.
"""
import numpy as np

DL = 5e-8


def gaussian(steps, t0, sigma, amp=1.0):
    t = np.arange(steps)
    return amp * np.exp(-(t - t0) ** 2 / 2 / sigma ** 2)


def modulated(steps, t0, sigma, period, amp=1.0):
    t = np.arange(steps)
    return amp * np.exp(-(t - t0) ** 2 / 2 / sigma ** 2) * np.cos(2 * np.pi * t / period)


def one_hot(shape, idx, val=1.0):
    a = np.zeros(shape)
    a[idx] = val
    return a


def _c1(component):
    """BASELINE config 1: 2-D 200x200, npml 20, point dipole, 1000 steps (tests/test_fields_fdtd.py)."""
    shape = (200, 200, 1)
    eps = 1 + np.random.default_rng(0).random(shape)
    steps = 1000
    prof = one_hot(shape, (100, 100, 0), 10.0)
    if component == "z":
        probes = [("Ez", one_hot(shape, (100, 100, 0))), ("Ez", one_hot(shape, (130, 95, 0))),
                  ("Hx", one_hot(shape, (100, 120, 0))), ("Hy", one_hot(shape, (60, 100, 0)))]
    else:
        probes = [("Ex", one_hot(shape, (100, 100, 0))), ("Ey", one_hot(shape, (130, 95, 0))),
                  ("Hz", one_hot(shape, (100, 120, 0))), ("Dx", one_hot(shape, (60, 100, 0)))]
    return dict(eps=eps, dL=DL, npml=[20, 20, 0], steps=steps,
                sources=[(component, prof, gaussian(steps, 300, 20, 10.0))],
                probes=probes, snapshots=(250, 500, 1000))


def _small3d(shape, npml, steps, seed, snapshots=None):
    rng = np.random.default_rng(seed)
    eps = 1 + 3 * rng.random(shape)
    sheet = np.zeros(shape)
    sheet[shape[0] // 3, :, :] = rng.random(shape[1:])
    pt = one_hot(shape, (shape[0] // 2, shape[1] // 2, shape[2] // 2), 2.0)
    sources = [("z", sheet, modulated(steps, steps / 3, steps / 10, 14.0, 3.0)),
               ("x", pt, gaussian(steps, steps / 4, steps / 12)),
               ("y", rng.random(shape) * (rng.random(shape) < 0.1), gaussian(steps, steps / 2, steps / 8, 0.5))]
    probes = [("Ez", rng.random(shape)), ("Hy", rng.random(shape) * (rng.random(shape) < 0.3)),
              ("Dx", one_hot(shape, (shape[0] - 1, 0, shape[2] - 1))), ("Ex", rng.standard_normal(shape)),
              ("Hz", one_hot(shape, (0, shape[1] - 1, 0)))]
    return dict(eps=eps, dL=DL, npml=list(npml), steps=steps, sources=sources, probes=probes,
                snapshots=tuple(snapshots or (steps // 2, steps)))


def field_case(name):
    if name == "c1_tm":
        return _c1("z")
    if name == "c1_te":
        return _c1("x")
    table = {
        "pml3d":     ((16, 14, 12), (4, 3, 5), 150, 11),
        "mixed_pml": ((12, 10, 9), (3, 0, 2), 100, 12),
        "periodic":  ((9, 8, 7), (0, 0, 0), 80, 13),
        "nx1":       ((1, 12, 10), (0, 3, 2), 80, 14),
        "odd2d":     ((31, 17, 1), (5, 4, 0), 200, 15),
        "line1d":    ((40, 1, 1), (6, 0, 0), 120, 16),
        "tall_z":    ((6, 5, 70), (0, 0, 9), 90, 17),
    }
    return _small3d(*table[name])


FIELD_CASES = ("c1_tm", "c1_te", "pml3d", "mixed_pml", "periodic", "nx1", "odd2d", "line1d", "tall_z")
SMALL_FIELD_CASES = FIELD_CASES[2:]


def grad_case(name):
    """Gradient cases.  'ref_rev_E'/'ref_rev_H'/'ref_fwd_E'/'ref_fwd_H' restate
    tests/test_gradients_fdtd.py:66-165 (8x8x1, npml [2,2,0], 500 steps) with a seeded eps."""
    if name.startswith("ref_"):
        shape = (8, 8, 1)
        steps = 500
        eps = np.random.default_rng(21).random(shape) + 1
        prof = one_hot(shape, (4, 4, 0), 1.0)
        comp = "z" if name == "ref_rev_E" else "x"
        keys = ("Ex", "Ey", "Ez") if name.endswith("E") else ("Hx", "Hy", "Hz")
        return dict(eps=eps, dL=DL, npml=[2, 2, 0], steps=steps,
                    sources=[(comp, prof, gaussian(steps, 300, 20, 1.0))],
                    probes=[(k, np.ones(shape)) for k in keys], mode=name.split("_")[1])
    if name == "c4_small":
        return scaled_case("c4_small")
    if name == "probe3d":
        # SURVEY appendix B check: PML on all axes, two sources, dense weights on E / H / D
        case = _small3d((10, 9, 7), (3, 2, 2), 60, 31)
        case["mode"] = "rev"
        return case
    raise KeyError(name)


GRAD_CASES = ("ref_rev_E", "ref_rev_H", "ref_fwd_E", "ref_fwd_H", "probe3d", "c4_small")


def objective_weights(steps, n_probes, seed=5):
    """Fixed cotangent for scalar objectives L = sum(w * series**2) on the probe series."""
    return np.random.default_rng(seed).random((steps, n_probes))


# ---- scaled copies of the BASELINE configs -------------------------------------------------------------
def _guide(shape, core=5.9536, split=True):
    """Config-2 style geometry: a guide along x (half-widths scale with the grid) that splits into two arms."""
    Nx, Ny, Nz = shape
    eps = np.ones(shape)
    cy, cz = Ny // 2, Nz // 2
    hy, hz = max(1, Ny // 24), max(1, Nz // 32) if Nz > 1 else 1
    off_max = Ny // 6
    for i in range(Nx):
        d = 0
        if split and i >= Nx // 2:
            d = int(round(off_max * min(1.0, (i - Nx // 2) / max(1, Nx // 2 - Nx // 8))))
        for c in ({cy} if d == 0 else {cy - d, cy + d}):
            eps[i, c - hy:c + hy, max(0, cz - hz):cz + hz if Nz > 1 else 1] = core
    return eps, (cy, cz, hy, hz, off_max)


def scaled_case(name):
    """'c2_96': config 2 (3-D splitter, PML on all axes, Jz sheet source, arm probes) at 96^3 for 2000 steps.
    'c4_small': config 4 (gradient of a windowed mode-overlap-like objective w.r.t. eps_r) at 32x28x20, 300 steps.
    'c5_small': config 5 (forward-mode JVP over eps_r directions, 2-D TM) at 256x256, 1000 steps, 4 directions."""
    from .fdtd_numpy import C_0, time_step
    dt = time_step(DL)
    omega = 2 * np.pi * C_0 / 2e-6
    if name == "c2_96":
        shape, steps = (96, 96, 96), 2000
        eps, (cy, cz, hy, hz, off) = _guide(shape)
        prof = np.zeros(shape); prof[14, cy - hy:cy + hy, cz - hz:cz + hz] = 1.0
        t = np.arange(steps)
        wave = 5 * np.exp(-(t - 300) ** 2 / (2 * 60.0 ** 2)) * np.cos(omega * dt * t)
        probes = []
        for c in (cy - off, cy + off):
            m = np.zeros(shape); m[82, c - hy:c + hy, cz - hz:cz + hz] = 1.0
            probes.append(("Ez", m))
        m = np.zeros(shape); m[40, cy - hy:cy + hy, cz - hz:cz + hz] = 1.0
        probes.append(("Hy", m))
        return dict(eps=eps, dL=DL, npml=[12, 12, 12], steps=steps, sources=[("z", prof, wave)], probes=probes,
                    snapshots=(steps,))
    if name == "c4_small":
        shape, steps = (32, 28, 20), 300
        eps, (cy, cz, hy, hz, off) = _guide(shape, split=False)
        rng = np.random.default_rng(1)
        eps[12:20, cy - 4:cy + 4, cz - 2:cz + 2] = 1 + 4.95 * rng.random((8, 8, 4))      # the design box
        prof = np.zeros(shape); prof[6, cy - hy:cy + hy, cz - hz:cz + hz] = 1.0
        t = np.arange(steps)
        wave = 3 * np.exp(-(t - 80) ** 2 / (2 * 25.0 ** 2)) * np.cos(omega * dt * t)
        yy = np.exp(-((np.arange(shape[1]) - cy) / 2.0) ** 2)[:, None] * np.exp(-((np.arange(shape[2]) - cz) / 2.0) ** 2)[None, :]
        m = np.zeros(shape); m[26] = yy                                              # a mode-like overlap mask
        return dict(eps=eps, dL=DL, npml=[4, 4, 4], steps=steps, sources=[("z", prof, wave)], probes=[("Ez", m), ("Hy", m)],
                    mode="rev")
    if name == "c5_small":
        shape, steps, B = (256, 256, 1), 1000, 4
        eps = np.full(shape, 1.44 ** 2)
        eps[:, 120:136, 0] = 3.48 ** 2                                               # slab waveguide
        V = np.zeros((B,) + shape)
        for b in range(B):                                                           # four grating-tooth groups
            x0 = 70 + b * 36
            eps[x0:x0 + 12, 136:144, 0] = 3.48 ** 2
            V[b, x0:x0 + 12, 136:144, 0] = 1.0
        prof = np.zeros(shape); prof[30, 120:136, 0] = 1.0
        mask = np.zeros(shape); mask[60:220, 180, 0] = 1.0
        t = np.arange(steps)
        wave = np.exp(-(t - 150) ** 2 / (2 * 40.0 ** 2)) * np.cos(omega * dt * t)
        return dict(eps=eps, dL=DL, npml=[10, 10, 0], steps=steps, sources=[("z", prof, wave)],
                    probes=[("Ez", mask), ("Hx", mask)], directions=V)
    raise KeyError(name)
