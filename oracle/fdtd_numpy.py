"""numpy restatement of ceviche's FDTD step (TEST INFRASTRUCTURE - see oracle/__init__.py).

Follows, operation for operation, the reference at /root/reference:

* constants            ceviche/constants.py:7-10
* time step            ceviche/fdtd.py:213-222
* sigma profiles       ceviche/fdtd.py:224-263
* update coefficients  ceviche/fdtd.py:265-316
* Yee averaging        ceviche/utils.py:153-176
* curls                ceviche/derivatives.py:16-30
* one leap-frog step   ceviche/fdtd.py:74-144
* caller loop          ceviche/utils.py:316-332 (measure_fields)

The arithmetic (operand order included) is the reference's, so the result is
bit-identical to it; only the bookkeeping differs: the three vector components
live in length-3 lists indexed by a cyclic helper instead of 60 named
attributes, and the sigma arrays are built from six 1-D profiles (each
reference sigma array varies along one axis only) broadcast to the grid.

Parity: pinned (tests/test_oracle_vs_reference.py, tests/golden/).
"""
import numpy as np

# ceviche/constants.py:7-9 -- NOT the SI values; C_0 is derived.
EPSILON_0 = 8.85418782e-12
MU_0 = 1.25663706e-6
C_0 = 1 / np.sqrt(EPSILON_0 * MU_0)

COMPONENTS = ("x", "y", "z")
FIELD_KEYS = ("Ex", "Ey", "Ez", "Dx", "Dy", "Dz", "Hx", "Hy", "Hz")


def pad_to_3d(arr):
    """ceviche/utils.py:206-214 -- trailing singleton axes, ValueError beyond 3-D."""
    arr = np.asarray(arr)
    if arr.ndim > 3:
        raise ValueError("array is larger than 3 dimensional, given shape {}".format(arr.shape))
    return arr.reshape(arr.shape + (1,) * (3 - arr.ndim))


def time_step(dL, stability_factor=0.5):
    """ceviche/fdtd.py:219-222 (always the 3-D Courant bound)."""
    dL_sum = 3 / dL ** 2
    dL_avg = 1 / np.sqrt(dL_sum)
    return dL_avg / C_0 * stability_factor


def sigma_profiles(shape, npml, dt):
    """Six 1-D sigma profiles (sH_x, sH_y, sH_z), (sD_x, sD_y, sD_z).

    ceviche/fdtd.py:229-263: cubic grading on the doubled grid; H samples the
    odd doubled-grid indices of its own axis, D the even ones.
    """
    sH, sD = [], []
    for n_cells, p in zip(shape, npml):
        s2 = np.zeros(2 * n_cells)
        for n in range(2 * p):
            val = (0.5 * EPSILON_0 / dt) * (n / 2 / p) ** 3
            s2[2 * p - n + 1] = val
            s2[2 * n_cells - 2 * p + n] = val
        sH.append(s2[1::2].copy())
        sD.append(s2[0::2].copy())
    return sH, sD


def _along(profile, axis, shape, materialize):
    view = [1, 1, 1]
    view[axis] = shape[axis]
    arr = np.broadcast_to(profile.reshape(view), shape)
    return np.ascontiguousarray(arr) if materialize else arr


def yee_average(eps_r):
    """ceviche/utils.py:153-176: eps felt by Ex/Ey/Ez = mean with the previous cell (periodic)."""
    return [(eps_r + np.roll(eps_r, shift=1, axis=a)) / 2 for a in range(3)]


def curl_fwd(c, F, dL):
    """Component c of the forward-difference curl (ceviche/derivatives.py:16-22)."""
    u, v = (c + 1) % 3, (c + 2) % 3          # x:(y,z)  y:(z,x)  z:(x,y)
    return (np.roll(F[v], shift=-1, axis=u) - F[v]) / dL - (np.roll(F[u], shift=-1, axis=v) - F[u]) / dL


def curl_bwd(c, F, dL):
    """Component c of the backward-difference curl (ceviche/derivatives.py:24-30)."""
    u, v = (c + 1) % 3, (c + 2) % 3
    return (F[v] - np.roll(F[v], shift=1, axis=u)) / dL - (F[u] - np.roll(F[u], shift=1, axis=v)) / dL


class OracleFDTD:
    """State + one-step update with the reference's semantics (fp64)."""

    def __init__(self, eps_r, dL, npml, materialize=True):
        eps_r = pad_to_3d(np.asarray(eps_r, dtype=np.float64))
        self.shape = eps_r.shape
        self.dL = dL
        self.npml = list(npml)
        self.dt = time_step(dL)
        self.sH, self.sD = sigma_profiles(self.shape, self.npml, self.dt)
        self._materialize = materialize
        self._coefficients()
        self.set_eps(eps_r)

    # -- set-up ---------------------------------------------------------
    def _coefficients(self):
        """ceviche/fdtd.py:272-311 with mu_r = 1.  For component c the pair
        (a, b) is the sigma of the two other axes and `own` that of axis c."""
        dt, shape, mat = self.dt, self.shape, self._materialize
        self.mH = [None] * 3
        self.mD = [None] * 3
        for c in range(3):
            u, v = sorted(((c + 1) % 3, (c + 2) % 3))
            for name, prof in (("mH", self.sH), ("mD", self.sD)):
                a = _along(prof[u], u, shape, mat)
                b = _along(prof[v], v, shape, mat)
                own = _along(prof[c], c, shape, mat)
                m0 = (1 / dt + (a + b) / 2 / EPSILON_0 + a * b * dt / 4 / EPSILON_0 ** 2)
                m1 = (1 / m0 * (1 / dt - (a + b) / 2 / EPSILON_0 - a * b * dt / 4 / EPSILON_0 ** 2))
                if name == "mH":
                    m2 = (-1 / m0 * C_0 / 1.0)
                    m3 = (-1 / m0 * C_0 * dt * own / EPSILON_0 / 1.0)
                else:
                    m2 = (1 / m0 * C_0)
                    m3 = (1 / m0 * C_0 * dt * own / EPSILON_0)
                m4 = (-1 / m0 * dt * a * b / EPSILON_0 ** 2)
                getattr(self, name)[c] = (m1, m2, m3, m4)

    def set_eps(self, eps_r):
        """ceviche/fdtd.py:63-72: new permittivity => new mE and a field reset."""
        self.eps_r = eps_r
        self.eps_yee = yee_average(eps_r)
        self.mE = [1 / e for e in self.eps_yee]
        self.reset()

    def reset(self):
        """ceviche/fdtd.py:147-211."""
        z = lambda: [np.zeros(self.shape) for _ in range(3)]
        self.t_index = 0
        self.H, self.D, self.E = z(), z(), z()
        self.ICE, self.IH, self.ICH, self.ID = z(), z(), z(), z()

    # -- the step -------------------------------------------------------
    def step(self, Jx=None, Jy=None, Jz=None):
        """ceviche/fdtd.py:74-144.  Integrals accumulate before use, from the
        OLD H / D; J is added after the D update, unscaled."""
        dL = self.dL
        self.t_index += 1
        CE = [curl_fwd(c, self.E, dL) for c in range(3)]
        self.ICE = [self.ICE[c] + CE[c] for c in range(3)]
        self.IH = [self.IH[c] + self.H[c] for c in range(3)]
        self.H = [self.mH[c][0] * self.H[c] + self.mH[c][1] * CE[c]
                  + self.mH[c][2] * self.ICE[c] + self.mH[c][3] * self.IH[c] for c in range(3)]
        CH = [curl_bwd(c, self.H, dL) for c in range(3)]
        self.ICH = [self.ICH[c] + CH[c] for c in range(3)]
        self.ID = [self.ID[c] + self.D[c] for c in range(3)]
        D = [self.mD[c][0] * self.D[c] + self.mD[c][1] * CH[c]
             + self.mD[c][2] * self.ICH[c] + self.mD[c][3] * self.ID[c] for c in range(3)]
        for c, J in enumerate((Jx, Jy, Jz)):
            D[c] += 0 if J is None else J
        self.D = D
        self.E = [self.mE[c] * self.D[c] for c in range(3)]
        return self.fields()

    def fields(self):
        out = {}
        for c, n in enumerate(COMPONENTS):
            out["E" + n], out["D" + n], out["H" + n] = self.E[c], self.D[c], self.H[c]
        return out

    # -- the caller loop (ceviche/utils.py:316-332) ----------------------
    def run(self, steps, sources=(), probes=(), snapshots=()):
        """sources: [(component 'x'|'y'|'z', profile ndarray, waveform[steps])]
        probes:  [(field key e.g. 'Ez', mask ndarray)]
        Returns (series[steps, n_probes], {t: fields copy for t in snapshots})."""
        series = np.zeros((steps, len(probes)))
        snaps = {}
        for t in range(steps):
            J = {"x": None, "y": None, "z": None}
            for comp, profile, wave in sources:
                term = pad_to_3d(profile) * wave[t]
                J[comp] = term if J[comp] is None else J[comp] + term
            f = self.step(Jx=J["x"], Jy=J["y"], Jz=J["z"])
            for p, (key, mask) in enumerate(probes):
                series[t, p] = np.sum(f[key] * pad_to_3d(mask))
            if (t + 1) in snapshots:
                snaps[t + 1] = {k: v.copy() for k, v in f.items()}
        return series, snaps


def rel_l2(a, b):
    """||a-b|| / ||b|| with b the oracle; an all-zero oracle demands an all-zero a."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    nb = np.linalg.norm(b)
    if nb == 0.0:
        return 0.0 if np.linalg.norm(a) == 0.0 else np.inf
    return float(np.linalg.norm(a - b) / nb)
