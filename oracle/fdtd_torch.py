"""torch-CPU fp64 restatement of the ceviche FDTD step, used as the GRADIENT oracle
(TEST INFRASTRUCTURE - see oracle/__init__.py).

The reference gets its derivatives from HIPS autograd tracing every
``autograd.numpy`` op of ``ceviche/fdtd.py:74-144`` (``ceviche/jacobians.py:29-51``).
HIPS autograd (requirements.txt:1, ``autograd>=1.3``, unpinned, not vendored
under /root/reference) is not installed here, so the same op sequence is written
over ``torch.roll`` and differentiated with ``torch.autograd`` /
``torch.func.jvp``.  Forward values are checked bit-for-bit against
``oracle.fdtd_numpy`` (itself pinned to the reference) in
``tests/test_oracle_vs_reference.py``; gradients are checked against finite
differences through the reference's numpy code -- the reference's own
criterion, tests/test_gradients_fdtd.py:52-64.
"""
import numpy as np
import torch

from . import fdtd_numpy as onp


class TorchFDTD:
    """Differentiable (w.r.t. eps_r) FDTD; eps_r enters only through mE = 1/eps_yee
    (ceviche/fdtd.py:67, 314-316)."""

    def __init__(self, eps_r, dL, npml):
        if not torch.is_tensor(eps_r):
            eps_r = torch.as_tensor(onp.pad_to_3d(np.asarray(eps_r, dtype=np.float64)))
        while eps_r.dim() < 3:
            eps_r = eps_r.unsqueeze(-1)
        self.shape = tuple(eps_r.shape)
        self.dL = dL
        # coefficients are plain constants in the reference too (fdtd.py:229-311 use np, not npa)
        helper = onp.OracleFDTD(np.ones(self.shape), dL, npml, materialize=False)
        self.dt = helper.dt
        as_t = lambda a: torch.as_tensor(np.ascontiguousarray(a))
        self.mH = [tuple(as_t(m) for m in helper.mH[c]) for c in range(3)]
        self.mD = [tuple(as_t(m) for m in helper.mD[c]) for c in range(3)]
        self.set_eps(eps_r)

    def set_eps(self, eps_r):
        self.eps_r = eps_r
        self.mE = [1 / ((eps_r + torch.roll(eps_r, 1, a)) / 2) for a in range(3)]
        self.reset()

    def reset(self):
        z = lambda: [torch.zeros(self.shape, dtype=torch.float64) for _ in range(3)]
        self.H, self.D, self.E = z(), z(), z()
        self.ICE, self.IH, self.ICH, self.ID = z(), z(), z(), z()

    def _curl_fwd(self, c, F):
        u, v = (c + 1) % 3, (c + 2) % 3
        dL = self.dL
        return (torch.roll(F[v], -1, u) - F[v]) / dL - (torch.roll(F[u], -1, v) - F[u]) / dL

    def _curl_bwd(self, c, F):
        u, v = (c + 1) % 3, (c + 2) % 3
        dL = self.dL
        return (F[v] - torch.roll(F[v], 1, u)) / dL - (F[u] - torch.roll(F[u], 1, v)) / dL

    def step(self, Jx=None, Jy=None, Jz=None):
        CE = [self._curl_fwd(c, self.E) for c in range(3)]
        self.ICE = [self.ICE[c] + CE[c] for c in range(3)]
        self.IH = [self.IH[c] + self.H[c] for c in range(3)]
        self.H = [self.mH[c][0] * self.H[c] + self.mH[c][1] * CE[c]
                  + self.mH[c][2] * self.ICE[c] + self.mH[c][3] * self.IH[c] for c in range(3)]
        CH = [self._curl_bwd(c, self.H) for c in range(3)]
        self.ICH = [self.ICH[c] + CH[c] for c in range(3)]
        self.ID = [self.ID[c] + self.D[c] for c in range(3)]
        D = [self.mD[c][0] * self.D[c] + self.mD[c][1] * CH[c]
             + self.mD[c][2] * self.ICH[c] + self.mD[c][3] * self.ID[c] for c in range(3)]
        for c, J in enumerate((Jx, Jy, Jz)):
            if J is not None:
                D[c] = D[c] + J
        self.D = D
        self.E = [self.mE[c] * self.D[c] for c in range(3)]
        out = {}
        for c, n in enumerate("xyz"):
            out["E" + n], out["D" + n], out["H" + n] = self.E[c], self.D[c], self.H[c]
        return out

    def run(self, steps, sources=(), probes=()):
        """Same conventions as OracleFDTD.run; returns series[steps, n_probes] (torch)."""
        rows = []
        for t in range(steps):
            J = {"x": None, "y": None, "z": None}
            for comp, profile, wave in sources:
                term = torch.as_tensor(onp.pad_to_3d(np.asarray(profile))) * float(wave[t])
                J[comp] = term if J[comp] is None else J[comp] + term
            f = self.step(Jx=J["x"], Jy=J["y"], Jz=J["z"])
            rows.append(torch.stack([torch.sum(f[key] * torch.as_tensor(onp.pad_to_3d(np.asarray(mask))))
                                     for key, mask in probes]))
        return torch.stack(rows)


def series_fn(shape, dL, npml, steps, sources, probes):
    """eps_r (torch, shape) -> probe series, a pure function for torch.autograd / torch.func."""
    def fn(eps_r):
        sim = TorchFDTD(eps_r.reshape(shape), dL, npml)
        return sim.run(steps, sources, probes)
    return fn
