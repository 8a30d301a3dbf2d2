"""A numpy/torch-CPU stand-in for one slab's compute (TEST INFRASTRUCTURE), so that the host-side slab
logic of ceviche_b200/slab.py -- partitioning, ring halo exchange schedule, point-set localisation,
probe reduction -- runs under gloo on CPU.  Arithmetic follows oracle/fdtd_numpy.py (the reference's op
order), restricted to an x-range with explicit halo planes."""
import numpy as np
import torch

from oracle import fdtd_numpy as onp


class NumpySlabBackend:
    is_cuda = False

    def __init__(self, shape_local, dL, dt, sH, sD, inv_eps):
        self.nx, self.Ny, self.Nz = self.shape = shape_local
        self.dL, self.dt = dL, dt
        z = lambda s=None: torch.zeros(s or self.shape, dtype=torch.float64)
        self.H, self.D = [z() for _ in range(3)], [z() for _ in range(3)]
        self.mE = [torch.as_tensor(np.ascontiguousarray(m)) for m in inv_eps]
        plane = (self.Ny, self.Nz)
        self.D_hi, self.mE_hi, self.H_lo = [None, z(plane), z(plane)], [None, z(plane), z(plane)], [None, z(plane), z(plane)]
        self.I = {k: [z().numpy() for _ in range(3)] for k in ("ICE", "IH", "ICH", "ID")}
        self.halo = False
        # full coefficient arrays exactly as the oracle builds them, from the LOCAL sigma profiles
        helper = onp.OracleFDTD.__new__(onp.OracleFDTD)
        helper.dt, helper.shape, helper.sH, helper.sD, helper._materialize = dt, self.shape, sH, sD, True
        helper._coefficients()
        self.mH_c, self.mD_c = helper.mH, helper.mD
        self.sources, self.probes, self.n_slots = [], [], 0

    def set_points(self, sources, probes):
        self.sources, self.probes = sources, probes
        self.n_slots = len(probes)
        self.n_sources = len(sources)

    def new_partials(self, steps):
        self.partials = torch.zeros((steps, len(self.probes)), dtype=torch.float64)

    def _E(self, c, lo, hi):
        """E_c = mE*D on planes [lo, hi] inclusive of the halo plane nx when hi == nx."""
        D, m = self.D[c].numpy(), self.mE[c].numpy()
        out = m[lo:min(hi, self.nx - 1) + 1] * D[lo:min(hi, self.nx - 1) + 1]
        if hi == self.nx:
            if self.halo:
                extra = (self.mE_hi[c].numpy() * self.D_hi[c].numpy())[None] if c > 0 else np.zeros((1, self.Ny, self.Nz))
            else:
                extra = (m[0] * D[0])[None]
            out = np.concatenate([out, extra], 0)
        return out

    def _probe(self, which, t):
        for p, (field, idx, w) in enumerate(self.probes):
            if (field >= 6) != (which == 1) or t < 0:
                continue
            c = field % 3
            arr = (self.mE[c] * self.D[c]) if field < 3 else (self.D[c] if field < 6 else self.H[c])
            self.partials[t, p] = float(np.sum(arr.numpy().reshape(-1)[idx] * w))

    def step_H(self, x0, x1, probe_t):
        if x1 <= x0:
            return
        self._probe(0, probe_t)
        dL = self.dL
        E = [self._E(c, x0, x1) for c in range(3)]          # planes x0..x1
        cur = [e[:-1] for e in E]
        nxt = [e[1:] for e in E]
        r = lambda a, ax: np.roll(a, -1, axis=ax)
        CE = [(r(cur[2], 1) - cur[2]) / dL - (r(cur[1], 2) - cur[1]) / dL,
              (r(cur[0], 2) - cur[0]) / dL - (nxt[2] - cur[2]) / dL,
              (nxt[1] - cur[1]) / dL - (r(cur[0], 1) - cur[0]) / dL]
        for c in range(3):
            H = self.H[c].numpy()
            m1, m2, m3, m4 = [m[x0:x1] for m in self.mH_c[c]]
            self.I["ICE"][c][x0:x1] = self.I["ICE"][c][x0:x1] + CE[c]
            self.I["IH"][c][x0:x1] = self.I["IH"][c][x0:x1] + H[x0:x1]
            H[x0:x1] = m1 * H[x0:x1] + m2 * CE[c] + m3 * self.I["ICE"][c][x0:x1] + m4 * self.I["IH"][c][x0:x1]

    def step_D(self, x0, x1, probe_t, wave_row):
        if x1 <= x0:
            return
        self._probe(1, probe_t)
        dL = self.dL
        Hn = [h.numpy() for h in self.H]

        def with_prev(c):
            if x0 > 0:
                return Hn[c][x0 - 1:x1]
            first = (self.H_lo[c].numpy() if (self.halo and c > 0) else Hn[c][self.nx - 1])[None]
            return np.concatenate([first, Hn[c][0:x1]], 0)
        ext = [with_prev(c) for c in range(3)]
        cur = [e[1:] for e in ext]
        prv = [e[:-1] for e in ext]
        r = lambda a, ax: np.roll(a, 1, axis=ax)
        CH = [(cur[2] - r(cur[2], 1)) / dL - (cur[1] - r(cur[1], 2)) / dL,
              (cur[0] - r(cur[0], 2)) / dL - (cur[2] - prv[2]) / dL,
              (cur[1] - prv[1]) / dL - (cur[0] - r(cur[0], 1)) / dL]
        for c in range(3):
            D = self.D[c].numpy()
            m1, m2, m3, m4 = [m[x0:x1] for m in self.mD_c[c]]
            self.I["ICH"][c][x0:x1] = self.I["ICH"][c][x0:x1] + CH[c]
            self.I["ID"][c][x0:x1] = self.I["ID"][c][x0:x1] + D[x0:x1]
            D[x0:x1] = m1 * D[x0:x1] + m2 * CH[c] + m3 * self.I["ICH"][c][x0:x1] + m4 * self.I["ID"][c][x0:x1]
        plane = self.Ny * self.Nz
        for s, (comp, idx, w) in enumerate(self.sources):      # sources of the planes just updated
            sel = (idx >= x0 * plane) & (idx < x1 * plane)
            self.D[comp].numpy().reshape(-1)[idx[sel]] += w[sel] * float(wave_row[s])

    def sample(self, which, t):
        self._probe(which, t)

    def series(self):
        return self.partials.clone()

    def field(self, key):
        c = "xyz".index(key[1])
        if key[0] == "E":
            return self.mE[c] * self.D[c]
        return (self.D if key[0] == "D" else self.H)[c]
