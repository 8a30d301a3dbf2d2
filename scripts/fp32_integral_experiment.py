"""Where does the late-time fp32 error of the sigma-PML FDTD come from?  CPU experiment (numpy oracle, reference update
order fdtd.py:74-144) on a 2-D TM grid: fields / integrals stored in fp32 or fp64 in four combinations, against the all-fp64
run.  Prints ||diff|| / ||peak field|| and the rel-L2 at milestones.   python scripts/fp32_integral_experiment.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.fdtd_numpy import OracleFDTD  # noqa: E402
from oracle import cases  # noqa: E402


class Mixed(OracleFDTD):
    """OracleFDTD whose stored state is rounded after every step: fields to `fdt`, PML integrals to `idt`."""

    def __init__(self, *a, fdt=np.float32, idt=np.float32, **k):
        super().__init__(*a, **k)
        self.fdt, self.idt = fdt, idt

    def step(self, **J):
        f = super().step(**J)
        r = lambda arrs, dt: [a.astype(dt).astype(np.float64) for a in arrs]
        self.H, self.D = r(self.H, self.fdt), r(self.D, self.fdt)
        self.E = [self.mE[c] * self.D[c] for c in range(3)]
        self.ICE, self.IH, self.ICH, self.ID = (r(x, self.idt) for x in (self.ICE, self.IH, self.ICH, self.ID))
        return self.fields()


def main():
    shape, npml, steps = (120, 100, 1), [20, 20, 0], 10000
    rng = np.random.default_rng(0)
    eps = np.ones(shape)
    eps[40:80, 45:55, 0] = 5.9536
    prof = cases.one_hot(shape, (30, 50, 0), 1.0)
    t = np.arange(steps)
    wave = 5 * np.exp(-(t - 300) ** 2 / (2 * 60.0 ** 2)) * np.cos(0.15 * t)
    miles = (500, 1000, 2000, 5000, 10000)
    runs = {"ref": OracleFDTD(eps, cases.DL, npml),
            "f32 fields, f32 integrals": Mixed(eps, cases.DL, npml, fdt=np.float32, idt=np.float32),
            "f32 fields, f64 integrals": Mixed(eps, cases.DL, npml, fdt=np.float32, idt=np.float64),
            "f64 fields, f32 integrals": Mixed(eps, cases.DL, npml, fdt=np.float64, idt=np.float32)}
    peak = 0.0
    for n in range(1, steps + 1):
        J = prof * wave[n - 1]
        out = {k: r.step(Jz=J) for k, r in runs.items()}
        ref = np.concatenate([out["ref"][k].ravel() for k in ("Ez", "Hx", "Hy")])
        peak = max(peak, np.linalg.norm(ref))
        if n in miles:
            for k in runs:
                if k == "ref":
                    continue
                v = np.concatenate([out[k][q].ravel() for q in ("Ez", "Hx", "Hy")])
                d = np.linalg.norm(v - ref)
                print("step %5d  %-28s rel %.2e   diff/peak %.2e   (|ref|/peak %.1e)" % (n, k, d / np.linalg.norm(ref), d / peak, np.linalg.norm(ref) / peak), flush=True)


if __name__ == "__main__":
    main()
