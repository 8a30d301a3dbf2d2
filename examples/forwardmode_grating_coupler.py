"""Forward-mode sensitivities of a grating coupler on the B200 FDTD path -- the time-domain part of the reference's
examples/forwardmode_grating_coupler.py (:138-240): a grating whose teeth are the sigmoid projection of a smooth density
around 1 - fill_factor, a pulsed source in the slab, the power through a line above the grating, and d(power)/d(fill
factor) by forward-mode differentiation.  The reference traces one complete run per parameter (jacobians.py:38-51); here
ALL fill factors (one per group of teeth) ride along in ONE sweep (fdtd.jvp_run: batched tangent launches), also when
asked for through `jacobian(objective, mode='forward')` as the reference's example does (`sensitivities_by_jacobian`).

    python examples/forwardmode_grating_coupler.py [N] [groups] [steps]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceviche_b200
from ceviche_b200.constants import C_0
from ceviche_b200.parametrization import grating_coupler


def build(N=512, groups=8, steps=1200, dL=5e-8, npml=20):
    G = grating_coupler(N, N, dL, npml, groups=groups)
    shape = (N, N, 1)
    prof = np.zeros(shape)
    prof[G.source_x, G.y_base[0]:G.y_teeth[1], 0] = 1.0                       # sheet across the slab, right of the grating
    mask = np.zeros(shape)
    mask[G.x_grids[0]:G.x_grids[-1], G.probe_y, 0] = 1.0                      # line above the grating
    dt = 0.5 * dL / (np.sqrt(3) * C_0)
    t = np.arange(steps) * dt
    omega0 = 2 * np.pi * C_0 / G.lambda0
    wave = np.exp(-(t - 280 * dt) ** 2 / (2 * (80 * dt) ** 2)) * np.cos(omega0 * t)
    return dict(G=G, prof=prof, mask=mask, wave=wave, steps=steps, dL=dL, npml=npml, groups=groups)


def power_and_sensitivities(P, ff, dtype=torch.float64):
    """ff: fill factors [groups] -> (power through the line = sum_t series^2, d power / d ff [groups]) in one sweep."""
    G = P["G"]
    ff = torch.as_tensor(ff, dtype=torch.float64, device="cuda")
    F = ceviche_b200.fdtd(G.eps_r(ff), P["dL"], [P["npml"], P["npml"], 0], dtype=dtype)
    series, dseries = F.jvp_run(P["steps"], G.fill_factor_directions(ff), [("z", P["prof"], P["wave"])], [("Ez", P["mask"])])
    power = (series ** 2).sum()
    dpower = (2 * series[None] * dseries).sum(dim=(1, 2))
    return power, dpower


def sensitivities_by_jacobian(P, ff, dtype=torch.float64):
    """The same numbers the way the reference's example asks for them (forwardmode_grating_coupler.py:200-240):
    `jacobian(objective, mode='forward')` of the power as a function of the fill factors.  `objective` is evaluated ONCE
    for all groups (torch.vmap over torch.func.jvp): the sigmoid projection is differentiated by torch, the run() inside
    becomes one batched tangent sweep."""
    G = P["G"]

    def objective(ff):
        F = ceviche_b200.fdtd(G.eps_r(ff.to("cuda")), P["dL"], [P["npml"], P["npml"], 0], dtype=dtype)
        series = F.run(P["steps"], [("z", P["prof"], P["wave"])], [("Ez", P["mask"])])
        return (series ** 2).sum()
    return ceviche_b200.jacobian(objective, mode='forward')(np.asarray(ff, dtype=np.float64))[0]


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    groups = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1200
    P = build(N, groups, steps)
    ff = np.full(groups, 0.5)
    power, dpower = power_and_sensitivities(P, ff)
    print("power through the line: %.6e" % float(power))
    for g, d in enumerate(dpower.tolist()):
        print("  d power / d ff[%d] = %+.6e" % (g, d))
    # the reference checks its forward-mode numbers against finite differences: one group here
    h, g = 1e-5, int(torch.argmax(dpower.abs()))
    e = np.zeros(groups); e[g] = h
    fd = (float(power_and_sensitivities(P, ff + e)[0]) - float(power_and_sensitivities(P, ff - e)[0])) / (2 * h)
    print("finite difference for group %d: %+.6e (forward mode: %+.6e)" % (g, fd, float(dpower[g])))
    dj = sensitivities_by_jacobian(P, ff)
    print("jacobian(mode='forward') [%s]: max |difference| to the sweep above %.3e" % (
        ceviche_b200.jacobians.last_forward_path, float((dj - dpower).abs().max())))
