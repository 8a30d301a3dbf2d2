"""Inverse design on the B200 FDTD path, end to end (the loop of the reference's examples/optimize_*.py on the
time-domain solver): a waveguide-mode source (ceviche_b200.modes.insert_mode), a mode-overlap probe, the objective
from the probe series, its gradient with respect to the design region through the checkpointed adjoint FDTD, ADAM.

    python examples/optimize_mode_overlap.py [Nsteps]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceviche_b200
from ceviche_b200 import modes
from ceviche_b200.constants import C_0
from ceviche_b200.optimizers import adam_optimize


def build(Nx=120, Ny=80, npml=10, dL=5e-8, lambda0=1.0e-6, steps=600):
    omega = 2 * np.pi * C_0 / lambda0
    eps = np.ones((Nx, Ny, 1))
    core = slice(Ny // 2 - 4, Ny // 2 + 4)
    eps[:, core, 0] = 4.0
    eps[Nx // 2 - 12:Nx // 2 + 12, core, 0] = 1.0                        # a gap to be bridged by the design
    design = np.zeros((Nx, Ny, 1), dtype=bool)
    design[Nx // 2 - 12:Nx // 2 + 12, Ny // 2 - 10:Ny // 2 + 10, 0] = True
    # mode profiles of the (unbroken) guide cross-section at the source and probe planes
    eps_wg = np.ones((Nx, Ny)); eps_wg[:, core] = 4.0
    src = modes.insert_mode(omega, dL, npml + 6, slice(None), eps_wg, npml=npml, m=1).real[:, :, None]
    prb = modes.insert_mode(omega, dL, Nx - npml - 7, slice(None), eps_wg, npml=npml, m=1).real[:, :, None]
    F0 = ceviche_b200.fdtd(eps, dL, [npml, npml, 0])
    t = np.arange(steps) * F0.dt
    wave = np.exp(-(t - 150 * F0.dt) ** 2 / (2 * (40 * F0.dt) ** 2)) * np.cos(omega * t)
    return dict(eps=eps, design=design, src=src, prb=prb, wave=wave, dL=dL, npml=npml, steps=steps)


def make_objective(P, device="cuda", blur_radius=0, beta=None, eta=0.5):
    """blur_radius > 0 / beta given: the design density goes through the reference examples' parametrisation first --
    blur inside the design region (make_rho) and tanh projection (operator_proj), examples/optimize_mode_converter.py:51-72
    -- and torch differentiates through it on the way back."""
    from ceviche_b200.parametrization import make_rho, operator_proj
    eps0 = torch.as_tensor(P["eps"], device=device)
    design = torch.as_tensor(P["design"], device=device)
    n_design = int(design.sum())
    F = ceviche_b200.fdtd(eps0, P["dL"], [P["npml"], P["npml"], 0])
    # the reverse sweep only needs dL/d(1/eps) inside the design region: no recomputation where the grid allows it
    ix, iy = torch.nonzero(design[:, :, 0], as_tuple=True)
    F.design_region = ((int(ix.min()), int(ix.max()) + 1), (int(iy.min()), int(iy.max()) + 1), (0, 1))

    def objective(rho):
        """rho in [0, 1] per design cell -> (transmitted mode energy, d/d rho)"""
        rho = rho.detach().clone().requires_grad_(True)
        dens = rho
        if blur_radius > 0 or beta is not None:
            full = torch.zeros(design.shape[:2], dtype=torch.float64, device=device)
            full = full.masked_scatter(design[:, :, 0], rho)
            if blur_radius > 0:
                full = make_rho(full, design[:, :, 0], radius=blur_radius)
            if beta is not None:
                full = operator_proj(full, eta=eta, beta=beta)
            dens = full[design[:, :, 0]]
        eps = eps0.clone()
        eps[design] = 1.0 + 3.0 * dens
        F.eps_r = eps                                                   # fields reset, new graph (fdtd.py:63-72)
        series = F.run(P["steps"], [("z", P["src"], P["wave"])], [("Ez", P["prb"])])
        val = (series ** 2).sum()
        (g,) = torch.autograd.grad(val, rho)
        return val.detach(), g
    return objective, n_design


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    P = build()
    objective, n_design = make_objective(P)
    rho0 = torch.full((n_design,), 0.5, dtype=torch.float64, device="cuda")
    rho, history = adam_optimize(objective, rho0, True, step_size=0.05, Nsteps=n, bounds=(0.0, 1.0), direction="max")
    print("objective: %.4e -> %.4e" % (history[0], history[-1]))
