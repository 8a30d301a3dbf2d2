"""Time-domain simulation of a waveguide splitter on the B200 FDTD path -- the workflow of the reference's
examples/simulate_splitter_fdtd.ipynb (cells 2-26): a pulsed Jz source in the input guide, the transmitted Ez recorded
through probe profiles in a straight reference guide and in the two arms of the splitter (`measure_fields`), field
snapshots (`aniplot`), power spectra (`get_spectral_power`, `get_max_power_freq`) and the arms' transmission normalised
to the straight guide.  The notebook loads its permittivity and source profiles from files produced by an FDFD inverse
design; here the geometry is built procedurally (same grid, same materials: 640 x 360 cells of 50 nm, eps 5.9536 in air,
20-cell PML, lambda = 2 um, 10 000 steps with the pulse at t0 = 2000).

    python examples/simulate_splitter_fdtd.py [Nx] [Ny] [steps]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CORE_EPS = 5.9536
LAMBDA0 = 2e-6


def geometry(Nx=640, Ny=360, npml=20, half_width=4):
    """(eps_wg, eps_r, J_in, J_wg, J_outs): straight guide, Y-splitter, the source profile in the input guide, the probe
    profile at the far end of the straight guide and one per arm of the splitter.  All (Nx, Ny, 1)."""
    cy = Ny // 2
    eps_wg = np.ones((Nx, Ny, 1))
    eps_wg[:, cy - half_width:cy + half_width, 0] = CORE_EPS
    eps_r = np.ones((Nx, Ny, 1))
    x_split, x_flat = Nx // 3, Nx - Nx // 4                  # the arms separate between these planes, then run straight
    off_max = Ny // 6
    offsets = np.zeros(Nx, dtype=int)
    for i in range(Nx):
        s = min(1.0, max(0.0, (i - x_split) / float(x_flat - x_split)))
        offsets[i] = int(round(off_max * (3 * s ** 2 - 2 * s ** 3)))      # smooth S-bend
        for c in ({cy} if offsets[i] == 0 else {cy - offsets[i], cy + offsets[i]}):
            eps_r[i, c - half_width:c + half_width, 0] = CORE_EPS
    # transverse profile of the source / probes: a cosine lobe over the core with evanescent-like tails (mode-shaped)
    y = np.arange(Ny)

    def lobe(centre):
        d = np.abs(y + 0.5 - centre)
        return np.where(d <= half_width, np.cos(0.35 * np.pi * d / half_width), np.cos(0.35 * np.pi) * np.exp(-(d - half_width) / 3.0))
    x_in, x_out = npml + 12, Nx - npml - 12
    J_in = np.zeros((Nx, Ny, 1)); J_in[x_in, :, 0] = lobe(cy)
    J_wg = np.zeros((Nx, Ny, 1)); J_wg[x_out, :, 0] = lobe(cy)
    J_outs = []
    for sign in (-1, +1):
        J = np.zeros((Nx, Ny, 1))
        prof = lobe(cy + sign * off_max)
        prof[(y < cy) if sign > 0 else (y >= cy)] = 0.0        # each arm's probe stays on its own side
        J[x_out, :, 0] = prof
        J_outs.append(J)
    return eps_wg, eps_r, J_in, J_wg, J_outs


def pulse(steps, dt, t0=2000, sigma=100, amp=5.0):
    """The notebook's source (cell 4): amp * exp(-(t - t0)^2 / 2 sigma^2) cos(omega dt t), omega = 2 pi C_0 / 2 um."""
    from ceviche_b200.constants import C_0
    t = np.arange(steps)
    return amp * np.exp(-(t - t0) ** 2 / 2.0 / sigma ** 2) * np.cos(2 * np.pi * C_0 / LAMBDA0 * dt * t)


def transmission(measured, measured_wg, dt):
    """Per-arm power transmission at the source's centre frequency: |FFT(arm)|^2 / |FFT(straight guide)|^2 there, plus the
    frequency of maximum power of the straight-guide signal (the notebook's plot_spectral_power / get_max_power_freq cells)."""
    from ceviche_b200.constants import C_0
    from ceviche_b200.utils import get_spectral_power
    freqs, p_wg = get_spectral_power(measured_wg, dt)
    _, p_arms = get_spectral_power(measured, dt)
    freqs = freqs.cpu().numpy()
    k = int(np.argmin(np.abs(freqs - C_0 / LAMBDA0)))
    T = (p_arms[k] / p_wg[k, 0]).cpu().numpy()
    half = len(freqs) // 2                                   # (the other half of a real series' spectrum is redundant)
    return T, float(freqs[int(np.argmax(p_wg[:half, 0].cpu().numpy()))])


def simulate(Nx=640, Ny=360, steps=10000, t0=2000, sigma=100, npml=20, dL=5e-8, dtype=None):
    import torch
    import ceviche_b200
    from ceviche_b200.utils import measure_fields
    eps_wg, eps_r, J_in, J_wg, J_outs = geometry(Nx, Ny, npml)
    dtype = dtype or torch.float64
    F = ceviche_b200.fdtd(eps_r, dL=dL, npml=[npml, npml, 0], dtype=dtype)
    F_wg = ceviche_b200.fdtd(eps_wg, dL=dL, npml=[npml, npml, 0], dtype=dtype)
    wave = pulse(steps, F.dt, t0, sigma)
    # (profile, waveform) sources run the fused device loop; a callable t -> J array runs the reference's own loop
    measured_wg = measure_fields(F_wg, (J_in, wave), steps, J_wg)
    measured = measure_fields(F, (J_in, wave), steps, J_outs)
    T, f_max = transmission(measured, measured_wg, F.dt)
    return dict(F=F, F_wg=F_wg, wave=wave, measured_wg=measured_wg, measured=measured, T=T, f_max=f_max,
                J_in=J_in, J_wg=J_wg, J_outs=J_outs)


if __name__ == "__main__":
    Nx = int(sys.argv[1]) if len(sys.argv) > 1 else 640
    Ny = int(sys.argv[2]) if len(sys.argv) > 2 else 360
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
    R = simulate(Nx, Ny, steps)
    print("straight guide: peak |Ez probe| %.4e, maximum power at %.4e Hz (source: %.4e Hz)" % (
        np.abs(R["measured_wg"]).max(), R["f_max"], 299792458.0 / LAMBDA0))
    print("splitter arms: power transmission at the centre frequency %s, total %.3f" % (
        np.array2string(R["T"], precision=3), float(R["T"].sum())))
    from ceviche_b200.utils import aniplot
    panels = aniplot(R["F"], lambda t: R["J_in"] * R["wave"][t], min(steps, 4000), num_panels=8, show=False)
    print("aniplot panels at time steps", [t for t, _ in panels])
