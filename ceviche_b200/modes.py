"""Waveguide-mode source / probe profiles for the FDTD path (mirror of ceviche/modes.py:10-66).

Set-up glue, on the host like the reference's (scipy sparse eigensolver): the hot path only consumes the
resulting profile as a source or probe mask.  The reference builds the cross-section operator from its FDFD
derivative matrices (ceviche/derivatives.py:34-60, 94-216); here the two 1-D operators it needs are written
down directly:

    A = diag(eps) + (1/k0^2) * Dxf @ Dxb,   Dxf = diag(1/s_f) * (periodic forward difference) / dL,
                                            Dxb = diag(1/s_b) * (periodic backward difference) / dL

with the complex coordinate stretch s(l) = 1 - 1j * sigma(l) / (omega * eps0), sigma(l) = sigma_max (l/d)^m,
sigma_max = -(m+1) lnR / (2 eta0 d) -- the reference's `sig_w` / `s_value` (derivatives.py:209-216: m = 3,
lnR = -30), sampled at the reference's half-cell offsets (derivatives.py:189-207).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl

from .constants import C_0, EPSILON_0, ETA_0


def _stretch(direction, omega, dL, N, n_pml, m=3, lnR=-30.0):
    """1-D s-factor profile (derivatives.py:173-216)."""
    s = np.ones(N, dtype=np.complex128)
    if n_pml == 0:
        return s
    d = n_pml * dL
    sigma_max = -(m + 1) * lnR / (2 * ETA_0 * d)
    off_lo, off_hi = (0.5, -0.5) if direction == "f" else (1.0, -1.0)
    i = np.arange(N)
    lo, hi = i <= n_pml, i > N - n_pml
    l = np.zeros(N)
    l[lo] = dL * (n_pml - i[lo] + off_lo)
    l[hi] = dL * (i[hi] - (N - n_pml) + off_hi)
    pml = lo | hi
    s[pml] = 1 - 1j * sigma_max * (l[pml] / d) ** m / (omega * EPSILON_0)
    return s


def cross_section_operator(eps_cross, omega, dL, npml):
    """The sparse operator whose eigenvalues are n_eff^2 of the slab modes (modes.py:24-33)."""
    eps = np.asarray(eps_cross).reshape(-1)
    N = eps.size
    k0 = omega / C_0
    fwd = sp.diags([-np.ones(N), np.ones(N - 1), np.ones(1)], [0, 1, -(N - 1)], shape=(N, N), dtype=np.complex128) / dL
    bwd = sp.diags([np.ones(N), -np.ones(N - 1), -np.ones(1)], [0, -1, N - 1], shape=(N, N), dtype=np.complex128) / dL
    Dxf = sp.diags(1 / _stretch("f", omega, dL, N, npml)) @ fwd
    Dxb = sp.diags(1 / _stretch("b", omega, dL, N, npml)) @ bwd
    return sp.diags(eps.astype(np.complex128)) + (Dxf @ Dxb) * (1 / k0) ** 2


def get_modes(eps_cross, omega, dL, npml, m=1, filtering=True):
    """ Solve for the modes of a waveguide cross section (signature and return values of ceviche/modes.py:10-49)
            eps_cross: permittivity profile of the cross section (1-D)
            omega:     angular frequency
            dL:        grid size
            npml:      PML cells on each side of the cross section
            m:         number of modes
            filtering: drop modes with Re(eigenvalue) <= 0
        RETURNS vals (the eigenvalues = squared effective indices, as the reference returns them) and
                vectors (n_points, n_modes), each normalised to sum |v|^2 = 1.
    """
    eps_cross = np.asarray(eps_cross.detach().cpu() if hasattr(eps_cross, "detach") else eps_cross, dtype=np.float64)
    A = cross_section_operator(eps_cross, omega, dL, npml)
    n_max = np.sqrt(np.max(eps_cross))
    vals, vecs = spl.eigs(A, k=m, sigma=n_max ** 2, v0=None, which="LM")
    if filtering:
        keep = np.where(np.real(vals) > 0.0)[0]
        vals, vecs = vals[keep], vecs[:, keep]
    if vals.size == 0:
        raise BaseException("Could not find any eigenmodes for this waveguide")
    vecs = vecs / np.sqrt(np.sum(np.square(np.abs(vecs)), axis=0))
    return vals, vecs


def insert_mode(omega, dx, x, y, epsr, target=None, npml=0, m=1, filtering=False):
    """Mode m of the cross section epsr[x, y] written into `target[x, y]` (modes.py:52-66)."""
    epsr = np.asarray(epsr.detach().cpu() if hasattr(epsr, "detach") else epsr)
    if target is None:
        target = np.zeros(epsr.shape, dtype=complex)
    _, mode_field = get_modes(epsr[x, y], omega, dx, npml, m=m, filtering=filtering)
    target[x, y] = np.atleast_2d(mode_field)[:, m - 1].squeeze()
    return target
