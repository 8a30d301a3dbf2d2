"""Host-side glue of SURVEY 8(f) ranks 3-4: the waveguide-mode profiles (ceviche/modes.py) against the reference's
own code (where /root/reference exists) and against a committed golden fixture generated from it; the ADAM loop
(ceviche/optimizers.py) against a plain numpy restatement."""
import os

import numpy as np
import pytest

from oracle import ref_loader

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "modes_ridge.npz")


def _ridge():
    """The reference's own self-test geometry (modes.py:140-165): a ridge of eps 4 in air."""
    from ceviche_b200.constants import C_0
    lambda0 = 1.550e-6
    dL = lambda0 / 100
    omega = 2 * np.pi * C_0 / lambda0
    Nx = int(lambda0 * 10 / dL)
    eps = np.ones((Nx,))
    w = int(lambda0 / dL / 2)
    eps[Nx // 2 - w:Nx // 2 + w] = 4.0
    return eps, omega, dL


def _match(vals, vecs, ref_vals, ref_vecs):
    order, rorder = np.argsort(-vals.real), np.argsort(-ref_vals.real)
    np.testing.assert_allclose(vals[order], ref_vals[rorder], rtol=1e-9, atol=1e-9)
    for a, b in zip(order, rorder):           # eigenvectors up to a phase
        ov = abs(np.vdot(ref_vecs[:, b], vecs[:, a]))
        assert abs(ov - 1) < 1e-7, ov


def test_modes_match_golden_fixture():
    from ceviche_b200 import modes
    eps, omega, dL = _ridge()
    g = np.load(GOLD)
    vals, vecs = modes.get_modes(eps, omega, dL, npml=10, m=6)
    _match(vals, vecs, g["vals"], g["vecs"])
    tgt = modes.insert_mode(omega, dL, slice(None), 3, np.tile(eps[:, None], (1, 8)), npml=10, m=2)
    assert tgt.shape == (eps.size, 8) and np.all(tgt[:, :3] == 0) and np.all(tgt[:, 4:] == 0)
    assert abs(abs(np.vdot(g["vecs"][:, np.argsort(-g["vals"].real)[1]], tgt[:, 3])) - 1) < 1e-7


@pytest.mark.skipif(not ref_loader.available(), reason="reference sources not present")
def test_modes_match_reference_code():
    from ceviche_b200 import modes
    ref = ref_loader.load_modes()
    eps, omega, dL = _ridge()
    for npml, m in ((10, 6), (0, 3), (25, 4)):
        rv, rvec = ref.get_modes(eps, omega, dL, npml, m=m)
        v, vec = modes.get_modes(eps, omega, dL, npml, m=m)
        _match(v, vec, rv, rvec)
    # the operator itself, entry for entry
    mats = ref_loader.load().__dict__  # noqa: F841  (reference fdtd module loaded => ceviche.derivatives present)
    import sys
    D = sys.modules["ceviche.derivatives"].compute_derivative_matrices(omega, (eps.size, 1), [10, 0], dL=dL)
    import scipy.sparse as sp
    A_ref = sp.spdiags(eps, [0], eps.size, eps.size) + D[0].dot(D[1]) * (1 / (omega / ref.C_0)) ** 2
    A = modes.cross_section_operator(eps, omega, dL, 10)
    assert abs(A - A_ref).max() <= 1e-12 * abs(A_ref).max()


def test_adam_matches_numpy_restatement():
    import torch
    from ceviche_b200.optimizers import adam_optimize
    rng = np.random.default_rng(0)
    Q = rng.standard_normal((6, 6)); Q = Q @ Q.T + np.eye(6)
    b = rng.standard_normal(6)
    f = lambda p: 0.5 * p @ Q @ p - b @ p
    df = lambda p: Q @ p - b
    p0 = rng.standard_normal(6)
    # optimizers.py:5-59 restated in plain numpy
    p, m, v = p0.copy(), np.zeros(6), np.zeros(6)
    hist = []
    for it in range(40):
        hist.append(f(p)); g = df(p)
        m = 0.9 * m + 0.1 * g; v = 0.999 * v + 0.001 * g * g
        p = p - 0.05 * (m / (1 - 0.9 ** (it + 1))) / (np.sqrt(v / (1 - 0.999 ** (it + 1))) + 1e-8)
        p = np.clip(p, -0.8, 0.8)
    pt, of = adam_optimize(lambda q: (0.5 * q @ torch.as_tensor(Q) @ q - torch.as_tensor(b) @ q, torch.as_tensor(Q) @ q - torch.as_tensor(b)),
                           torch.as_tensor(p0), True, step_size=0.05, Nsteps=40, bounds=(-0.8, 0.8), verbose=False)
    np.testing.assert_allclose(pt.numpy(), p, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(of, hist, rtol=1e-12)
    pn, of2 = adam_optimize(f, p0, df, step_size=0.05, Nsteps=40, bounds=(-0.8, 0.8), verbose=False)
    np.testing.assert_allclose(pn, p, rtol=1e-12, atol=1e-14)
    with pytest.raises(ValueError):
        adam_optimize(f, p0, df, direction='sideways', verbose=False)
