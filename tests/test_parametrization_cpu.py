"""Design parametrisations (ceviche_b200/parametrization.py) against the formulas of the reference's examples
(examples/optimize_mode_converter.py:51-72, examples/forwardmode_grating_coupler.py:147-162) restated with numpy / scipy.
The examples themselves import skimage and HIPS autograd, which are not installed: the disc kernel is restated from
skimage.draw.circle's definition (pixels with centre distance^2 < radius^2 of the call, here radius + 1)."""
import numpy as np
import torch
from scipy.signal import convolve2d

from ceviche_b200 import parametrization as P


def _ref_proj(rho, eta=0.5, beta=100):
    return np.divide(np.tanh(beta * eta) + np.tanh(beta * (rho - eta)), np.tanh(beta * eta) + np.tanh(beta * (1 - eta)))


def _ref_kernel(radius):
    k = np.zeros((2 * radius + 1, 2 * radius + 1))
    for r in range(2 * radius + 1):
        for c in range(2 * radius + 1):
            if (r - radius) ** 2 + (c - radius) ** 2 < (radius + 1) ** 2:      # skimage.draw.circle(radius, radius, radius + 1)
                k[r, c] = 1
    return k / k.sum()


def _ref_blur(rho, radius):
    return convolve2d(rho, _ref_kernel(radius), mode="full")[radius:-radius, radius:-radius]


def test_proj_blur_make_rho_match_the_example_formulas():
    rng = np.random.default_rng(0)
    rho = rng.random((31, 27))
    region = np.zeros((31, 27)); region[6:25, 5:22] = 1
    for eta, beta in ((0.5, 100), (0.3, 8.0)):
        assert np.allclose(P.operator_proj(torch.as_tensor(rho), eta, beta).numpy(), _ref_proj(rho, eta, beta), rtol=1e-13, atol=1e-15)
    for radius in (1, 2, 3, 5):
        assert np.allclose(P.disc_kernel(radius), _ref_kernel(radius))
        assert np.allclose(P.operator_blur(torch.as_tensor(rho), radius).numpy(), _ref_blur(rho, radius), rtol=1e-12, atol=1e-14)
        want = rho * (region == 0) + _ref_blur(rho, radius) * region
        assert np.allclose(P.make_rho(rho, region, radius).numpy(), want, rtol=1e-12, atol=1e-14)
    # the radius-3 disc is not the full square (its corners are cut), the radius-2 one is
    assert P.disc_kernel(2).min() > 0 and P.disc_kernel(3)[0, 0] == 0


def test_parametrisation_is_differentiable_in_both_modes():
    rng = np.random.default_rng(1)
    rho = torch.as_tensor(rng.random((12, 10)), dtype=torch.float64).requires_grad_(True)
    region = torch.zeros((12, 10), dtype=torch.float64); region[3:9, 2:8] = 1
    f = lambda r: (P.operator_proj(P.make_rho(r, region, 2), 0.5, 6.0) ** 2).sum()
    (g,) = torch.autograd.grad(f(rho), rho)
    v = torch.as_tensor(rng.standard_normal((12, 10)))
    _, jv = torch.func.jvp(f, (rho.detach(),), (v,))
    assert abs(float(jv) - float((g * v).sum())) <= 1e-12 * abs(float(jv))
    h = 1e-6
    fd = (f(rho.detach() + h * v) - f(rho.detach() - h * v)) / (2 * h)
    assert abs(float(fd) - float(jv)) <= 1e-7 * abs(float(jv))


def test_grating_coupler_geometry_and_fill_factor_directions():
    G = P.grating_coupler(256, 256, 5e-8, 20, groups=4)
    ff = torch.full((4,), 0.5, dtype=torch.float64)
    eps = G.eps_r(ff)
    assert eps.shape == (256, 256, 1) and float(eps.min()) == 1.0
    assert abs(float(eps.max()) - 3.48 ** 2) < 1e-9
    # the sigmoid projection of forwardmode_grating_coupler.py:151-162 on the teeth rows
    col, row = int(G.x_grids[len(G.x_grids) // 3]), G.y_teeth[0]
    dens = np.sin(2 * np.pi * G.dl * col / G.Lambda / 2) ** 2
    want = G.eps_base[col, row] + (3.48 ** 2 - 1.44 ** 2) / (np.exp(-15 * (dens - 0.5)) + 1)
    assert abs(float(eps[col, row, 0]) - want) <= 1e-12 * want
    # analytic directions = autograd Jacobian = finite differences; one group per column, supports disjoint
    V = G.fill_factor_directions(ff)
    J = torch.autograd.functional.jacobian(lambda f: G.eps_r(f), ff)          # [Nx, Ny, 1, groups]
    assert torch.allclose(V, J.permute(3, 0, 1, 2), rtol=1e-12, atol=1e-14)
    h = 1e-6
    for g in range(4):
        e = torch.zeros(4, dtype=torch.float64); e[g] = h
        fd = (G.eps_r(ff + e) - G.eps_r(ff - e)) / (2 * h)
        assert float((fd - V[g]).abs().max()) <= 1e-6 * float(V[g].abs().max())
        assert float(V[g].abs().max()) > 0
    assert float((V.abs() > 0).sum(0).max()) == 1
