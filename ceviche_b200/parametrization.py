"""Design parametrisations of the reference's inverse-design examples, on the device and differentiable by torch in
both modes (reverse for the adjoint sweep, forward for the batched tangent sweep):

* `operator_proj`, `operator_blur`, `make_rho`  -- examples/optimize_mode_converter.py:51-72 (tanh density
  projection; blur by a disc kernel, applied inside the design region only);
* `sigmoid`, `projection`                        -- examples/forwardmode_grating_coupler.py:147-157 (sigmoid projection
  of a teeth density around 1 - fill_factor);
* `grating_coupler`                              -- the grating geometry of forwardmode_grating_coupler.py:33-98, 138-162
  scaled to an (Nx, Ny) grid with one fill factor per tooth group; `fill_factor_directions` gives d eps_r / d ff_g,
  the perturbation directions of BASELINE config 5.

The reference builds these with HIPS-autograd numpy on the host; values here are torch tensors wherever `rho` lives.
"""
import math

import numpy as np
import torch


def _t(x, like=None):
    if torch.is_tensor(x):
        return x
    return torch.as_tensor(np.asarray(x, dtype=np.float64), device=None if like is None else like.device)


def operator_proj(rho, eta=0.5, beta=100):
    """Density projection (optimize_mode_converter.py:51-55):
    (tanh(beta eta) + tanh(beta (rho - eta))) / (tanh(beta eta) + tanh(beta (1 - eta)))."""
    rho = _t(rho)
    return (math.tanh(beta * eta) + torch.tanh(beta * (rho - eta))) / (math.tanh(beta * eta) + math.tanh(beta * (1 - eta)))


def disc_kernel(radius):
    """The blur kernel of optimize_mode_converter.py:60-63: ones on skimage.draw.circle(radius, radius, radius + 1),
    i.e. the pixels of the (2 radius + 1)^2 window whose centre distance d obeys d^2 < (radius + 1)^2, normalised."""
    r = np.arange(2 * radius + 1) - radius
    k = ((r[:, None] ** 2 + r[None, :] ** 2) < (radius + 1) ** 2).astype(np.float64)
    return k / k.sum()


def operator_blur(rho, radius=2):
    """Blur by 2-D convolution with the disc kernel, 'full' mode cropped back to rho's shape, zero padding outside
    (optimize_mode_converter.py:57-65)."""
    rho = _t(rho)
    k = torch.as_tensor(disc_kernel(radius), dtype=rho.dtype, device=rho.device)
    # true convolution = cross-correlation with the flipped kernel (the disc is symmetric; flipped anyway)
    out = torch.nn.functional.conv2d(rho[None, None], torch.flip(k, (0, 1))[None, None], padding=radius)
    return out[0, 0]


def make_rho(rho, design_region, radius=2):
    """Blur inside the design region only (optimize_mode_converter.py:67-72)."""
    rho = _t(rho)
    design_region = _t(design_region, rho).to(rho.dtype)
    lpf_rho = operator_blur(rho, radius=radius) * design_region
    bg_rho = rho * (design_region == 0).to(rho.dtype)
    return bg_rho + lpf_rho


def sigmoid(x, strength=1):
    """Smooth projection from (-inf, inf) to (0, 1) (forwardmode_grating_coupler.py:147-149)."""
    return 1 / (torch.exp(-strength * _t(x)) + 1)


def projection(density, center, eps_min, eps_max, strength=15):
    """(eps_max - eps_min) sigmoid(strength (density - center)) (forwardmode_grating_coupler.py:151-157)."""
    return (eps_max - eps_min) * sigmoid(_t(density) - center, strength=strength)


class grating_coupler:
    """The grating coupler of examples/forwardmode_grating_coupler.py on an Nx x Ny grid (2-D, Nz = 1).

    Same construction as the example (`:33-98`): SiO2 substrate band, Si slab of thickness h1, teeth of thickness h0 - h1
    on top of it, period Lambda = lambda0 / (neff - sin(theta)); the teeth are the sigmoid projection of the density
    sin^2(pi x / Lambda) around 1 - ff (`:138-162`).  The example has 7 teeth and ONE fill factor; here the grating
    is as long as the grid allows and the teeth are cut into `groups` contiguous groups with one fill factor each, so
    that d eps_r / d ff_g, g = 0 .. groups-1, are `groups` independent perturbation directions (BASELINE config 5)."""

    lambda0 = 1550e-9
    neff_teeth, neff_hole = 2.846, 2.534
    theta = 20 / 360 * 2 * np.pi
    h0, h1 = 220e-9, 150e-9
    sub_eps, grating_eps = 1.44 ** 2, 3.48 ** 2

    def __init__(self, Nx, Ny, dl, npml, groups=16, ff=0.5, spc=1.5e-6, subs=1.5e-6, strength=15):
        self.Nx, self.Ny, self.dl, self.npml, self.groups, self.strength = Nx, Ny, dl, npml, groups, strength
        neff = ff * self.neff_teeth + (1 - ff) * self.neff_hole
        self.Lambda = self.lambda0 / (neff - np.sin(self.theta))
        period = int(self.Lambda / dl)
        x0 = npml + int(spc / dl)
        self.num_teeth = max(groups, (Nx - 2 * x0) // period - 1)
        self.x_grids = np.arange(x0, min(Nx - x0, x0 + int(self.Lambda * self.num_teeth / dl)))
        # vertical layout: centred on the grid (the example stacks spc / subs from the PML; on a square grid the slab
        # sits in the middle)
        y_slab = Ny // 2 - int(self.h0 / dl) // 2
        self.y_sub = (max(npml, y_slab - int(subs / dl)), min(Ny - npml, y_slab + int((subs + self.h0) / dl)))
        self.y_base = (y_slab, y_slab + int(self.h1 / dl))
        self.y_teeth = (y_slab + int(self.h1 / dl), y_slab + max(int(self.h0 / dl), int(self.h1 / dl) + 1))
        eps_base = np.ones((Nx, Ny))
        eps_base[:, self.y_sub[0]:self.y_sub[1]] = self.sub_eps
        eps_base[:, self.y_base[0]:self.y_base[1]] = self.grating_eps
        self.eps_base = eps_base
        density = np.zeros((Nx, Ny))
        density[self.x_grids, self.y_teeth[0]:self.y_teeth[1]] = np.square(np.sin(2 * np.pi * dl * self.x_grids / self.Lambda / 2))[:, None]
        self.teeth_density = density
        # tooth group of every column of the grating (contiguous groups of teeth)
        tooth = np.minimum((self.x_grids - x0) // period, self.num_teeth - 1)
        self.group_of_column = np.full(Nx, -1)
        self.group_of_column[self.x_grids] = np.minimum(tooth * groups // self.num_teeth, groups - 1)
        self.source_x = Nx - npml - int(spc / dl) // 2
        self.probe_y = min(Ny - npml - 2, self.y_sub[1] + int(spc / dl) // 2)

    def eps_r(self, ff, device=None, dtype=torch.float64):
        """eps_r (Nx, Ny, 1) for per-group fill factors `ff` [groups] (tensor: differentiable)."""
        ff = _t(ff).to(dtype=dtype, device=device)
        dens = torch.as_tensor(self.teeth_density, dtype=dtype, device=ff.device)
        grp = torch.as_tensor(self.group_of_column, device=ff.device)
        center = torch.where(grp >= 0, 1 - ff[grp.clamp(min=0)], torch.ones((), dtype=dtype, device=ff.device))[:, None]
        in_teeth = torch.zeros((self.Nx, self.Ny), dtype=dtype, device=ff.device)
        in_teeth[torch.as_tensor(self.x_grids, device=ff.device), self.y_teeth[0]:self.y_teeth[1]] = 1.0
        eps = torch.as_tensor(self.eps_base, dtype=dtype, device=ff.device) + \
            projection(dens, center, self.sub_eps, self.grating_eps, self.strength) * in_teeth
        return eps[:, :, None]

    def fill_factor_directions(self, ff, device=None, dtype=torch.float64):
        """d eps_r / d ff_g, shape [groups, Nx, Ny, 1] (analytic: each column belongs to one group)."""
        ff = _t(ff).to(dtype=dtype, device=device)
        dens = torch.as_tensor(self.teeth_density, dtype=dtype, device=ff.device)
        grp = torch.as_tensor(self.group_of_column, device=ff.device)
        center = torch.where(grp >= 0, 1 - ff[grp.clamp(min=0)], torch.ones((), dtype=dtype, device=ff.device))[:, None]
        sg = sigmoid(dens - center, self.strength)
        d_dff = (self.grating_eps - self.sub_eps) * self.strength * sg * (1 - sg)       # d/d ff = - d/d center
        in_teeth = torch.zeros((self.Nx, self.Ny), dtype=dtype, device=ff.device)
        in_teeth[torch.as_tensor(self.x_grids, device=ff.device), self.y_teeth[0]:self.y_teeth[1]] = 1.0
        d_dff = d_dff * in_teeth
        onehot = (grp[None, :] == torch.arange(self.groups, device=ff.device)[:, None]).to(dtype)      # [groups, Nx]
        return (onehot[:, :, None] * d_dff[None])[..., None]
