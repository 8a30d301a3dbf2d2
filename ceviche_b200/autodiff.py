"""Differentiation of the FDTD path w.r.t. eps_r (custom VJP / JVP).  Filled in below."""
import torch


def needs_grad(sim, J):
    if not torch.is_grad_enabled():
        return False
    ts = list(sim._mE64) + [j for j in J if j is not None] + list(sim._H) + list(sim._D)
    return any(t.requires_grad for t in ts)


def step(sim, J):
    raise NotImplementedError


def run(sim, steps, waveforms):
    raise NotImplementedError
