"""HIPS-autograd binding of the FDTD run (soft dependency).

The reference differentiates everything with HIPS autograd and registers each operator that autograd cannot trace in
its operator-extension style: `@primitive` + `defvjp(f, vjp_maker_arg0, None, ...)` with
`vjp_maker(ans, *args) -> (v -> cotangent)` and `defjvp(f, jvp_arg0, ...)` with `jvp(g, ans, *args) -> tangent`
(ceviche/primitives.py:28-54, used at :58-259 and ceviche/utils.py:350-370).  The FDTD loop itself is traced op by op
there; here it is ONE operator, registered the same way, so that objectives written with `autograd.numpy` around it keep
working when HIPS autograd is installed:

    from ceviche_b200.hips import fdtd_series            # None when autograd is not importable
    series = fdtd_series(eps_r, F, steps, sources, probes)   # numpy in, numpy out; differentiable w.r.t. eps_r

The VJP is the checkpointed adjoint FDTD (ceviche_b200.autodiff._RunFn.backward), the JVP the tangent sweep
(fdtd.jvp_run).  HIPS autograd is not installed in the build image: `register()` is exercised in the tests against a
stand-in `extend` module that records the registrations, and the registered callables are checked on the GPU against
torch.autograd / jvp_run directly.
"""
import importlib.util

import numpy as np
import torch


def register(extend):
    """Define the primitive with `extend.primitive / defvjp / defjvp` and return it."""

    @extend.primitive
    def fdtd_series(eps_r, sim, steps, sources, probes):
        """probe series [steps, n_probes] (numpy) of `sim` run with permittivity `eps_r` (numpy, sim.grid_shape)"""
        sim.eps_r = torch.as_tensor(np.asarray(eps_r, dtype=np.float64)).reshape(sim.grid_shape)
        return sim.run(steps, sources, probes).detach().cpu().numpy()

    def vjp_maker_eps(ans, eps_r, sim, steps, sources, probes):
        def vjp(v):
            eps = torch.as_tensor(np.asarray(eps_r, dtype=np.float64), device=sim.device).reshape(sim.grid_shape)
            eps.requires_grad_(True)
            sim.eps_r = eps
            series = sim.run(steps, sources, probes)
            (g,) = torch.autograd.grad(series, eps, grad_outputs=torch.as_tensor(np.asarray(v), device=sim.device,
                                                                                 dtype=series.dtype))
            return g.cpu().numpy().reshape(np.shape(eps_r))
        return vjp

    def jvp_eps(g, ans, eps_r, sim, steps, sources, probes):
        sim.eps_r = torch.as_tensor(np.asarray(eps_r, dtype=np.float64)).reshape(sim.grid_shape)
        direction = torch.as_tensor(np.asarray(g, dtype=np.float64)).reshape((1,) + tuple(sim.grid_shape))
        _, dseries = sim.jvp_run(steps, direction, sources, probes)
        return dseries[0].cpu().numpy()

    extend.defvjp(fdtd_series, vjp_maker_eps, None, None, None, None)     # only eps_r (argnum 0) is differentiable
    extend.defjvp(fdtd_series, jvp_eps, None, None, None, None)
    return fdtd_series


fdtd_series = None
if importlib.util.find_spec("autograd") is not None:      # soft dependency
    try:
        import autograd.extend as _extend
        fdtd_series = register(_extend)
    except Exception:                                      # a stub / broken install: stay unregistered
        fdtd_series = None
