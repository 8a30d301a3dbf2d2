"""Inverse-design loop helpers (mirror of ceviche/optimizers.py:5-59) over torch tensors: parameters and
gradients stay on the device between iterations (the reference round-trips numpy arrays through autograd boxes)."""
import time

import numpy as np
import torch


def step_adam(gradient, mopt_old, vopt_old, iteration, beta1, beta2, epsilon=1e-8):
    """ One step of ADAM (optimizers.py:49-59) """
    mopt = beta1 * mopt_old + (1 - beta1) * gradient
    mopt_t = mopt / (1 - beta1 ** (iteration + 1))
    vopt = beta2 * vopt_old + (1 - beta2) * (gradient * gradient)
    vopt_t = vopt / (1 - beta2 ** (iteration + 1))
    grad_adam = mopt_t / ((vopt_t ** 0.5) + epsilon)
    return grad_adam, mopt, vopt


def adam_optimize(objective, params, jac, step_size=1e-2, Nsteps=100, bounds=None, direction='min', beta1=0.9,
                  beta2=0.999, callback=None, verbose=True):
    """Nsteps of ADAM on `objective` with gradient `jac` (same arguments and behaviour as optimizers.py:5-46:
    jac=True means objective returns (value, gradient); bounds clip abruptly).  `params` may be a torch tensor
    (any device) or a numpy array; values and gradients of either kind are accepted."""
    if direction not in ('min', 'max'):
        raise ValueError("The 'direction' parameter should be either 'min' or 'max'")
    of_list = []
    as_like = (lambda g, p: torch.as_tensor(g, dtype=p.dtype, device=p.device)) if torch.is_tensor(params) \
        else (lambda g, p: np.asarray(g.detach().cpu() if torch.is_tensor(g) else g))
    mopt = vopt = None
    for iteration in range(Nsteps):
        if callback:
            callback(iteration, of_list, params)
        t_start = time.time()
        if jac is True:
            of, grad = objective(params)
        else:
            of = objective(params)
            grad = jac(params)
        t_elapsed = time.time() - t_start
        of_list.append(float(of.detach()) if torch.is_tensor(of) else of)
        if verbose:
            print("Epoch: %3d/%3d | Duration: %.2f secs | Value: %5e" % (iteration + 1, Nsteps, t_elapsed, of_list[-1]))
        grad = as_like(grad, params).reshape(params.shape)
        if iteration == 0:
            mopt, vopt = grad * 0, grad * 0
        grad_adam, mopt, vopt = step_adam(grad, mopt, vopt, iteration, beta1, beta2)
        params = params - step_size * grad_adam if direction == 'min' else params + step_size * grad_adam
        if bounds:
            params = params.clamp(bounds[0], bounds[1]) if torch.is_tensor(params) else np.clip(params, bounds[0], bounds[1])
    return params, of_list
