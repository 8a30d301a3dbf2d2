"""`jacobian(fun, argnum, mode, step_size)` with the reference's signature and modes
(ceviche/jacobians.py:16-73), over torch instead of HIPS autograd.

reverse   -> torch.autograd: one backward sweep per OUTPUT element (jacobians.py:29-35)
forward   -> torch forward-mode AD.  The reference runs `fun` once per INPUT direction (jacobians.py:38-51); here a
             batch of directions goes through ONE evaluation of `fun` (torch.vmap over torch.func.jvp, in chunks of
             `forward_chunk` directions): every fdtd.run() inside it becomes one batched tangent sweep
             (autodiff._RunTanFn.vmap -> cev_fdtd_jvp_run with B tangent states) and every fdtd.forward() one primal
             step plus B tangent steps.  A `fun` that torch.vmap cannot trace (in-place writes into unbatched tensors,
             .item(), numpy round trips) falls back to one pass per direction; `last_forward_path` says which ran.
numerical -> one-sided finite differences (jacobians.py:54-73)
Returns an (n_out, n_in) array (torch tensor on the input's device)."""
import numpy as np
import torch
import torch.autograd.forward_ad as fwAD


forward_chunk = 16          # directions per evaluation of `fun` in mode='forward' (B tangent states live at once)
last_forward_path = None    # 'batched' | 'per-direction: <why the batched pass was not possible>'


def _as_tensor(x):
    if torch.is_tensor(x):
        return x.detach().clone().to(torch.float64)
    return torch.as_tensor(np.asarray(x, dtype=np.float64))


def _flat(y):
    if not torch.is_tensor(y):
        y = torch.as_tensor(np.asarray(y, dtype=np.float64))
    return y.reshape(-1)


def jacobian(fun, argnum=0, mode='reverse', step_size=1e-6):
    if mode == 'reverse':
        inner = _reverse
    elif mode == 'forward':
        inner = _forward
    elif mode == 'numerical':
        inner = lambda f, x: _numerical(f, x, step_size)
    else:
        raise ValueError("'mode' kwarg must be either 'reverse' or 'forward' or 'numerical', given {}".format(mode))

    def wrapped(*args, **kwargs):
        def unary(x):
            a = list(args)
            a[argnum] = x
            return fun(*a, **kwargs)
        return inner(unary, _as_tensor(args[argnum]))
    return wrapped


def _reverse(fun, x):
    x = x.clone().requires_grad_(True)
    y = _flat(fun(x))
    rows = []
    for q in range(y.numel()):
        (g,) = torch.autograd.grad(y[q], x, retain_graph=True, allow_unused=True)
        rows.append(torch.zeros_like(x).reshape(-1) if g is None else g.reshape(-1))
    return torch.stack(rows)


def _forward(fun, x):
    global last_forward_path
    try:
        J = _forward_batched(fun, x)
        last_forward_path = "batched"
        return J
    except Exception as e:          # not traceable by torch.vmap: the reference's schedule, one pass per direction
        last_forward_path = "per-direction: {}: {}".format(type(e).__name__, str(e).splitlines()[0] if str(e) else "")
    return _forward_per_direction(fun, x)


def _forward_batched(fun, x):
    n = x.numel()
    if n == 0:
        raise ValueError("no input directions")
    E = torch.eye(n, dtype=x.dtype, device=x.device).reshape((n,) + tuple(x.shape))

    def column(t):
        return torch.func.jvp(lambda z: _flat(fun(z)), (x,), (t,))[1]
    cols = torch.vmap(column, chunk_size=max(1, min(int(forward_chunk), n)))(E)        # [n_in, n_out]
    return cols.detach().transpose(0, 1).contiguous()


def _forward_per_direction(fun, x):
    cols = []
    flat = x.reshape(-1)
    for q in range(flat.numel()):
        e = torch.zeros_like(flat)
        e[q] = 1.0
        with fwAD.dual_level():
            y = fun(fwAD.make_dual(x, e.reshape(x.shape)))
            t = fwAD.unpack_dual(_flat(y)).tangent
            cols.append(torch.zeros(_flat(y).numel(), dtype=torch.float64) if t is None else t.detach().clone())
    return torch.stack(cols, dim=1)


def _numerical(fun, x, step_size):
    with torch.no_grad():
        y0 = _flat(fun(x)).clone()
        flat = x.reshape(-1)
        cols = []
        for q in range(flat.numel()):
            xq = flat.clone()
            xq[q] += step_size
            cols.append((_flat(fun(xq.reshape(x.shape))) - y0) / step_size)
    return torch.stack(cols, dim=1)
