"""Differentiation of the FDTD path with respect to eps_r.

The reference gets FDTD derivatives from HIPS autograd tracing every numpy op of every step
(ceviche/fdtd.py:2, ceviche/jacobians.py:29-51); each non-traceable operator there is registered with
`defvjp` / `defjvp` (ceviche/primitives.py:28-54).  Here the FDTD step itself is that operator:

* `_StepFn`  - one `forward()` call as a torch.autograd.Function whose backward is ONE transposed step
               (cev_fdtd_adjoint_step), so reference-style loops `for t: fields = F.forward(...)` stay
               differentiable in reverse mode with O(steps x 3 arrays) memory (only D is saved: the step
               is linear in the state and bilinear in (1/eps, D));
* `_RunFn`   - the fused `run()` with a custom backward: checkpointed, time-reversed adjoint FDTD
               (forward snapshots every K steps, each segment recomputed once storing D per step);
* `jvp_run`  - forward mode: the primal and a BATCH of tangent states advance in one sweep
               (cev_fdtd_jvp_run), instead of one complete traced run per direction (jacobians.py:43).

Both Functions are written in the `setup_context` style and hand their forward-mode rule to a second Function
(`_StepTanFn`, `_RunTanFn`) that carries a `vmap` rule: `torch.vmap(torch.func.jvp(fun))` -- which is how
`jacobian(mode='forward')` pushes a batch of input directions through `fun` in ONE evaluation -- then reaches
the tangent kernels with all B directions at once (`_RunTanFn.vmap`: one cev_fdtd_jvp_run sweep with B tangent states;
`_StepTanFn.vmap`: the tangent step of each state).  The kernels only ever see plain tensors: Function.forward runs
below every torch.func level, and `base()` strips the wrappers where host code keeps tensors on the object.

eps_r enters only through mE = 1/eps_yee (fdtd.py:67, 314-316); both Functions take the three fp64 mE
arrays as differentiable inputs and torch chains through `1/x` and the Yee averaging on its own.
"""
import contextlib
import ctypes as C
import math

import torch
import torch.autograd.forward_ad as fwAD
from torch._C._functorch import get_unwrapped as _unwrap_one, is_batchedtensor as _is_batched, \
    is_functorch_wrapped_tensor as _is_wrapped

from . import _lib

_FAMS = ("ICE", "IH", "ICH", "ID")


def base(t):
    """The plain tensor under the wrappers of torch.func transforms (jvp / grad levels); `t` itself when there are none.
    A tensor batched by torch.vmap has no single plain value: a batch of permittivities is not supported."""
    while _is_wrapped(t):
        if _is_batched(t):
            raise NotImplementedError("torch.vmap over eps_r / the FDTD state is not supported (only over tangent "
                                      "directions, which is what jacobian(mode='forward') does)")
        t = _unwrap_one(t)
    return t


def plain():
    """Context for host-side bookkeeping (state allocation, point sets, staging buffers, plain kernel launches): while a
    torch.func transform is active EVERY op, even a factory call, returns a tensor wrapped at the current level, and a
    wrapper has no data pointer to hand to the C ABI.  Inside this context the interpreter stack is cleared, so plain
    tensors in give plain tensors out."""
    if torch._C._are_functorch_transforms_active():
        from torch._functorch.pyfunctorch import temporarily_clear_interpreter_stack
        return temporarily_clear_interpreter_stack()
    return contextlib.nullcontext()


def needs_grad(sim, J):
    """True when the step must go through the differentiable Functions: some input carries a
    reverse-mode graph or a forward-mode tangent."""
    ts = list(sim._mE64) + [j for j in J if j is not None] + list(sim._H) + list(sim._D)
    if sim._pml is not None:
        ts += [t for fam in _FAMS for t in sim._pml[fam]]
    if torch._C._are_functorch_transforms_active() and any(_is_wrapped(t) for t in ts):
        from torch._functorch.pyfunctorch import GradInterpreter, retrieve_all_functorch_interpreters
        if any(isinstance(i, GradInterpreter) for i in retrieve_all_functorch_interpreters()):
            # (their wrappers do not say requires_grad: the step would silently run as a constant)
            raise NotImplementedError("torch.func.grad / vjp / jacrev through the FDTD step are not supported: use "
                                      "torch.autograd (jacobian(mode='reverse')); torch.func.jvp and torch.vmap over "
                                      "tangent directions are")
    if torch.is_grad_enabled() and any(t.requires_grad for t in ts):
        return True
    if fwAD._current_level >= 0:
        return any(fwAD.unpack_dual(t).tangent is not None for t in ts)
    return False


def _p3(ts):
    return _lib.c_void_p3(*[None if (t is None or t.numel() == 0) else t.data_ptr() for t in ts])


def _state(H, D, mE, pml):
    st = _lib.cev_state()
    st.H, st.D, st.inv_eps = _p3(H), _p3(D), _p3(mE)
    for f, fam in enumerate(_FAMS):
        setattr(st, fam, _p3(pml[3 * f:3 * f + 3]))
    return st


def _adjoint(lH, lD, lpml, gC2, G, box=None, gC=None):
    adj = _lib.cev_adjoint()
    adj.lH, adj.lD, adj.gC2, adj.G_mE = _p3(lH), _p3(lD), _p3(gC2), _p3(G)
    if gC is not None:
        adj.gC = _p3(gC)
    for f, fam in enumerate(_FAMS):
        setattr(adj, "l" + fam, _p3(lpml[3 * f:3 * f + 3]))
    if box is not None:
        adj.g_box = (C.c_int64 * 6)(*[int(v) for pair in box for v in pair])
    return adj


def _grad_box(sim):
    """The box of 1/eps_yee cells whose G_mE the design region needs: eps_r[i,j,k] enters eps_xx at (i,j,k) and
    (i+1,j,k) etc. (utils.py:167-174), so one more cell on the high side of every axis; None = the whole grid (also
    when the +1 would wrap)."""
    region = getattr(sim, "design_region", None)
    if region is None:
        return None
    box = []
    for (lo, hi), n in zip(region, sim.grid_shape):
        lo, hi = int(lo), int(hi)
        if not (0 <= lo < hi <= n):
            raise ValueError("design_region {} outside the grid {}".format(region, sim.grid_shape))
        if lo == 0 and hi == n:          # the whole axis (e.g. the z axis of a 2-D grid): nothing to add
            box.append((0, n))
            continue
        if hi + 1 > n:
            return None
        box.append((lo, hi + 1))
    return box


def _record_fits(sim, box, steps):
    """Can the reverse sweep run from a D-box record (no checkpoints, no recomputation)?  The tensor-map adjoint
    kernels must serve the grid, and the record (steps + 1 slots of the box) must fit in a quarter of the free memory."""
    plan = sim._ensure_plan()
    if not plan.lib.cev_fdtd_adjoint_boxed_supported(plan.handle) or sim._n_mon_pts > 0:
        return False
    cells = 1
    for lo, hi in box:
        cells *= hi - lo
    need = (steps + 1) * 3 * cells * (8 if sim.dtype == torch.float64 else 4)
    free, _ = torch.cuda.mem_get_info(sim.device)
    return need <= free // 4


def _flat_pml(sim):
    return [t for fam in _FAMS for t in sim._pml[fam]]


# ----------------------------------------------------------------------------- one step
class _StepTape:
    """What one step leaves behind for its derivative rules: 1/eps in storage precision, D before and after."""
    __slots__ = ("mE", "D_in", "D_out", "has_J")


def _tangent_step(sim, tape, dts):
    """Tangent of one step: the same linear step on the tangent state with J = dJ, except
    dE = mE dD + dmE D (product rule on fdtd.py:135-137), once for the E feeding curl_E (primal D
    before the step) and once for the returned E' (primal D after the step).  `dts`: 24 plain tensors / None."""
    mE, D_in, D_out = tape.mE, tape.D_in, tape.D_out
    plan = sim._ensure_plan()
    lib, h, s = plan.lib, plan.handle, sim._stream()
    shapes = plan.pml_shapes
    with torch.cuda.device(sim.device), torch.no_grad():
        zf = lambda: torch.zeros(sim.grid_shape, dtype=sim.dtype, device=sim.device)
        dmE = [zf() if t is None else t.to(sim.dtype).contiguous() for t in dts[0:3]]
        dH = [zf() if t is None else t.contiguous() for t in dts[3:6]]
        dD = [zf() if t is None else t.contiguous() for t in dts[6:9]]
        dJ = [None if (t is None or not tape.has_J[c]) else t.contiguous() for c, t in enumerate(dts[9:12])]
        dP = [torch.zeros(shapes[q], dtype=sim.dtype, device=sim.device) if t is None else t.clone().contiguous()
              for q, t in enumerate(dts[12:24])]
        dHn, dDn, dEn = [zf() for _ in range(3)], [zf() for _ in range(3)], [zf() for _ in range(3)]
        tst = _state(dH, dD, mE, dP)
        tan = _lib.cev_tangent()
        tan.d_inv_eps, tan.D_primal = _p3(dmE), _p3(D_in)
        _lib.check(lib.cev_fdtd_step_H_ex(h, C.byref(tst), C.byref(tan), _p3(dHn), 0, sim.Nx, -1, None, s))
        tst.H = _p3(dHn)
        _lib.check(lib.cev_fdtd_step_D(h, C.byref(tst), _p3(dDn), None, _p3(dJ), _lib.c_double3(1.0, 1.0, 1.0),
                                       0, sim.Nx, s))
        tst.D = _p3(dDn)
        tan.D_primal = _p3(D_out)
        _lib.check(lib.cev_fdtd_compute_E(h, C.byref(tst), C.byref(tan), _p3(dEn), s))
    return tuple(dHn + dDn + dEn + dP)


class _StepTanFn(torch.autograd.Function):
    """The forward-mode rule of `_StepFn` as an operator of its own, so that it can carry a vmap rule."""

    @staticmethod
    def forward(sim, tape, *dts):
        return _tangent_step(sim, tape, dts)

    @staticmethod
    def setup_context(ctx, inputs, output):
        pass

    @staticmethod
    def vmap(info, in_dims, sim, tape, *dts):
        """B tangent states through the same primal step (torch.vmap over tangent directions)."""
        dims = in_dims[2:]
        outs = []
        for b in range(info.batch_size):
            one = [None if t is None else (t if d is None else t.movedim(d, 0)[b]) for t, d in zip(dts, dims)]
            outs.append(_tangent_step(sim, tape, one))
        return tuple(torch.stack([o[k] for o in outs]) for k in range(21)), (0,) * 21


class _StepFn(torch.autograd.Function):
    """inputs : mE64[3], H[3], D[3], J[3] (zeros-size tensor = absent), pml[12]
    outputs: H'[3], D'[3], E'[3], pml'[12]"""

    @staticmethod
    def forward(sim, *ts):
        mE64, H, D, J, pml = ts[0:3], ts[3:6], ts[6:9], ts[9:12], ts[12:24]
        plan = sim._ensure_plan()
        lib, h, s = plan.lib, plan.handle, sim._stream()
        with torch.cuda.device(sim.device):
            mE = [m.detach().to(sim.dtype).contiguous() for m in mE64]
            Hn = [torch.empty_like(t) for t in H]
            Dn = [torch.empty_like(t) for t in D]
            En = [torch.empty_like(t) for t in D]
            pml_n = [t.detach().clone() for t in pml]           # integrals advance in place on the copies
            st = _state([t.detach() for t in H], [t.detach() for t in D], mE, pml_n)
            _lib.check(lib.cev_fdtd_step_H(h, C.byref(st), _p3(Hn), 0, sim.Nx, s))
            st.H = _p3(Hn)
            Jp = [None if j.numel() == 0 else j.detach().contiguous() for j in J]
            _lib.check(lib.cev_fdtd_step_D(h, C.byref(st), _p3(Dn), _p3(En), _p3(Jp), _lib.c_double3(1.0, 1.0, 1.0),
                                           0, sim.Nx, s))
        tape = _StepTape()
        tape.mE, tape.D_in, tape.D_out = mE, [t.detach() for t in D], [t.detach() for t in Dn]   # (aliases without grad_fn: no cycle)
        tape.has_J = [j.numel() != 0 for j in J]
        sim._step_tape = tape           # picked up by setup_context (which has no other channel from forward)
        return tuple(Hn + Dn + En + pml_n)

    @staticmethod
    def setup_context(ctx, inputs, output):
        sim = inputs[0]
        tape = sim._step_tape
        ctx.sim, ctx.tape, ctx.has_J = sim, tape, tape.has_J
        ctx.save_for_backward(*tape.mE, *tape.D_in, *tape.D_out)

    @staticmethod
    def jvp(ctx, _sim, *dts):
        return _StepTanFn.apply(ctx.sim, ctx.tape, *dts)

    @staticmethod
    def vmap(info, in_dims, sim, *ts):
        raise NotImplementedError("torch.vmap over eps_r / the FDTD state is not supported (only over tangent directions)")

    @staticmethod
    def backward(ctx, *gs):
        sim = ctx.sim
        saved = ctx.saved_tensors
        mE, D_in, D_out = saved[0:3], saved[3:6], saved[6:9]
        plan = sim._ensure_plan()
        shapes = plan.pml_shapes
        with torch.cuda.device(sim.device):
            z = lambda ref: torch.zeros_like(ref)
            gH = [g.detach().clone().contiguous() if g is not None else z(D_in[c]) for c, g in enumerate(gs[0:3])]
            gD = [g.detach().clone().contiguous() if g is not None else z(D_in[c]) for c, g in enumerate(gs[3:6])]
            gE = gs[6:9]
            gp = [g.detach().clone().contiguous() if g is not None
                  else torch.zeros(shapes[q], dtype=sim.dtype, device=sim.device) for q, g in enumerate(gs[9:21])]
            G = [torch.zeros(D_in[c].shape, dtype=torch.float64, device=sim.device) for c in range(3)]
            for c in range(3):       # E' = mE * D'  (fdtd.py:135-137)
                if gE[c] is not None:
                    gD[c] += mE[c] * gE[c]
                    G[c] += gE[c].double() * D_out[c].double()
            gJ = [gD[c].clone() if ctx.has_J[c] else None for c in range(3)]   # D' = ... + J
            gC2 = [torch.empty_like(D_in[c]) for c in range(3)]
            fwd = _state(D_in, D_in, mE, [None] * 12)      # only inv_eps and D (= D before the step) are read
            adj = _adjoint(gH, gD, gp, gC2, G)
            _lib.check(plan.lib.cev_fdtd_adjoint_step(plan.handle, C.byref(fwd), C.byref(adj), sim._stream()))
        grads_J = [g if g is not None else None for g in gJ]
        return (None, *G, *gH, *gD, *grads_J, *gp)


def step(sim, J):
    """Differentiable `forward()`: returns (H', D', E', pml dict)."""
    empty = torch.zeros(0, dtype=sim.dtype, device=sim.device)
    Jt = [empty if j is None else j for j in J]
    out = _StepFn.apply(sim, *sim._mE64, *sim._H, *sim._D, *Jt, *_flat_pml(sim))
    sim._step_tape = None           # every level's setup_context has taken it by now
    H, D, E, p = list(out[0:3]), list(out[3:6]), list(out[6:9]), out[9:21]
    pml = {fam: list(p[3 * f:3 * f + 3]) for f, fam in enumerate(_FAMS)}
    return H, D, E, pml


# ----------------------------------------------------------------------------- fused run
class _RunTape:
    """What a differentiable run() leaves behind: the D-box record or the checkpoints (reverse mode) and the inputs of a
    tangent sweep (forward mode)."""
    record = box = every = checkpoints = None


def _tangent_sweep(sim, tape, dmE_batch):
    """Forward mode through the fused run: ONE sweep (cev_fdtd_jvp_run) advancing a scratch primal state and the B tangent
    states of `dmE_batch` = [[d(1/eps_x), d(1/eps_y), d(1/eps_z)] per direction] from zero fields.  Returns the tangents of
    the probe series, [B, steps, n_probes]."""
    plan = sim._ensure_plan()
    B = len(dmE_batch)
    sim._apply_active(tape.active)
    with torch.cuda.device(sim.device), torch.no_grad():
        z = lambda ts: [torch.zeros_like(t) for t in ts]
        H, D, P = z(sim._H), z(sim._D), z(_flat_pml(sim))
        keep, tsts, tans = [], (_lib.cev_state * B)(), (_lib.cev_tangent * B)()
        for b, dm in enumerate(dmE_batch):
            dmE = [torch.zeros_like(tape.mE[c]) if t is None else t.detach().to(sim.dtype).contiguous() for c, t in enumerate(dm)]
            tH, tD, tP = z(sim._H), z(sim._D), z(_flat_pml(sim))
            keep.append((dmE, tH, tD, tP))
            tsts[b] = _state(tH, tD, tape.mE, tP)
            tans[b].d_inv_eps, tans[b].D_primal = _p3(dmE), _p3(D)
        partials = torch.zeros((tape.steps, sim._n_slots), dtype=torch.float64, device=sim.device)
        tpart = torch.zeros((B, tape.steps, sim._n_slots), dtype=torch.float64, device=sim.device)
        st = _state(H, D, tape.mE, P)
        _lib.check(plan.lib.cev_fdtd_jvp_run(plan.handle, C.byref(st), B, tsts, tans, tape.steps, _ptr(tape.waveforms),
                                             _ptr(partials), _ptr(tpart), sim._stream()))
        sim._apply_active(sim._active)
        if tape.n_probes == 0:
            return torch.zeros((B, tape.steps, 0), dtype=torch.float64, device=sim.device)
        from .fdtd import fold_probes
        return fold_probes(plan, tpart, tape.n_probes, sim._stream())


class _RunTanFn(torch.autograd.Function):
    """The forward-mode rule of `_RunFn` as an operator of its own: under torch.vmap (a batch of directions, as
    jacobian(mode='forward') builds it) its vmap rule advances ALL of them in one sweep -- the reference does one complete
    traced run per direction (jacobians.py:38-51)."""

    @staticmethod
    def forward(sim, tape, dx, dy, dz):
        return _tangent_sweep(sim, tape, [[dx, dy, dz]])[0]

    @staticmethod
    def setup_context(ctx, inputs, output):
        pass

    @staticmethod
    def vmap(info, in_dims, sim, tape, dx, dy, dz):
        dims = in_dims[2:]
        batch = [[None if t is None else (t if d is None else t.movedim(d, 0)[b]) for t, d in zip((dx, dy, dz), dims)]
                 for b in range(info.batch_size)]
        return _tangent_sweep(sim, tape, batch), 0


class _RunFn(torch.autograd.Function):
    @staticmethod
    def forward(sim, steps, waveforms, every, mEx, mEy, mEz):
        tape = _RunTape()
        tape.steps, tape.waveforms = steps, waveforms
        sim._run_tape = tape            # picked up by setup_context
        box = _grad_box(sim)
        if box is None and getattr(sim, "design_region", None) is None and getattr(sim, "record_whole_grid", True):
            # gradient of every cell: the same trick with the whole grid as the box, where the record fits in memory
            box = [(0, int(n)) for n in sim.grid_shape]
        reverse = any(bool(m.requires_grad) for m in (mEx, mEy, mEz))     # (grad mode is off inside forward)
        if reverse and box is not None and every is None and steps > 0 and _record_fits(sim, box, steps):
            # gradients wanted inside a design box only: record D of the box after every step instead of checkpointing the
            # state -- the reverse sweep then needs no recomputation at all (cev_fdtd_adjoint_run_boxed)
            plan = sim._ensure_plan()
            with torch.cuda.device(sim.device):
                bx, by, bz = [hi - lo for lo, hi in box]
                rec = torch.empty((steps + 1, 3, bx, by, bz), dtype=sim.dtype, device=sim.device)
                sl = tuple(slice(lo, hi) for lo, hi in box)
                for c in range(3):
                    rec[0, c].copy_(sim._D[c][sl])
                flat = (C.c_int64 * 6)(*[int(v) for pair in box for v in pair])
                _lib.check(plan.lib.cev_fdtd_set_recorder(plan.handle, flat, rec[1].data_ptr(), steps))
                try:
                    out = sim._run_raw(steps, waveforms, refresh=True, fused=False)
                finally:
                    _lib.check(plan.lib.cev_fdtd_set_recorder(plan.handle, None, None, 0))
            tape.record, tape.box = rec, box
            tape.mE = [m.clone() for m in sim._mE]
            tape.n_probes = sim._n_probes
            tape.active = sim._active
            return out
        if not reverse:
            # forward mode only (no reverse graph wanted): nothing to checkpoint, one plain run
            with torch.cuda.device(sim.device):
                out = sim._run_raw(steps, waveforms, refresh=True)
            tape.checkpoints = None
            tape.mE = [m.clone() for m in sim._mE]
            tape.n_probes = sim._n_probes
            tape.active = sim._active
            return out
        every = max(1, int(every or math.ceil(math.sqrt(max(steps, 1)))))
        tape.every = every
        tape.checkpoints = []
        chunks = []
        with torch.cuda.device(sim.device):
            for t0 in range(0, steps, every):
                t1 = min(steps, t0 + every)
                tape.checkpoints.append((t0, t1, [t.clone() for t in sim._H], [t.clone() for t in sim._D],
                                         [t.clone() for t in _flat_pml(sim)]))
                chunks.append(sim._run_raw(t1 - t0, waveforms[t0:t1], refresh=False))
            sim._refresh_E()
        tape.mE = [m.clone() for m in sim._mE]
        tape.n_probes = sim._n_probes
        tape.active = sim._active
        return torch.cat(chunks) if chunks else torch.zeros((0, sim._n_probes), dtype=torch.float64, device=sim.device)

    @staticmethod
    def setup_context(ctx, inputs, output):
        ctx.sim = inputs[0]
        ctx.tape = inputs[0]._run_tape

    @staticmethod
    def jvp(ctx, _sim, _steps, _wf, _every, *dmE64):
        """Forward mode through the fused run (torch.autograd.forward_ad / jacobian(mode='forward')): a tangent sweep on
        scratch states restarted from zero fields, with d(1/eps_yee) as torch hands it over -- one direction under plain
        forward_ad, the whole batch in one sweep under torch.vmap (`_RunTanFn.vmap`)."""
        return _RunTanFn.apply(ctx.sim, ctx.tape, *dmE64)

    @staticmethod
    def vmap(info, in_dims, sim, *a):
        raise NotImplementedError("torch.vmap over eps_r / the FDTD state is not supported (only over tangent directions)")

    @staticmethod
    def backward(ctx, gbar):
        sim, tape = ctx.sim, ctx.tape
        if tape.record is None and tape.checkpoints is None:
            raise RuntimeError("this run() kept nothing for a reverse sweep: 1/eps did not require grad when it ran "
                               "(reverse mode through torch.func transforms is not supported: use torch.autograd)")
        plan = sim._ensure_plan()
        lib, h, s = plan.lib, plan.handle, sim._stream()
        shapes = plan.pml_shapes
        sim._apply_active(tape.active)       # the recomputation runs the forward kernels: same component set
        with torch.cuda.device(sim.device):
            gbar = gbar.detach().to(torch.float64).contiguous()
            zf = lambda: [torch.zeros(sim.grid_shape, dtype=sim.dtype, device=sim.device) for _ in range(3)]
            lH, lD = zf(), zf()
            gC2 = [torch.empty(sim.grid_shape, dtype=sim.dtype, device=sim.device) for _ in range(3)]
            gC = [torch.empty(sim.grid_shape, dtype=sim.dtype, device=sim.device) for _ in range(3)]
            lp = [torch.zeros(shapes[q], dtype=sim.dtype, device=sim.device) for q in range(12)]
            G = [torch.zeros(sim.grid_shape, dtype=torch.float64, device=sim.device) for _ in range(3)]
            if tape.record is not None:      # no recomputation: the forward run recorded D of the design box
                adj = _adjoint(lH, lD, lp, gC2, G, tape.box, gC)
                st = _state(lH, lH, tape.mE, [None] * 12)          # only inv_eps is read
                _lib.check(lib.cev_fdtd_adjoint_run_boxed(h, C.byref(st), tape.steps, _ptr(gbar) if tape.n_probes else None,
                                                          tape.record.data_ptr(), C.byref(adj), s))
                sim._apply_active(sim._active)
                return (None, None, None, None, *G)
            adj = _adjoint(lH, lD, lp, gC2, G, _grad_box(sim), gC)
            # D after every step of a segment (the only forward quantity the transposed step needs: the step is linear
            # in the state): one ring of slots for the whole sweep, written straight by the out-of-place D half-steps
            longest = max(t1 - t0 for t0, t1, *_ in tape.checkpoints)
            hist = torch.empty((longest + 1, 3) + tuple(sim.grid_shape), dtype=sim.dtype, device=sim.device)
            slots = (_lib.c_void_p3 * (longest + 1))(*[_p3(list(hist[k])) for k in range(longest + 1)])
            # scratch state of the recomputation: the same buffers for every segment (H and the integrals are copies, so a
            # second backward() finds the checkpoints intact; fixed buffers let the C side replay a captured segment)
            Hs = [torch.empty_like(t) for t in tape.checkpoints[0][2]]
            Ps = [torch.empty_like(t) for t in tape.checkpoints[0][4]]
            st = _state(Hs, list(hist[0]), tape.mE, Ps)
            for t0, t1, H0, D0, P0 in reversed(tape.checkpoints):
                # one C call per checkpoint segment: recompute + transposed steps (cev_fdtd_adjoint_run); the recomputation
                # advances H and the PML integrals it is given in place
                for c in range(3):
                    hist[0, c].copy_(D0[c])
                for dst, src in zip(Hs + Ps, list(H0) + list(P0)):
                    dst.copy_(src)
                _lib.check(lib.cev_fdtd_adjoint_run(h, C.byref(st), t1 - t0,
                                                    _ptr(tape.waveforms[t0:t1]) if sim._n_sources else None,
                                                    _ptr(gbar[t0:t1]) if tape.n_probes else None, slots, C.byref(adj), s))
            del hist
        sim._apply_active(sim._active)
        return (None, None, None, None, *G)


def _ptr(t):
    return None if t is None or t.numel() == 0 else t.data_ptr()


def run(sim, steps, waveforms, checkpoint_every=None):
    if sim.t_index != 0 or any(bool(t.requires_grad) for t in sim._H + sim._D):
        raise RuntimeError("a differentiable run() must start from initialize_fields(): gradients do not chain "
                           "across run() calls (use the per-step forward() API for that)")
    out = _RunFn.apply(sim, steps, waveforms.contiguous(), checkpoint_every, *sim._mE64)
    sim._run_tape = None            # the graph node(s) own the record / checkpoints from here on
    return out


# ----------------------------------------------------------------------------- forward mode
def jvp_run(sim, steps, waveforms, eps_tangents):
    """Primal + B tangents.  eps_tangents: [B, Nx, Ny, Nz] directions in eps_r.
    Returns (series [steps, P], dseries [B, steps, P])."""
    plan = sim._ensure_plan()
    if sim.t_index != 0 or any(bool(t.requires_grad) for t in sim._H + sim._D):
        # the tangent states start at zero: a primal state that already depends on eps_r (earlier run() / forward()
        # calls with this permittivity) would drop d(state)/d(eps) of that history -- same rule as the reverse-mode run()
        raise RuntimeError("jvp_run() must start from initialize_fields(): tangents do not chain across run() calls "
                           "(use the per-step forward() API under torch.autograd.forward_ad for that)")
    if sim._n_mon_pts > 0:
        raise RuntimeError("jvp_run() does not accumulate running-DFT monitors: clear them with set_monitors([], [])")
    v = eps_tangents.to(device=sim.device, dtype=torch.float64)
    if v.dim() == 3:
        v = v.unsqueeze(0)
    B = v.shape[0]
    eps_yee = (sim.eps_xx, sim.eps_yy, sim.eps_zz)
    with torch.cuda.device(sim.device), torch.no_grad():
        if sim._published:
            sim._H = [t.clone() for t in sim._H]
            sim._D = [t.clone() for t in sim._D]
            sim._published = False
        keep, tsts, tans = [], (_lib.cev_state * max(1, B))(), (_lib.cev_tangent * max(1, B))()
        for b in range(B):
            # d(eps_yee) = Yee average of v (utils.py:167-174); d(1/x) = -dx/x^2
            dmE = [(-((v[b] + torch.roll(v[b], 1, a)) / 2) / eps_yee[a].detach() ** 2).to(sim.dtype).contiguous()
                   for a in range(3)]
            tH = [torch.zeros_like(t) for t in sim._H]
            tD = [torch.zeros_like(t) for t in sim._D]
            tP = [torch.zeros_like(t) for t in _flat_pml(sim)]
            keep.append((dmE, tH, tD, tP))
            tsts[b] = _state(tH, tD, sim._mE, tP)
            tans[b].d_inv_eps = _p3(dmE)
            tans[b].D_primal = _p3(sim._D)
        partials = torch.zeros((steps, sim._n_slots), dtype=torch.float64, device=sim.device)
        tpart = torch.zeros((B, steps, sim._n_slots), dtype=torch.float64, device=sim.device)
        st = sim._state()
        _lib.check(plan.lib.cev_fdtd_jvp_run(plan.handle, C.byref(st), B, tsts, tans, steps, _ptr(waveforms),
                                             _ptr(partials), _ptr(tpart), sim._stream()))
        sim.t_index += steps
        sim._refresh_E()
        sim._tangent_states = keep
        if sim._n_probes == 0:
            return (torch.zeros((steps, 0), dtype=torch.float64, device=sim.device),
                    torch.zeros((B, steps, 0), dtype=torch.float64, device=sim.device))
        from .fdtd import fold_probes
        return (fold_probes(plan, partials, sim._n_probes, sim._stream()), fold_probes(plan, tpart, sim._n_probes, sim._stream()))
