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

eps_r enters only through mE = 1/eps_yee (fdtd.py:67, 314-316); both Functions take the three fp64 mE
arrays as differentiable inputs and torch chains through `1/x` and the Yee averaging on its own.
"""
import ctypes as C
import math

import torch
import torch.autograd.forward_ad as fwAD

from . import _lib

_FAMS = ("ICE", "IH", "ICH", "ID")


def needs_grad(sim, J):
    """True when the step must go through the differentiable Functions: some input carries a
    reverse-mode graph or a forward-mode tangent."""
    ts = list(sim._mE64) + [j for j in J if j is not None] + list(sim._H) + list(sim._D)
    if sim._pml is not None:
        ts += [t for fam in _FAMS for t in sim._pml[fam]]
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
class _StepFn(torch.autograd.Function):
    """inputs : mE64[3], H[3], D[3], J[3] (zeros-size tensor = absent), pml[12]
    outputs: H'[3], D'[3], E'[3], pml'[12]"""

    @staticmethod
    def forward(ctx, sim, *ts):
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
        ctx.sim = sim
        ctx.has_J = [j.numel() != 0 for j in J]
        ctx.save_for_backward(*mE, *[t.detach() for t in D], *Dn)
        ctx.fw = (mE, [t.detach() for t in D], Dn)            # for forward mode (jvp)
        return tuple(Hn + Dn + En + pml_n)

    @staticmethod
    def jvp(ctx, _sim, *dts):
        """Tangent of one step: the same linear step on the tangent state with J = dJ, except
        dE = mE dD + dmE D (product rule on fdtd.py:135-137), once for the E feeding curl_E (primal D
        before the step) and once for the returned E' (primal D after the step)."""
        sim = ctx.sim
        mE, D_in, D_out = ctx.fw
        plan = sim._ensure_plan()
        lib, h, s = plan.lib, plan.handle, sim._stream()
        shapes = plan.pml_shapes
        with torch.cuda.device(sim.device):
            zf = lambda: torch.zeros(sim.grid_shape, dtype=sim.dtype, device=sim.device)
            dmE = [zf() if t is None else t.to(sim.dtype).contiguous() for t in dts[0:3]]
            dH = [zf() if t is None else t.contiguous() for t in dts[3:6]]
            dD = [zf() if t is None else t.contiguous() for t in dts[6:9]]
            dJ = [None if (t is None or not ctx.has_J[c]) else t.contiguous() for c, t in enumerate(dts[9:12])]
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
    H, D, E, p = list(out[0:3]), list(out[3:6]), list(out[6:9]), out[9:21]
    pml = {fam: list(p[3 * f:3 * f + 3]) for f, fam in enumerate(_FAMS)}
    return H, D, E, pml


# ----------------------------------------------------------------------------- fused run
class _RunFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sim, steps, waveforms, every, mEx, mEy, mEz):
        ctx.sim, ctx.steps, ctx.waveforms = sim, steps, waveforms
        ctx.record = None
        box = _grad_box(sim)
        if box is None and getattr(sim, "design_region", None) is None and getattr(sim, "record_whole_grid", True):
            # gradient of every cell: the same trick with the whole grid as the box, where the record fits in memory
            box = [(0, int(n)) for n in sim.grid_shape]
        if box is not None and every is None and steps > 0 and _record_fits(sim, box, steps):
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
            ctx.record, ctx.box = rec, box
            ctx.mE = [m.clone() for m in sim._mE]
            ctx.n_probes = sim._n_probes
            ctx.active = sim._active
            return out
        every = max(1, int(every or math.ceil(math.sqrt(max(steps, 1)))))
        ctx.every = every
        ctx.checkpoints = []
        chunks = []
        with torch.cuda.device(sim.device):
            for t0 in range(0, steps, every):
                t1 = min(steps, t0 + every)
                ctx.checkpoints.append((t0, t1, [t.clone() for t in sim._H], [t.clone() for t in sim._D],
                                        [t.clone() for t in _flat_pml(sim)]))
                chunks.append(sim._run_raw(t1 - t0, waveforms[t0:t1], refresh=False))
            sim._refresh_E()
        ctx.mE = [m.clone() for m in sim._mE]
        ctx.n_probes = sim._n_probes
        ctx.active = sim._active
        return torch.cat(chunks) if chunks else torch.zeros((0, sim._n_probes), dtype=torch.float64, device=sim.device)

    @staticmethod
    def jvp(ctx, _sim, _steps, _wf, _every, *dmE64):
        """Forward mode through the fused run (torch.autograd.forward_ad / jacobian(mode='forward')): one tangent sweep
        (cev_fdtd_jvp_run, B = 1) on scratch states restarted from zero fields, with d(1/eps_yee) as torch hands it over.
        For many directions call fdtd.jvp_run, which advances all of them in ONE sweep."""
        sim = ctx.sim
        plan = sim._ensure_plan()
        sim._apply_active(ctx.active)
        with torch.cuda.device(sim.device), torch.no_grad():
            z = lambda ts: [torch.zeros_like(t) for t in ts]
            dmE = [torch.zeros_like(ctx.mE[c]) if t is None else t.detach().to(sim.dtype).contiguous() for c, t in enumerate(dmE64)]
            H, D, P = z(sim._H), z(sim._D), z(_flat_pml(sim))
            tH, tD, tP = z(sim._H), z(sim._D), z(_flat_pml(sim))
            tsts, tans = (_lib.cev_state * 1)(), (_lib.cev_tangent * 1)()
            tsts[0] = _state(tH, tD, ctx.mE, tP)
            tans[0].d_inv_eps, tans[0].D_primal = _p3(dmE), _p3(D)
            partials = torch.zeros((ctx.steps, sim._n_slots), dtype=torch.float64, device=sim.device)
            tpart = torch.zeros((1, ctx.steps, sim._n_slots), dtype=torch.float64, device=sim.device)
            st = _state(H, D, ctx.mE, P)
            _lib.check(plan.lib.cev_fdtd_jvp_run(plan.handle, C.byref(st), 1, tsts, tans, ctx.steps, _ptr(ctx.waveforms),
                                                 _ptr(partials), _ptr(tpart), sim._stream()))
            sim._apply_active(sim._active)
            if ctx.n_probes == 0:
                return torch.zeros((ctx.steps, 0), dtype=torch.float64, device=sim.device)
            from .fdtd import fold_probes
            return fold_probes(plan, tpart, ctx.n_probes, sim._stream())[0]

    @staticmethod
    def backward(ctx, gbar):
        sim = ctx.sim
        plan = sim._ensure_plan()
        lib, h, s = plan.lib, plan.handle, sim._stream()
        shapes = plan.pml_shapes
        sim._apply_active(ctx.active)       # the recomputation runs the forward kernels: same component set
        with torch.cuda.device(sim.device):
            gbar = gbar.detach().to(torch.float64).contiguous()
            zf = lambda: [torch.zeros(sim.grid_shape, dtype=sim.dtype, device=sim.device) for _ in range(3)]
            lH, lD = zf(), zf()
            gC2 = [torch.empty(sim.grid_shape, dtype=sim.dtype, device=sim.device) for _ in range(3)]
            gC = [torch.empty(sim.grid_shape, dtype=sim.dtype, device=sim.device) for _ in range(3)]
            lp = [torch.zeros(shapes[q], dtype=sim.dtype, device=sim.device) for q in range(12)]
            G = [torch.zeros(sim.grid_shape, dtype=torch.float64, device=sim.device) for _ in range(3)]
            if ctx.record is not None:      # no recomputation: the forward run recorded D of the design box
                adj = _adjoint(lH, lD, lp, gC2, G, ctx.box, gC)
                st = _state(lH, lH, ctx.mE, [None] * 12)          # only inv_eps is read
                _lib.check(lib.cev_fdtd_adjoint_run_boxed(h, C.byref(st), ctx.steps, _ptr(gbar) if ctx.n_probes else None,
                                                          ctx.record.data_ptr(), C.byref(adj), s))
                sim._apply_active(sim._active)
                return (None, None, None, None, *G)
            adj = _adjoint(lH, lD, lp, gC2, G, _grad_box(sim), gC)
            # D after every step of a segment (the only forward quantity the transposed step needs: the step is linear
            # in the state): one ring of slots for the whole sweep, written straight by the out-of-place D half-steps
            longest = max(t1 - t0 for t0, t1, *_ in ctx.checkpoints)
            hist = torch.empty((longest + 1, 3) + tuple(sim.grid_shape), dtype=sim.dtype, device=sim.device)
            slots = (_lib.c_void_p3 * (longest + 1))(*[_p3(list(hist[k])) for k in range(longest + 1)])
            # scratch state of the recomputation: the same buffers for every segment (H and the integrals are copies, so a
            # second backward() finds the checkpoints intact; fixed buffers let the C side replay a captured segment)
            Hs = [torch.empty_like(t) for t in ctx.checkpoints[0][2]]
            Ps = [torch.empty_like(t) for t in ctx.checkpoints[0][4]]
            st = _state(Hs, list(hist[0]), ctx.mE, Ps)
            for t0, t1, H0, D0, P0 in reversed(ctx.checkpoints):
                # one C call per checkpoint segment: recompute + transposed steps (cev_fdtd_adjoint_run); the recomputation
                # advances H and the PML integrals it is given in place
                for c in range(3):
                    hist[0, c].copy_(D0[c])
                for dst, src in zip(Hs + Ps, list(H0) + list(P0)):
                    dst.copy_(src)
                _lib.check(lib.cev_fdtd_adjoint_run(h, C.byref(st), t1 - t0,
                                                    _ptr(ctx.waveforms[t0:t1]) if sim._n_sources else None,
                                                    _ptr(gbar[t0:t1]) if ctx.n_probes else None, slots, C.byref(adj), s))
            del hist
        sim._apply_active(sim._active)
        return (None, None, None, None, *G)


def _ptr(t):
    return None if t is None or t.numel() == 0 else t.data_ptr()


def run(sim, steps, waveforms, checkpoint_every=None):
    if sim.t_index != 0 or any(bool(t.requires_grad) for t in sim._H + sim._D):
        raise RuntimeError("a differentiable run() must start from initialize_fields(): gradients do not chain "
                           "across run() calls (use the per-step forward() API for that)")
    return _RunFn.apply(sim, steps, waveforms.contiguous(), checkpoint_every, *sim._mE64)


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
