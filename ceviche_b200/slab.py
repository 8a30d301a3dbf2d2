"""x-slab decomposition of the FDTD path over the GPUs of one box (one process per GPU).

The reference has no parallelism at all (single-thread numpy); this is the new build's only
"parallel" axis.  The grid is cut into contiguous x-slabs (x is the slowest C-order axis, so a slab and
each halo plane are contiguous blocks: no packing).  The stencil reach is one plane:

* the H half-step (curl_E, derivatives.py:16-22: forward differences) needs D_y, D_z (and the static
  1/eps_y, 1/eps_z) at local i = nx   -> plane 0 of the RIGHT neighbour;
* the D half-step (curl_H, derivatives.py:24-30: backward differences) needs H_y, H_z at local i = -1
  -> plane nx-1 of the LEFT neighbour.

np.roll wraps, so the ranks form a ring (rank 0 <-> rank P-1).  Schedule per half-step: interior planes
first (they need no halo), then wait for the halo posted at the end of the previous half-step, then the
one boundary plane, then post this half-step's send/recv (NCCL send/recv over NVLink, on a side stream):
every message has a whole interior half-step to arrive.  Probe partial sums are reduced once at the end.
Results are bit-identical to the single-GPU run (no arithmetic is reordered).

The per-slab compute goes through a small backend interface (`CudaSlabBackend` = the C ABI); tests inject
a numpy backend to exercise the partitioning / exchange / reduction logic on CPU with gloo.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .constants import C_0
from .fdtd import _FIELD_CODE, _COMP, _PML_FAMILIES, reshape_to_ND, sigma_profiles


def partition(Nx, P):
    """Contiguous x-ranges [(lo, hi)] of P slabs, sizes differing by at most one plane."""
    base, extra = divmod(Nx, P)
    out, lo = [], 0
    for r in range(P):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def localize_points(arr, lo, hi, plane):
    """Global profile/mask -> (local flat idx int64, weights float64) of the non-zeros inside x-planes
    [lo, hi).  `arr` is a dense (Nx,Ny,Nz) array, or -- for grids too large to hold dense masks -- a dict
    {"ijk": int [n,3] global cell coordinates, "w": float [n], "Ny": .., "Nz": ..}."""
    if isinstance(arr, dict):
        ijk, w = np.asarray(arr["ijk"], dtype=np.int64).reshape(-1, 3), np.asarray(arr["w"], dtype=np.float64)
        keep = (ijk[:, 0] >= lo) & (ijk[:, 0] < hi)
        loc = ((ijk[keep, 0] - lo) * arr["Ny"] + ijk[keep, 1]) * arr["Nz"] + ijk[keep, 2]
        order = np.argsort(loc, kind="stable")
        return loc[order], w[keep][order]
    a = np.asarray(arr, dtype=np.float64)
    a = a.reshape(a.shape + (1,) * (3 - a.ndim))
    sub = a[lo:hi].reshape(-1)
    idx = np.flatnonzero(sub)
    return idx.astype(np.int64), sub[idx]


class CudaSlabBackend:
    """One slab on one GPU behind the C ABI (include/ceviche_b200.h)."""
    is_cuda = True

    def __init__(self, device, dtype, shape_local, dL, dt, sH, sD, inv_eps, arith_f64=False):
        from .fdtd import _Plan
        self.device, self.dtype = device, dtype
        self.nx, self.Ny, self.Nz = shape_local
        self.plan = _Plan(device, dtype, arith_f64 or dtype == torch.float64, shape_local, dL, dt, sH, sD)
        z = lambda s: torch.zeros(s, dtype=dtype, device=device)
        self.H = [z(shape_local) for _ in range(3)]
        self.D = [z(shape_local) for _ in range(3)]
        self.mE = [m.to(device=device, dtype=dtype).contiguous() for m in inv_eps]
        shapes = self.plan.pml_shapes
        self.pml = {fam: [z(shapes[f * 3 + c]) for c in range(3)] for f, fam in enumerate(_PML_FAMILIES)}
        plane = (self.Ny, self.Nz)
        self.D_hi = [None, z(plane), z(plane)]       # D_y, D_z at local i = nx
        self.mE_hi = [None, z(plane), z(plane)]
        self.H_lo = [None, z(plane), z(plane)]       # H_y, H_z at local i = -1
        self.halo = False
        self.n_slots = 0
        self.main = torch.cuda.current_stream(device)
        # the halo messages ride on a high-priority side stream, so that their few NCCL CTAs are dispatched while the
        # interior half-step kernel still has thousands of CTAs pending (measured neutral at config-3 sizes on 2 GPUs:
        # the messages are already hidden; profiles/r1_slab_thin_slabs_2gpu.log)
        self.comm = torch.cuda.Stream(device, priority=-1)
        self._st = None

    def _state(self):
        st = _lib.cev_state()
        p3 = lambda ts: _lib.c_void_p3(*[None if (t is None or t.numel() == 0) else t.data_ptr() for t in ts])
        st.H, st.D, st.inv_eps = p3(self.H), p3(self.D), p3(self.mE)
        for fam in _PML_FAMILIES:
            setattr(st, fam, p3(self.pml[fam]))
        if self.halo:
            st.D_xhi, st.inv_eps_xhi, st.H_xlo = p3(self.D_hi), p3(self.mE_hi), p3(self.H_lo)
        return st

    def _s(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_points(self, sources, probes):
        """sources: [(comp, idx, w)], probes: [(field code, idx, w)] with LOCAL flat indices."""
        keep = []

        def mk(field, idx, w):
            pts = _lib.cev_points()
            ti = torch.as_tensor(idx).to(self.device)
            tw = torch.as_tensor(w).to(self.device)
            keep.extend([ti, tw])
            pts.field, pts.n, pts.cell0 = field, len(idx), 0
            pts.idx = ti.data_ptr() if len(idx) else None
            pts.weight = tw.data_ptr() if len(idx) else None
            return pts
        sp = (_lib.cev_points * max(1, len(sources)))()
        for q, (comp, idx, w) in enumerate(sources):
            sp[q] = mk(3 + comp, idx, w)
        pp = (_lib.cev_points * max(1, len(probes)))()
        for q, (field, idx, w) in enumerate(probes):
            pp[q] = mk(field, idx, w)
        n_slots = C.c_int64()
        with torch.cuda.device(self.device):
            torch.cuda.current_stream(self.device).synchronize()     # blocking legacy-stream copies follow
            _lib.check(self.plan.lib.cev_fdtd_set_sources(self.plan.handle, len(sources), sp))
            _lib.check(self.plan.lib.cev_fdtd_set_probes(self.plan.handle, len(probes), pp, C.byref(n_slots)))
        owner = (C.c_int32 * max(1, n_slots.value))()
        _lib.check(self.plan.lib.cev_fdtd_probe_slots(self.plan.handle, owner))
        fold = torch.zeros((n_slots.value, len(probes)), dtype=torch.float64)
        for s in range(n_slots.value):
            fold[s, owner[s]] = 1.0
        self.fold = fold.to(self.device)
        self.n_slots, self.n_sources = n_slots.value, len(sources)

    def new_partials(self, steps):
        self.partials = torch.zeros((steps, self.n_slots), dtype=torch.float64, device=self.device)
        self._st = self._state()          # the pointers do not change during a run(): built once, not per half-step

    def step_H(self, x0, x1, probe_t):
        if x1 <= x0:
            return
        st = self._st or self._state()
        _lib.check(self.plan.lib.cev_fdtd_step_H_ex(self.plan.handle, C.byref(st), None, None, x0, x1, probe_t,
                                                    self.partials.data_ptr() if self.n_slots else None, self._s()))

    def step_D(self, x0, x1, probe_t, wave_row):
        """D half-step of planes [x0, x1); the sources lying in those planes are injected in-kernel."""
        if x1 <= x0:
            return
        st = self._st or self._state()
        _lib.check(self.plan.lib.cev_fdtd_step_D_ex(self.plan.handle, C.byref(st), None, None, None, None,
                                                    wave_row.data_ptr() if self.n_sources else None, x0, x1, probe_t,
                                                    self.partials.data_ptr() if self.n_slots else None, self._s()))

    def sample(self, which, t):
        if self.n_slots == 0:
            return
        st = self._state()
        _lib.check(self.plan.lib.cev_fdtd_sample_probes(self.plan.handle, C.byref(st), None, which, t,
                                                        self.partials.data_ptr(), self._s()))

    def series(self):
        return self.partials @ self.fold

    def field(self, key):
        c = "xyz".index(key[1])
        if key[0] == "E":
            return (self.mE[c].double() * self.D[c].double()).to(self.dtype)
        return (self.D if key[0] == "D" else self.H)[c]


class SlabFDTD:
    """FDTD on an x-slab per rank.  `eps_local` is this rank's slab of eps_r WITH one extra leading
    plane (global plane lo-1, periodic) as needed by the Yee averaging along x (utils.py:167)."""

    def __init__(self, global_shape, eps_local, dL, npml, *, dtype=torch.float64, device=None, group=None,
                 backend_factory=None):
        self.group = group
        self.P = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.Nx, self.Ny, self.Nz = self.global_shape = tuple(global_shape)
        if self.Ny <= 1 and self.Nz <= 1:
            raise ValueError("slab decomposition needs a 2-D or 3-D grid (Ny > 1 or Nz > 1)")
        self.lo, self.hi = partition(self.Nx, self.P)[self.rank]
        self.nx = self.hi - self.lo
        if self.P > 1 and self.nx < 2:
            raise ValueError("each slab needs at least 2 x-planes")
        self.dL, self.npml, self.dtype = dL, list(npml), dtype
        self.dt = 0.5 * (1 / np.sqrt(3 / dL ** 2)) / C_0        # fdtd.py:219-222
        sH, sD = sigma_profiles(self.global_shape, self.npml, self.dt)
        sH = [sH[0][self.lo:self.hi].copy(), sH[1], sH[2]]
        sD = [sD[0][self.lo:self.hi].copy(), sD[1], sD[2]]
        e = torch.as_tensor(np.asarray(eps_local, dtype=np.float64)) if not torch.is_tensor(eps_local) else eps_local.double()
        if tuple(e.shape) != (self.nx + 1, self.Ny, self.Nz):
            raise ValueError("eps_local must have shape (nx+1, Ny, Nz) = {}".format((self.nx + 1, self.Ny, self.Nz)))
        if device is not None:
            e = e.to(device)
        own = e[1:]
        inv_eps = [1 / ((own + e[:-1]) / 2), 1 / ((own + torch.roll(own, 1, 1)) / 2), 1 / ((own + torch.roll(own, 1, 2)) / 2)]
        local_shape = (self.nx, self.Ny, self.Nz)
        if backend_factory is None:
            dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
            self.be = CudaSlabBackend(dev, dtype, local_shape, dL, self.dt, sH, sD, inv_eps)
        else:
            self.be = backend_factory(local_shape, dL, self.dt, sH, sD, [m.cpu().numpy() for m in inv_eps])
        self.right = (self.rank + 1) % self.P
        self.left = (self.rank - 1) % self.P
        self.be.halo = self.P > 1
        self._pending = []
        self.t_index = 0
        if self.P > 1:
            # static halo: 1/eps_y, 1/eps_z at local i = nx  <- plane 0 of the right neighbour
            self._exchange([self.be.mE[1][0], self.be.mE[2][0]], self.left, [self.be.mE_hi[1], self.be.mE_hi[2]], self.right)
            self._wait()

    # ---- halo plumbing ------------------------------------------------------------------------
    def _exchange(self, send_planes, send_to, recv_planes, recv_from):
        """Post send/recv of contiguous planes; on CUDA the NCCL calls are enqueued on the side stream after
        everything already queued on the compute stream, so the interior kernels launched next overlap."""
        ops = [dist.P2POp(dist.isend, t, send_to, self.group) for t in send_planes]
        ops += [dist.P2POp(dist.irecv, t, recv_from, self.group) for t in recv_planes]
        if self.be.is_cuda:
            ev = torch.cuda.Event()
            ev.record(self.be.main)
            with torch.cuda.stream(self.be.comm):
                self.be.comm.wait_event(ev)
                self._pending = dist.batch_isend_irecv(ops)
        else:
            self._pending = dist.batch_isend_irecv(ops)

    def _wait(self):
        for req in self._pending:
            req.wait()          # CUDA/NCCL: the compute stream waits; gloo: the host waits
        self._pending = []

    # ---- caller loop ----------------------------------------------------------------------------
    def prepare(self, sources=(), probes=()):
        """sources [(comp, global profile)], probes [(field key, global mask)] -> local point sets."""
        plane = self.Ny * self.Nz
        src = [(_COMP[c],) + localize_points(p, self.lo, self.hi, plane) for c, p in sources]
        prb = [(_FIELD_CODE[k],) + localize_points(m, self.lo, self.hi, plane) for k, m in probes]
        self.be.set_points(src, prb)
        self.n_probes = len(probes)

    def run(self, steps, waveforms):
        """`steps` leap-frog steps; waveforms [steps, n_sources].  Returns the probe series
        [steps, n_probes] summed over ranks (identical on every rank)."""
        be, nx, P = self.be, self.nx, self.P
        wf = torch.as_tensor(np.ascontiguousarray(waveforms, dtype=np.float64)) if not torch.is_tensor(waveforms) else waveforms.double().contiguous()
        if be.is_cuda:
            wf = wf.to(be.device)
        be.new_partials(steps)
        for n in range(steps):
            # ---- H half-step (fdtd.py:80-97): interior, then the plane that needs the right neighbour's D
            if P > 1:
                be.step_H(0, nx - 1, n - 1)
                self._wait()
                be.step_H(nx - 1, nx, -1)
                self._exchange([be.H[1][nx - 1], be.H[2][nx - 1]], self.right, [be.H_lo[1], be.H_lo[2]], self.left)
                # ---- D half-step (fdtd.py:105-127): interior, then the plane that needs the left neighbour's H
                be.step_D(1, nx, n, wf[n])
                self._wait()
                be.step_D(0, 1, -1, wf[n])
                self._exchange([be.D[1][0], be.D[2][0]], self.left, [be.D_hi[1], be.D_hi[2]], self.right)
            else:
                be.step_H(0, nx, n - 1)
                be.step_D(0, nx, n, wf[n])
        if steps > 0:
            be.sample(0, steps - 1)
        self.t_index += steps
        series = be.series()
        if P > 1:
            self._wait()     # leave no message in flight; the D halo is in place for a following run()
            self._pending_none = True
            if self.n_probes:
                dist.all_reduce(series, group=self.group)
        return series

    def gather(self, key):
        """Full field on every rank (tests / small grids only)."""
        loc = self.be.field(key)
        loc = loc if torch.is_tensor(loc) else torch.as_tensor(loc)
        if self.P == 1:
            return loc
        sizes = [hi - lo for lo, hi in partition(self.Nx, self.P)]
        outs = [torch.empty((s, self.Ny, self.Nz), dtype=loc.dtype, device=loc.device) for s in sizes]
        dist.all_gather(outs, loc.contiguous(), group=self.group) if len(set(sizes)) == 1 else self._gather_uneven(outs, loc)
        return torch.cat(outs, 0)

    def _gather_uneven(self, outs, loc):
        for r, buf in enumerate(outs):
            if r == self.rank:
                buf.copy_(loc)
            dist.broadcast(buf, src=r, group=self.group)
