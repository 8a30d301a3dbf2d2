"""x-slab decomposition of the FDTD path over the GPUs of one box (one process per GPU).

The reference has no parallelism at all (single-thread numpy); this is the new build's only
"parallel" axis.  The grid is cut into contiguous x-slabs (x is the slowest C-order axis, so a slab and
each halo plane are contiguous blocks: no packing).  The stencil reach is one plane:

* the H half-step (curl_E, derivatives.py:16-22: forward differences) needs D_y, D_z (and the static
  1/eps_y, 1/eps_z) at local i = nx   -> plane 0 of the RIGHT neighbour;
* the D half-step (curl_H, derivatives.py:24-30: backward differences) needs H_y, H_z at local i = -1
  -> plane nx-1 of the LEFT neighbour.

np.roll wraps, so the ranks form a ring (rank 0 <-> rank P-1).  Two halo transports:

* "peer" (default wherever the tensor-map kernels serve the slab: 3-D, Ny a multiple of 4, Nz a multiple of the
  16-byte vector and >= 32 vectors): every rank owns an exchange block in device memory, shared with its two
  neighbours by CUDA IPC; the CTAs that produce a slab's boundary plane store it straight into the neighbour's
  block (NVLink peer stores) and bump an arrival counter there, the CTAs that read a halo plane wait for its
  counter (include/ceviche_b200.h, "x-slab decomposition").  One H and one D launch per time step per rank, the
  time loop runs in C (cev_fdtd_run): no collective, no extra launch, no Python per step.
* "nccl" (any grid the path serves, incl. 2-D): per half-step the interior planes first (they need no halo),
  then wait for the halo posted at the end of the previous half-step, then the one boundary plane, then post
  this half-step's ncclSend/ncclRecv on a side stream.

Probe partial sums are reduced once per run().  Results are bit-identical to the single-GPU run (no arithmetic
is reordered).  `ceviche_b200.fdtd(eps_r, dL, npml, devices=[...])` returns a SlabFDTD: the same object surface
(run / prepare / forward / initialize_fields / fields / t_index / dt / grid_shape), one process per GPU.

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


# Relative extra cost of an x-plane inside the x-PML in the half-step kernels (all three components take the PML path and
# read-modify-write their integral there): measured on B200, scripts/slab_plane_cost.py (0.26 fp64 / 0.22 fp32).
XPML_PLANE_COST = 0.25


def partition(Nx, P, npml_x=0, alpha=0.0):
    """Contiguous x-ranges [(lo, hi)] of P slabs.  alpha = 0: sizes differing by at most one plane.  alpha > 0: planes
    inside the x-PML (the first and last npml_x planes) count 1 + alpha, and the slabs get equal COST -- the ring runs at
    the pace of its slowest rank, and the two end ranks would otherwise carry all of the x-PML work."""
    if alpha <= 0 or npml_x <= 0 or P < 3:
        base, extra = divmod(Nx, P)
        out, lo = [], 0
        for r in range(P):
            hi = lo + base + (1 if r < extra else 0)
            out.append((lo, hi))
            lo = hi
        return out
    cost = np.ones(Nx)
    cost[:min(npml_x, Nx)] += alpha
    cost[max(0, Nx - npml_x):] += alpha
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    cuts = [0]
    for r in range(1, P):
        target = cum[-1] * r / P
        c = int(np.argmin(np.abs(cum - target)))
        c = max(c, cuts[-1] + 2)                 # every slab keeps at least 2 planes
        cuts.append(min(c, Nx - 2 * (P - r)))
    cuts.append(Nx)
    return [(cuts[r], cuts[r + 1]) for r in range(P)]


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
        self.mE = [m.detach().to(device=device, dtype=dtype).contiguous() for m in inv_eps]
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

    def run_c(self, steps, wf):
        """The whole time loop in C (cev_fdtd_run): the plan's attached exchange blocks carry the halos."""
        st = self._st or self._state()
        _lib.check(self.plan.lib.cev_fdtd_run(self.plan.handle, C.byref(st), steps,
                                              wf.data_ptr() if self.n_sources and steps else None,
                                              self.partials.data_ptr() if self.n_slots and steps else None, self._s()))

    def series(self):
        from .fdtd import fold_probes
        if self.fold.shape[1] == 0:
            return torch.zeros((self.partials.shape[0], 0), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            return fold_probes(self.plan, self.partials, self.fold.shape[1], self._s())

    def field(self, key):
        c = "xyz".index(key[1])
        if key[0] == "E":
            return (self.mE[c].double() * self.D[c].double()).to(self.dtype)
        return (self.D if key[0] == "D" else self.H)[c]


class _LazyFields:
    """`fields` of a slab simulator: fields[key] gathers the full (Nx, Ny, Nz) array on every rank (a collective:
    all ranks must ask for the same keys in the same order).  `local_fields[key]` is this rank's slab."""

    def __init__(self, sim):
        self._sim = sim

    def __getitem__(self, key):
        return self._sim.gather(key)

    def keys(self):
        return ("Ex", "Ey", "Ez", "Dx", "Dy", "Dz", "Hx", "Hy", "Hz")

    def __iter__(self):
        return iter(self.keys())

    def items(self):
        return [(k, self[k]) for k in self.keys()]


class _LocalFields(_LazyFields):
    def __getitem__(self, key):
        loc = self._sim.be.field(key)
        return loc if torch.is_tensor(loc) else torch.as_tensor(loc)


class SlabFDTD:
    """FDTD on an x-slab per rank (one process per GPU): what `ceviche_b200.fdtd(eps_r, dL, npml, devices=[...])`
    returns.  Same surface as the single-GPU object for the caller loop: prepare / run / forward /
    initialize_fields / set_option / fields / t_index / dt / grid_shape / Nx, Ny, Nz.

    eps_r: the GLOBAL permittivity (every rank passes the same array; the slab and the extra plane the Yee
    averaging along x needs, utils.py:167, are cut here), or -- with `global_shape` given -- this rank's slab WITH
    one extra leading plane (global plane lo-1, periodic), shape (nx+1, Ny, Nz), for grids too large to
    materialise on every rank."""

    def __init__(self, global_shape, eps_local, dL, npml, *, dtype=torch.float64, device=None, group=None,
                 backend_factory=None, path=None, arith=None, balance=0.0, _ring=None):
        self.group = group
        if _ring is not None:                     # several slabs in one process (tests): (rank, world)
            self.rank, self.P = _ring
        else:
            self.P = dist.get_world_size(group) if dist.is_initialized() else 1
            self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._in_process = _ring is not None
        self.Nx, self.Ny, self.Nz = self.global_shape = self.grid_shape = tuple(int(v) for v in global_shape)
        self.N = self.Nx * self.Ny * self.Nz
        if self.Ny <= 1 and self.Nz <= 1:
            raise ValueError("slab decomposition needs a 2-D or 3-D grid (Ny > 1 or Nz > 1)")
        self.balance = float(balance)      # x-PML plane cost of the partition (0 = equal plane counts); see partition()
        self._parts = partition(self.Nx, self.P, int(npml[0]), self.balance)
        self.lo, self.hi = self._parts[self.rank]
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
        self._mE64 = inv_eps              # (carries the autograd graph back to eps_r when it requires grad)
        self.design_region = None         # ((x0, x1), (y0, y1), (z0, z1)) GLOBAL cells: see fdtd.design_region
        local_shape = (self.nx, self.Ny, self.Nz)
        if backend_factory is None:
            dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
            self.be = CudaSlabBackend(dev, dtype, local_shape, dL, self.dt, sH, sD, inv_eps,
                                      arith_f64=arith in ("f64", torch.float64))
        else:
            self.be = backend_factory(local_shape, dL, self.dt, sH, sD, [m.detach().cpu().numpy() for m in inv_eps])
        self.right = (self.rank + 1) % self.P
        self.left = (self.rank - 1) % self.P
        self._pending = []
        self.t_index = 0
        self.n_probes = 0
        self.fields = _LazyFields(self)
        self.local_fields = _LocalFields(self)
        self._blocks = None
        # ---- which halo transport
        if path not in (None, "peer", "nccl"):
            raise ValueError("path must be 'peer', 'nccl' or None (auto)")
        can_peer = self.be.is_cuda and self.P > 1 and self._peer_supported()
        if path == "peer" and not can_peer:
            raise ValueError("the peer-memory halo path needs CUDA slabs of a 3-D grid with Ny a multiple of 4 and Nz a "
                             "multiple of the 16-byte vector and >= 32 vectors")
        self.path = "none" if self.P == 1 else ("peer" if (can_peer and path != "nccl") else "nccl")
        self.be.halo = self.path == "nccl"
        if self.path == "nccl":
            # static halo: 1/eps_y, 1/eps_z at local i = nx  <- plane 0 of the right neighbour
            self._exchange([self.be.mE[1][0], self.be.mE[2][0]], self.left, [self.be.mE_hi[1], self.be.mE_hi[2]], self.right)
            self._wait()
        elif self.path == "peer" and not self._in_process:
            self._peer_setup_ipc()

    def __repr__(self):
        return "FDTD(eps_r.shape={}, dL={}, NPML={}, x-slab {} of {})".format(self.grid_shape, self.dL, self.npml,
                                                                               self.rank, self.P)

    def slab_path(self):
        return {"peer": "direct stores into the neighbour's peer-mapped halo block from inside the half-step kernels "
                        "(CUDA IPC over NVLink), arrival counters, time loop in C",
                "nccl": "ncclSend/ncclRecv of the halo planes per half-step on a side stream, boundary plane launched after the interior",
                "none": "single slab (periodic wrap inside the array)"}[self.path]

    def set_option(self, name, value):
        _lib.check(self.be.plan.lib.cev_fdtd_set_option(self.be.plan.handle, name.encode(), int(value)))

    # ---- peer-memory halo plumbing ----------------------------------------------------------------
    def _peer_supported(self):
        v = 2 if self.dtype == torch.float64 else 4
        return (self.Ny > 1 and self.Nz > 1 and self.Ny % 4 == 0 and self.Nz % v == 0 and self.Nz >= 32 * v)

    def _layout(self):
        lay = _lib.cev_halo_layout()
        _lib.check(self.be.plan.lib.cev_fdtd_halo_layout(self.be.plan.handle, C.byref(lay)))
        return lay

    def _peer_alloc(self, want_handle=True):
        lib = self.be.plan.lib
        block, handle = C.c_void_p(), C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
        _lib.check(lib.cev_halo_alloc(self.be.device.index, self._layout().bytes, C.byref(block), handle if want_handle else None))
        return block, handle.raw

    def _peer_attach(self, own, left, right):
        """own / left / right: device pointers of the three exchange blocks (left may equal right)."""
        be, lib = self.be, self.be.plan.lib
        self._blocks = (own, left, right)
        _lib.check(lib.cev_fdtd_halo_attach(be.plan.handle, own, left, right))
        with torch.cuda.device(be.device):
            st = be._state()
            _lib.check(lib.cev_fdtd_halo_push_static(be.plan.handle, C.byref(st), be._s()))
            torch.cuda.current_stream(be.device).synchronize()

    def _peer_setup_ipc(self):
        """One process per GPU: allocate this rank's block, trade IPC handles, map the neighbours' blocks."""
        lib, dev = self.be.plan.lib, self.be.device.index
        own, handle = self._peer_alloc()
        handles = [None] * self.P
        dist.all_gather_object(handles, handle, group=self.group)
        opened = {}
        for r in {self.left, self.right}:
            q = C.c_void_p()
            _lib.check(lib.cev_halo_open(dev, handles[r], C.byref(q)))
            opened[r] = q
        self._opened, self._own_block = opened, own
        self._peer_attach(own, opened[self.left], opened[self.right])
        dist.barrier(group=self.group)            # every neighbour's 1/eps halo has landed before the first step

    def _peer_check(self):
        err = C.c_int(0)
        _lib.check(self.be.plan.lib.cev_fdtd_halo_error(self.be.plan.handle, C.byref(err)))
        if err.value:
            raise _lib.CevicheB200Error("x-slab halo wait timed out: a neighbouring rank never delivered its boundary plane "
                                        "(all ranks must run the same sequence of steps)")

    def __del__(self):
        try:
            if self._blocks is not None and not self._in_process:
                lib, dev = self.be.plan.lib, self.be.device.index
                lib.cev_fdtd_halo_attach(self.be.plan.handle, None, None, None)
                for q in self._opened.values():
                    lib.cev_halo_close(dev, q)
                lib.cev_halo_free(dev, self._own_block)
                self._blocks = None
        except Exception:
            pass

    # ---- NCCL halo plumbing -------------------------------------------------------------------------
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
    def initialize_fields(self):
        """fdtd.py:147-211 on every slab: zero state (and halo planes), t_index = 0.  A collective."""
        be = self.be
        for t in be.H + be.D:
            t.zero_()
        if be.is_cuda:
            for fam in _PML_FAMILIES:
                for t in be.pml[fam]:
                    t.zero_()
        else:
            for arrs in be.I.values():
                for arr in arrs:
                    arr[...] = 0.0
        for t in [x for x in (be.D_hi + be.H_lo) if x is not None]:
            t.zero_()
        if self.path == "peer":
            if not self._in_process:
                torch.cuda.current_stream(be.device).synchronize()
                dist.barrier(group=self.group)        # no neighbour is still writing into this rank's block
            _lib.check(be.plan.lib.cev_fdtd_halo_reset(be.plan.handle, be._s()))
            if not self._in_process:
                torch.cuda.current_stream(be.device).synchronize()
                dist.barrier(group=self.group)
        self.t_index = 0

    def prepare(self, sources=(), probes=()):
        """sources [(comp, global profile)], probes [(field key, global mask)] -> local point sets.  Profiles / masks
        are dense (Nx, Ny, Nz) arrays, or point lists {"ijk", "w", "Ny", "Nz"} on grids too large for dense masks."""
        plane = self.Ny * self.Nz
        src = [(_COMP[s[0]],) + localize_points(s[1], self.lo, self.hi, plane) for s in sources]
        prb = [(_FIELD_CODE[k],) + localize_points(m, self.lo, self.hi, plane) for k, m in probes]
        self.be.set_points(src, prb)
        self.n_probes = len(probes)
        self.n_sources = len(sources)

    def run(self, steps, sources=None, probes=None, waveforms=None, checkpoint_every=None):
        """`steps` leap-frog steps (the loop of ceviche/utils.py:325-331 on every slab).  Same arguments as the
        single-GPU `fdtd.run`: sources [(component, profile[, waveform])] / probes [(field key, mask)] (None = keep
        the prepared ones), waveforms [steps, n_sources].  Returns the probe series [steps, n_probes] summed over
        ranks (identical on every rank).  A collective: every rank calls it with the same arguments."""
        if sources is not None and not isinstance(sources, (list, tuple)):      # run(steps, waveforms): the older call
            sources, waveforms = None, sources
        steps = int(steps)
        if sources is not None or probes is not None:
            if sources is None or probes is None:
                raise ValueError("run(): give sources and probes together (or prepare() them once)")
            self.prepare(sources, probes)
            if waveforms is None:
                waveforms = (np.stack([np.asarray(s_[2], dtype=np.float64)[:steps] for s_ in sources], axis=1)
                             if len(sources) else np.zeros((steps, 0)))
        be, nx, P = self.be, self.nx, self.P
        if waveforms is None:
            if getattr(self, "n_sources", 0):
                raise ValueError("run(): prepared sources need `waveforms` [steps, n_sources]")
            waveforms = np.zeros((steps, 0))
        wf = torch.as_tensor(np.ascontiguousarray(waveforms, dtype=np.float64)) if not torch.is_tensor(waveforms) else waveforms.double().contiguous()
        if be.is_cuda:
            wf = wf.to(be.device)
        if tuple(wf.shape) != (steps, getattr(self, "n_sources", wf.shape[1] if wf.dim() == 2 else 0)):
            raise ValueError("waveforms must have shape (steps, n_sources)")
        if be.is_cuda and torch.is_grad_enabled() and any(bool(m.requires_grad) for m in self._mE64):
            # eps_r requires grad: the run is an autograd node whose backward is the adjoint FDTD on the slabs
            if self.t_index != 0:
                raise RuntimeError("a differentiable run() must start from initialize_fields()")
            return _SlabRunFn.apply(self, steps, wf, checkpoint_every, *self._mE64)
        return self._run_plain(steps, wf)

    def _run_plain(self, steps, wf):
        be, nx, P = self.be, self.nx, self.P
        be.new_partials(steps)
        if self.path in ("peer", "none") and be.is_cuda:
            be.run_c(steps, wf)                   # the time loop in C; peer path: halos travel inside the kernels
        else:
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
        if P > 1 and not self._in_process:
            if self.path == "nccl":
                self._wait()     # leave no message in flight; the D halo is in place for a following run()
            if self.n_probes:
                # sum of the ranks' partial series in RANK ORDER (all-gather + a fixed-order sum): an all-reduce picks its
                # reduction order by message size and ring / tree algorithm, so a run split into two run() calls differed
                # from the unsplit one in the last bit on 4 ranks
                parts = [torch.empty_like(series) for _ in range(P)]
                dist.all_gather(parts, series.contiguous(), group=self.group)
                series = parts[0].clone()
                for r in range(1, P):
                    series += parts[r]
            if self.path == "peer":
                self._peer_check()
        return series

    def _local_grad_box(self):
        """design_region (global cells) -> this slab's part of the G_mE box as six local ints (x range clipped to the slab,
        possibly empty), or None = everywhere.  One more cell on the high side of every axis: eps_r[i, j, k] enters the
        Yee averages of cells (i..i+1, j..j+1, k..k+1) (utils.py:167-174)."""
        if self.design_region is None:
            return None
        (x0, x1), (y0, y1), (z0, z1) = [(int(a), int(b)) for a, b in self.design_region]
        if x1 + 1 > self.Nx or y1 + 1 > self.Ny or z1 + 1 > self.Nz:
            return None
        lx0, lx1 = max(x0, self.lo) - self.lo, min(x1 + 1, self.hi) - self.lo
        if lx1 <= lx0:
            return [1, 1, 0, 0, 0, 0]      # the box does not touch this slab: an empty x-range
        return [lx0, lx1, y0, y1 + 1, z0, z1 + 1]

    def _restore_halos(self, peer):
        """After the reverse sweep put the end-of-run state's halo planes back in place (a collective)."""
        be = self.be
        if peer:        # the exchange blocks were paused, not touched: they still hold the end-of-run planes
            self.set_option("halo_pause", 0)
        else:
            self._exchange([be.D[1][0], be.D[2][0]], self.left, [be.D_hi[1], be.D_hi[2]], self.right)
            self._wait()

    def forward(self, Jx=None, Jy=None, Jz=None):
        """One time step with dense GLOBAL J arrays (the reference's per-step call, fdtd.py:74-144), for API
        compatibility on small grids: J is re-uploaded as a sparse source set on every call (use run() for time
        loops).  Returns `fields` (gathers lazily)."""
        srcs = []
        for comp, J in zip("xyz", (Jx, Jy, Jz)):
            if J is None:
                continue
            J = J.detach().cpu().numpy() if torch.is_tensor(J) else np.asarray(J, dtype=np.float64)
            J = np.broadcast_to(J.reshape(J.shape + (1,) * (3 - J.ndim)) if J.ndim else J, self.global_shape)
            srcs.append((comp, J, np.ones(1)))
        keep = getattr(self, "_probe_spec", None)
        self.prepare([(c, p) for c, p, _ in srcs], [])
        self.run(1, waveforms=np.ones((1, len(srcs))))
        self._sources_stale = True
        del keep
        return self.fields

    def time_local_step_H(self, reps=20):
        """CUDA-event time (ms) of the H half-step over this rank's whole slab, halos paused (periodic wrap inside
        the slab: the same work, no neighbour involved)."""
        be = self.be
        if self.path == "peer":
            self.set_option("halo_pause", 1)
        be.new_partials(1)
        for _ in range(3):
            be.step_H(0, be.nx, -1)
        torch.cuda.synchronize(be.device)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            be.step_H(0, be.nx, -1)
        b.record()
        torch.cuda.synchronize(be.device)
        if self.path == "peer":
            self.set_option("halo_pause", 0)
        return a.elapsed_time(b) / reps

    def gather(self, key):
        """Full field on every rank (tests / small grids only)."""
        loc = self.be.field(key)
        loc = loc if torch.is_tensor(loc) else torch.as_tensor(loc)
        if self.P == 1 or self._in_process:
            return loc
        sizes = [hi - lo for lo, hi in self._parts]
        outs = [torch.empty((s, self.Ny, self.Nz), dtype=loc.dtype, device=loc.device) for s in sizes]
        dist.all_gather(outs, loc.contiguous(), group=self.group) if len(set(sizes)) == 1 else self._gather_uneven(outs, loc)
        return torch.cat(outs, 0)

    def _gather_uneven(self, outs, loc):
        for r, buf in enumerate(outs):
            if r == self.rank:
                buf.copy_(loc)
            dist.broadcast(buf, src=r, group=self.group)


class _AllReduceGrad(torch.autograd.Function):
    """Identity whose backward sums the gradient over the ranks: every rank differentiates its own slab's share of the
    objective w.r.t. the GLOBAL eps_r, and ends up with the complete gradient."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        if dist.is_initialized() and dist.get_world_size(ctx.group) > 1:
            if g.is_cuda or dist.get_backend(ctx.group) != "nccl":
                dist.all_reduce(g, group=ctx.group)
            else:                                   # NCCL groups reduce device tensors only
                dev = torch.device("cuda", torch.cuda.current_device())
                gd = g.to(dev)
                dist.all_reduce(gd, group=ctx.group)
                g = gd.to(g.device)
        return g, None


class _SlabRunFn(torch.autograd.Function):
    """run() on x-slabs as an autograd node: forward = the plain slab run with state checkpoints every K steps;
    backward = per segment, recompute the slab's forward steps keeping D per step, then the transposed steps in reverse
    order (cev_fdtd_adjoint_part: cell-local D part, H part, E part) with one plane pair exchanged between the parts --
    the +x plane of gC for the H part, the -x plane of gC2 for the E part, mirroring the forward half-steps' halos."""

    @staticmethod
    def forward(ctx, sim, steps, wf, every, mEx, mEy, mEz):
        be = sim.be
        every = max(1, int(every or np.ceil(np.sqrt(max(steps, 1)))))
        ctx.sim, ctx.steps, ctx.wf = sim, steps, wf
        ctx.checkpoints, chunks = [], []
        for t0 in range(0, steps, every):
            t1 = min(steps, t0 + every)
            ctx.checkpoints.append((t0, t1, [t.clone() for t in be.H], [t.clone() for t in be.D],
                                    {fam: [t.clone() for t in be.pml[fam]] for fam in _PML_FAMILIES}))
            chunks.append(sim._run_plain(t1 - t0, wf[t0:t1]))
        return torch.cat(chunks) if chunks else torch.zeros((0, sim.n_probes), dtype=torch.float64, device=be.device)

    @staticmethod
    def backward(ctx, gbar):
        sim, be = ctx.sim, ctx.sim.be
        lib, h = be.plan.lib, be.plan.handle
        shape, dev, dt_ = (be.nx, be.Ny, be.Nz), be.device, be.dtype
        P = sim.P
        with torch.cuda.device(dev):
            gbar = gbar.detach().to(torch.float64).contiguous()
            z3 = lambda dtype=dt_: [torch.zeros(shape, dtype=dtype, device=dev) for _ in range(3)]
            lH, lD, gC, gC2, G = z3(), z3(), z3(), z3(), z3(torch.float64)
            lp = {fam: [torch.zeros_like(t) for t in be.pml[fam]] for fam in _PML_FAMILIES}
            plane = (be.Ny, be.Nz)
            gC_hi = [None] + [torch.zeros(plane, dtype=dt_, device=dev) for _ in range(2)]
            gC2_lo = [None] + [torch.zeros(plane, dtype=dt_, device=dev) for _ in range(2)]
            p3 = lambda ts: _lib.c_void_p3(*[None if (t is None or t.numel() == 0) else t.data_ptr() for t in ts])
            adj = _lib.cev_adjoint()
            adj.lH, adj.lD, adj.gC, adj.gC2, adj.G_mE = p3(lH), p3(lD), p3(gC), p3(gC2), p3(G)
            for fam in _PML_FAMILIES:
                setattr(adj, "l" + fam, p3(lp[fam]))
            box = sim._local_grad_box()
            if box is not None:
                adj.g_box = (C.c_int64 * 6)(*box)
            skip_G = box is not None and box[0] >= box[1]          # the design box does not touch this slab

            def fwd_state(D):
                st = _lib.cev_state()
                st.D, st.inv_eps = p3(D), p3(be.mE)
                return st

            # the state at the end of the run comes back after the sweep (the recomputation runs in the slab's own arrays)
            end = ([t.clone() for t in be.H], [t.clone() for t in be.D],
                   {fam: [t.clone() for t in be.pml[fam]] for fam in _PML_FAMILIES}, sim.t_index)
            peer = sim.path == "peer"
            if peer:         # recompute with caller-owned halo planes (NCCL), the exchange blocks paused
                sim.set_option("halo_pause", 1)
                if not getattr(sim, "_mE_hi_filled", False):       # the static 1/eps halo lives in the blocks on this path
                    sim._exchange([be.mE[1][0], be.mE[2][0]], sim.left, [be.mE_hi[1], be.mE_hi[2]], sim.right)
                    sim._wait()
                    sim._mE_hi_filled = True
            be.halo = P > 1
            be._st = None
            try:
                for t0, t1, H0, D0, P0 in reversed(ctx.checkpoints):
                    for dst, src in zip(be.H + be.D, H0 + D0):
                        dst.copy_(src)
                    for fam in _PML_FAMILIES:
                        for dst, src in zip(be.pml[fam], P0[fam]):
                            dst.copy_(src)
                    n = t1 - t0
                    hist = [[t.clone() for t in be.D]]
                    be.new_partials(1)
                    be._st = None
                    if P > 1:     # D halo of the checkpoint time: plane 0 of the right neighbour
                        sim._exchange([be.D[1][0], be.D[2][0]], sim.left, [be.D_hi[1], be.D_hi[2]], sim.right)
                        sim._wait()
                    for k in range(n):
                        be.step_H(0, be.nx, -1)
                        if P > 1:
                            sim._exchange([be.H[1][be.nx - 1], be.H[2][be.nx - 1]], sim.right, [be.H_lo[1], be.H_lo[2]], sim.left)
                            sim._wait()
                        be.step_D(0, be.nx, -1, ctx.wf[t0 + k])
                        if P > 1:
                            sim._exchange([be.D[1][0], be.D[2][0]], sim.left, [be.D_hi[1], be.D_hi[2]], sim.right)
                            sim._wait()
                        hist.append([t.clone() for t in be.D])
                    s = be._s()
                    for k in range(n, 0, -1):
                        if sim.n_probes:
                            _lib.check(lib.cev_fdtd_adjoint_seed(h, C.byref(fwd_state(hist[k])), C.byref(adj), gbar[t0 + k - 1].data_ptr(), s))
                        st = fwd_state(hist[k - 1])
                        _lib.check(lib.cev_fdtd_adjoint_part(h, 0, C.byref(st), C.byref(adj), None, s))
                        halo = None
                        if P > 1:
                            sim._exchange([gC[1][0], gC[2][0]], sim.left, [gC_hi[1], gC_hi[2]], sim.right)
                            sim._wait()
                            halo = C.byref(p3(gC_hi))
                        _lib.check(lib.cev_fdtd_adjoint_part(h, 1, C.byref(st), C.byref(adj), halo, s))
                        if P > 1:
                            sim._exchange([gC2[1][be.nx - 1], gC2[2][be.nx - 1]], sim.right, [gC2_lo[1], gC2_lo[2]], sim.left)
                            sim._wait()
                            halo = C.byref(p3(gC2_lo))
                        _lib.check(lib.cev_fdtd_adjoint_part(h, 2, C.byref(st), C.byref(adj), halo, s))
                    del hist
            finally:
                for dst, src in zip(be.H + be.D, end[0] + end[1]):
                    dst.copy_(src)
                for fam in _PML_FAMILIES:
                    for dst, src in zip(be.pml[fam], end[2][fam]):
                        dst.copy_(src)
                be.halo = sim.path == "nccl"
                be._st = None
                if P > 1:      # the halo planes / exchange blocks of the end-of-run state, for a following run()
                    sim._restore_halos(peer)
            if skip_G:
                G = [torch.zeros_like(g) for g in G]
        return (None, None, None, None, *G)


def make_slab_fdtd(eps_r, dL, npml, *, devices, global_shape=None, dtype=torch.float64, device=None, arith=None,
                   group=None, path=None, balance=None):
    """What `ceviche_b200.fdtd(eps_r, dL, npml, devices=[...])` builds: one SlabFDTD per process.  `devices` lists the
    CUDA device of every slab in rank order (one process per GPU, torch.distributed initialised with that many ranks)."""
    P = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if len(devices) != P:
        raise ValueError("devices lists {} slabs but the process group has {} ranks (one process per GPU)".format(len(devices), P))
    if device is None:
        d = devices[rank]
        device = torch.device("cuda", d) if isinstance(d, int) else torch.device(d)
    # balance: x-PML plane cost of the partition (see partition()); None / 0 = equal plane counts, the default: the
    # cost-balanced cut (XPML_PLANE_COST) was measured neutral on 8 B200 (280.7 vs 277.3 Gcell/s fp64, 545.3 vs 545.9 fp32:
    # under the power cap the ring is paced by its slowest GPU, not by the x-PML planes of the end ranks)
    if balance is None:
        balance = 0.0
    if global_shape is None:
        e = torch.as_tensor(np.asarray(eps_r, dtype=np.float64)) if not torch.is_tensor(eps_r) else eps_r
        e = reshape_to_ND(e, 3)
        global_shape = tuple(e.shape)
        lo, hi = partition(global_shape[0], P, int(npml[0]), balance)[rank]
        idx = torch.arange(lo - 1, hi) % global_shape[0]
        if torch.is_tensor(e) and e.requires_grad:     # every rank ends up with the complete gradient w.r.t. eps_r
            e = _AllReduceGrad.apply(e, group)
        eps_r = e[idx.to(e.device)]
    return SlabFDTD(global_shape, eps_r, dL, npml, dtype=dtype, device=device, group=group, path=path, arith=arith,
                    balance=balance)
