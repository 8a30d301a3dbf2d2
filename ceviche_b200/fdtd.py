"""The `fdtd` simulator object: a drop-in for ceviche.fdtd (reference ceviche/fdtd.py:10-316)
whose arrays are torch CUDA tensors and whose time step runs in hand-written sm_100a kernels
behind the C ABI of include/ceviche_b200.h.

Same constructor, properties (with the reference's side effects), `forward(Jx, Jy, Jz)` returning
the nine-field dict, `initialize_fields()`, and attributes (`dt`, `t_index`, `Nx/Ny/Nz`,
`grid_shape`, `N`, `eps_arr`, `eps_xx/yy/zz`, `Hx..`, `Dx..`, `Ex..`).  Additions are keyword-only
(`dtype`, `device`, `arith`) plus the fused `run()` entry (the caller loop of
ceviche/utils.py:316-332 executed on the device with in-kernel sources and probes).

Host-side set-up (dt, the six 1-D sigma profiles) is fp64 numpy following the reference
formulas; the 30 full-grid coefficient arrays of fdtd.py:265-311 are never built.
"""
import ctypes as C
import functools

import numpy as np
import torch

from . import _lib
from .constants import C_0, EPSILON_0


def _base(t):
    from .autodiff import base
    return base(t)

FIELD_KEYS = ("Ex", "Ey", "Ez", "Dx", "Dy", "Dz", "Hx", "Hy", "Hz")
_FIELD_CODE = {k: i for i, k in enumerate(FIELD_KEYS)}
_COMP = {"x": 0, "y": 1, "z": 2, "Jx": 0, "Jy": 1, "Jz": 2, 0: 0, 1: 1, 2: 2}
_DTYPES = {torch.float64: 1, torch.float32: 0}
_PML_FAMILIES = ("ICE", "IH", "ICH", "ID")


def reshape_to_ND(arr, N):
    """ceviche/utils.py:206-214: trailing singleton axes up to N dims, ValueError beyond."""
    ND = len(arr.shape)
    if ND > N:
        raise ValueError("array is larger than {} dimensional, given shape {}".format(N, tuple(arr.shape)))
    return arr.reshape(tuple(arr.shape) + (N - ND) * (1,))


def sigma_profiles(shape, npml, dt):
    """The six 1-D PML profiles equivalent to fdtd._compute_sigmas (fdtd.py:224-263).

    Each reference sigma array varies along its own axis only; on the doubled grid of axis a
    (2*N_a samples) the cubic grading sits at indices 2p-n+1 (low side) and 2N_a-2p+n (high
    side) for n = 0..2p-1; H samples the odd entries, D the even ones."""
    sH, sD = [], []
    for n_cells, p in zip(shape, npml):
        p = int(p)
        s2 = np.zeros(2 * n_cells)
        for n in range(2 * p):
            val = (0.5 * EPSILON_0 / dt) * (n / 2 / p) ** 3
            s2[2 * p - n + 1] = val
            s2[2 * n_cells - 2 * p + n] = val
        sH.append(np.ascontiguousarray(s2[1::2]))
        sD.append(np.ascontiguousarray(s2[0::2]))
    return sH, sD


def coupled_components(shape, mask):
    """Closure of a 6-bit component mask (bits 0-2: D/E x,y,z; bits 3-5: H x,y,z) under the curl
    couplings of ceviche/derivatives.py:16-30.  A periodic difference along an axis of extent 1 is
    identically zero, so on 2-D / 1-D grids the step splits into uncoupled polarisations
    (Nz = 1: TM = {Ez, Hx, Hy}, TE = {Ex, Ey, Hz}); on a full 3-D grid any source reaches all six."""
    live = [n > 1 for n in shape]
    # H_c is driven by E_a through d/d(b) and by E_b through d/d(a), with (a, b) the other two axes
    # (and D_c by H_a, H_b the same way)
    while True:
        new = mask
        for c in range(3):
            a, b = (c + 1) % 3, (c + 2) % 3
            for src, dst in ((0, 3), (3, 0)):
                if ((mask >> (src + a)) & 1 and live[b]) or ((mask >> (src + b)) & 1 and live[a]):
                    new |= 1 << (dst + c)
        if new == mask:
            return mask
        mask = new


def _ptr(t):
    return None if t is None or t.numel() == 0 else t.data_ptr()


def _ptr3(ts):
    return _lib.c_void_p3(*[_ptr(t) for t in ts])


_ONES3 = _lib.c_double3(1.0, 1.0, 1.0)


class _Plan:
    """Owner of one C-side plan handle."""

    def __init__(self, device, dtype, arith_f64, shape, dL, dt, sH, sD):
        self.lib = _lib.load()
        self.handle = C.c_void_p()
        self._keep = [np.ascontiguousarray(a, dtype=np.float64) for a in list(sH) + list(sD)]
        for q, a in enumerate(self._keep):      # the C side reads shape[axis] entries of each profile
            if a.shape != (shape[q % 3],):
                raise ValueError("sigma profile {} has shape {}, expected ({},)".format(q, a.shape, shape[q % 3]))
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        cH = _lib.c_dptr3(*[dp(a) for a in self._keep[:3]])
        cD = _lib.c_dptr3(*[dp(a) for a in self._keep[3:]])
        _lib.check(self.lib.cev_fdtd_create(C.byref(self.handle), device.index, _DTYPES[dtype], int(arith_f64),
                                            shape[0], shape[1], shape[2], float(dL), float(dt), cH, cD))
        shapes = (C.c_int64 * 3 * 12)()
        _lib.check(self.lib.cev_fdtd_pml_shapes(self.handle, C.byref(shapes)))
        self.pml_shapes = [tuple(int(v) for v in shapes[q]) for q in range(12)]

    def __del__(self):
        try:
            if self.handle:
                self.lib.cev_fdtd_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass


def _host(method):
    """Host-side bookkeeping of the object runs below any active torch.func transform (autodiff.plain): its tensors go to
    the C ABI by pointer."""
    @functools.wraps(method)
    def wrapped(self, *args, **kwargs):
        if torch._C._are_functorch_transforms_active():
            from .autodiff import plain
            with plain():
                return method(self, *args, **kwargs)
        return method(self, *args, **kwargs)
    return wrapped


def fold_probes(plan, partials, n_probes, stream):
    """Slot partial sums [..., n_slots] -> probe series [..., n_probes], summed in slot order on the device
    (cev_fdtd_fold_probes: the same bits whatever the number of rows; a matmul with a 0/1 matrix is not)."""
    out = torch.empty(partials.shape[:-1] + (n_probes,), dtype=torch.float64, device=partials.device)
    rows = out.numel() // max(1, n_probes)
    if rows == 0 or n_probes == 0:
        return out
    _lib.check(plan.lib.cev_fdtd_fold_probes(plan.handle, _ptr(partials), rows, _ptr(out), stream))
    return out


class fdtd:

    def __new__(cls, eps_r=None, dL=None, npml=None, *, devices=None, global_shape=None, **kw):
        """`devices=[...]` (more than one): the grid is cut into x-slabs over the GPUs of one box, one process per GPU
        (ceviche_b200/slab.py); the object returned has the same caller-loop surface."""
        if devices is not None and len(devices) > 1:
            from .slab import make_slab_fdtd
            return make_slab_fdtd(eps_r, dL, npml, devices=devices, global_shape=global_shape, **kw)
        return super().__new__(cls)

    def __init__(self, eps_r, dL, npml, *, dtype=torch.float64, device=None, arith=None, devices=None, global_shape=None):
        """ Makes an FDTD object (signature of ceviche/fdtd.py:12)
                eps_r: relative permittivity, 1-/2-/3-D numpy array or torch tensor
                dL: the grid size (scalar, as the reference: fdtd.py:219)
                npml: the number of PML cells on each axis (3 ints, 0 = periodic axis)
            keyword-only: dtype (torch.float64 | torch.float32 storage), device, arith
            ('f64' | 'f32': arithmetic of the fp32 path; fp64 storage always computes in fp64).
        """
        if dtype not in _DTYPES:
            raise ValueError("dtype must be torch.float64 or torch.float32")
        if devices is not None and len(devices) == 1 and device is None:
            device = devices[0] if not isinstance(devices[0], int) else torch.device("cuda", devices[0])
        if global_shape is not None:
            raise ValueError("global_shape is only meaningful with several `devices` (x-slabs)")
        if not torch.cuda.is_available():
            raise _lib.CevicheB200Error("ceviche_b200 needs a CUDA device: there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise _lib.CevicheB200Error("ceviche_b200 runs on CUDA devices only")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.dtype = dtype
        # fp32 storage computes in fp32 by default (meets the 1e-5 bar at <= 1000 steps)
        self.arith_f64 = True if dtype == torch.float64 else (arith in ("f64", torch.float64))
        self._options = {}
        self._plan = None
        self._fused_step = True
        # skip field components that are provably zero (2-D TM / TE, 1-D): see coupled_components()
        self.specialise_components = True
        # ((x0, x1), (y0, y1), (z0, z1)) or None: reverse-mode gradients w.r.t. eps_r are only wanted inside this box
        # (the region being optimised); outside it the adjoint sweep skips the dL/d(1/eps) accumulation and returns 0
        self.design_region = None
        # without a design region: record D of the WHOLE grid per step when that fits in a quarter of the free memory (no
        # checkpoints, no recomputation in the reverse sweep; the forward run pays one extra copy of D per step)
        self.record_whole_grid = True

        eps_r = self._as_eps(eps_r, pad=True)
        self.Nx, self.Ny, self.Nz = self.grid_shape = tuple(eps_r.shape)

        self.dL = dL
        self.npml = npml
        self.eps_r = eps_r

    def __repr__(self):
        return "FDTD(eps_r.shape={}, dL={}, NPML={})".format(self.grid_shape, self.dL, self.npml)

    def __str__(self):
        return "FDTD object:\n\tdomain size = {}\n\tdL = {}\n\tNPML = {}".format(self.grid_shape, self.dL, self.npml)

    # ------------------------------------------------------------------ properties
    @property
    def dL(self):
        return self.__dL

    @dL.setter
    def dL(self, new_dL):
        """Resets the time step (only: sigma is NOT recomputed -- the reference's behaviour, fdtd.py:41-45)."""
        self.__dL = new_dL
        self._set_time_step()
        self._plan = None

    @property
    def npml(self):
        return self.__npml

    @npml.setter
    def npml(self, new_npml):
        self.__npml = new_npml
        self._compute_sigmas()
        self._plan = None

    @property
    def eps_r(self):
        return self.__eps_r

    @eps_r.setter
    def eps_r(self, new_eps):
        """New permittivity => new Yee averages and 1/eps, and a FIELD RESET (fdtd.py:63-72)."""
        new_eps = self._as_eps(new_eps, pad=False)
        if tuple(new_eps.shape) != tuple(self.grid_shape):
            # the reference fails here too: its sigma arrays keep the old shape and _compute_update_parameters
            # (fdtd.py:265-311) raises numpy's broadcast ValueError
            raise ValueError("eps_r of shape {} assigned to an FDTD object of grid shape {}: operands could not be "
                             "broadcast together (make a new fdtd object)".format(tuple(new_eps.shape), tuple(self.grid_shape)))
        self.__eps_r = new_eps
        e64 = new_eps.to(torch.float64)
        # ceviche/utils.py:153-176 (grid_center_to_xyz): mean with the previous cell, periodic
        self.eps_xx, self.eps_yy, self.eps_zz = [(e64 + torch.roll(e64, 1, a)) / 2 for a in range(3)]
        self.eps_arr = new_eps.flatten()
        self.N = self.eps_arr.numel()
        self.grid_shape = self.Nx, self.Ny, self.Nz = tuple(new_eps.shape)
        self._compute_update_parameters()
        self.initialize_fields()

    def _as_eps(self, eps, pad):
        if not torch.is_tensor(eps):
            eps = torch.as_tensor(np.asarray(eps, dtype=np.float64))
        if pad:
            eps = reshape_to_ND(eps, N=3)
        elif eps.dim() != 3:
            raise ValueError("eps_r must be 3-dimensional when assigned, given shape {}".format(tuple(eps.shape)))
        return eps.to(self.device)

    # ------------------------------------------------------------------ set-up
    def _set_time_step(self, stability_factor=0.5):
        """fdtd.py:213-222: always the 3-D Courant bound."""
        dL_sum = 3 / self.dL ** 2
        dL_avg = 1 / np.sqrt(dL_sum)
        courant_stability = dL_avg / C_0
        self.dt = courant_stability * stability_factor

    def _compute_sigmas(self):
        npml = list(self.npml)
        if len(npml) != 3:
            raise IndexError("npml needs 3 entries (fdtd.py:249 indexes npml[2])")
        self.sigH, self.sigD = sigma_profiles(self.grid_shape, npml, self.dt)

    def _compute_update_parameters(self, mu_r=1.0):
        """Only the D->E coefficients are arrays (fdtd.py:314-316); everything else lives in the
        plan's 1-D tables."""
        from .autodiff import base, plain
        self._mE64 = [1 / e for e in (self.eps_xx, self.eps_yy, self.eps_zz)]
        # what the kernels read: plain tensors (under torch.func transforms _mE64 is wrapped; see autodiff.base / plain)
        with plain():
            self._mE = [base(m).to(self.dtype).contiguous() for m in self._mE64]
        self.mEx1, self.mEy1, self.mEz1 = self._mE
        # the reference's m*1..4 arrays freeze dt at THIS call (a later `F.dL = ...` changes dt and the curls,
        # fdtd.py:41-45, 80, but not the coefficients until eps_r is assigned again)
        if self.__dict__.get("_coef_dt") != self.dt:
            self._plan = None
        self._coef_dt = self.dt

    @_host
    def _ensure_plan(self):
        if self._plan is None:
            self._plan = _Plan(self.device, self.dtype, self.arith_f64, self.grid_shape, self.dL, self._coef_dt,
                               self.sigH, self.sigD)
            self._alloc_pml()
            for name, value in self._options.items():
                _lib.check(self._plan.lib.cev_fdtd_set_option(self._plan.handle, name.encode(), int(value)))
            self._n_sources = self._n_probes = self._n_slots = 0
            self._n_mon_pts, self._mon_counts, self._mon_acc, self._mon_freqs, self.monitor_points = 0, [], None, None, []
            self._source_mask = 0
            self._slot_fold = None
        return self._plan

    def set_option(self, name, value):
        """Kernel tuning / test knobs of the C ABI ('kernel_variant', 'xchunk'); results do not change."""
        if name == "fused_step":        # host-side switch: 0 = run() never uses the fused full-step kernel
            self._fused_step = bool(value)
            return
        self._options[name] = int(value)
        if self._plan is not None:
            _lib.check(self._plan.lib.cev_fdtd_set_option(self._plan.handle, name.encode(), int(value)))

    def _drive(self, d_mask):
        """Record that the D components in `d_mask` (bits 0-2) are being driven and tell the plan which
        components can be non-zero from now on (the in-place kernels skip the others entirely)."""
        self._active = coupled_components(self.grid_shape, self._active | int(d_mask))
        self._apply_active(self._active)

    def _apply_active(self, mask):
        mask = int(mask) if self.specialise_components else 63
        if self._options.get("active_components", 63) != mask:
            self.set_option("active_components", mask)

    def _alloc_pml(self):
        shapes = self._plan.pml_shapes
        z = lambda s: torch.zeros(s, dtype=self.dtype, device=self.device)
        self._pml = {fam: [z(shapes[f * 3 + c]) for c in range(3)] for f, fam in enumerate(_PML_FAMILIES)}
        self._shadow = None
        self._st_cache = None

    @_host
    def initialize_fields(self):
        """fdtd.py:147-211: zero state, t_index = 0, a NEW fields dict."""
        self.t_index = 0
        self._st_cache = None
        self._ptr_cache = None      # (H list, D list, their pointer structs) as stored by the last plain forward()
        # could the state / 1/eps carry an autograd graph?  (eps_r that requires grad; set again by differentiable steps)
        self._grad_state = any(bool(m.requires_grad) for m in self.__dict__.get("_mE64", ()))
        if self.__dict__.get("_mon_acc") is not None:
            self._mon_acc.zero_()
        self._shadow = None         # ping-pong scratch of the fused kernel (allocated on first use)
        self._active = 0            # components that may be non-zero (bits 0-2 D/E, 3-5 H)
        z = lambda: [torch.zeros(self.grid_shape, dtype=self.dtype, device=self.device) for _ in range(3)]
        self._H, self._D, self._E = z(), z(), z()
        if self._plan is not None:
            self._alloc_pml()
        else:
            self._pml = None
        self.fields = {}
        self._publish()

    def _publish(self):
        self._published = True   # these tensors are now in the caller's hands: never mutate them
        for c, n in enumerate("xyz"):
            self.fields["E" + n] = self._E[c]
            self.fields["D" + n] = self._D[c]
            self.fields["H" + n] = self._H[c]

    def __getattr__(self, name):
        # Hx.., Dx.., Ex.., ICEx.. attribute surface of the reference (fdtd.py:157-199)
        if len(name) >= 2 and name[-1] in "xyz" and not name.startswith("_"):
            c = "xyz".index(name[-1])
            fam = name[:-1]
            d = self.__dict__
            if fam in ("H", "D", "E") and "_" + fam in d:
                return d["_" + fam][c]
            if fam in _PML_FAMILIES and d.get("_pml") is not None:
                return d["_pml"][fam][c]
            if fam in ("sigH", "sigD") and fam in d:
                view = [1, 1, 1]
                view[c] = -1
                return np.broadcast_to(d[fam][c].reshape(view), d["grid_shape"])
        raise AttributeError(name)

    # ------------------------------------------------------------------ C-ABI state
    def _state(self, H=None, D=None, mE=None):
        st = _lib.cev_state()
        st.H = _ptr3(H or self._H)
        st.D = _ptr3(D or self._D)
        st.inv_eps = _ptr3(mE or self._mE)
        for fam in _PML_FAMILIES:
            setattr(st, fam, _ptr3(self._pml[fam]))
        return st

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _sync_for_upload(self):
        """The point-set uploads of the C ABI (cev_fdtd_set_sources / _probes / _monitors) are blocking copies on the
        legacy default stream, which does not wait for torch's non-blocking side streams: finish the work that
        produced the index / weight tensors first."""
        torch.cuda.current_stream(self.device).synchronize()

    def _as_J(self, J):
        if J is None:
            return None
        if (torch.is_tensor(J) and J.dtype == self.dtype and J.device == self.device and tuple(J.shape) == self.grid_shape
                and J.is_contiguous()):
            return J                      # the common case of the reference-style loop: nothing to convert
        if not torch.is_tensor(J):
            J = torch.as_tensor(np.asarray(J))
        if J.is_complex():
            # the reference raises here too: `self.Dx += Jx` (fdtd.py:125-127) adds in place into a float64 array, which
            # numpy refuses for a complex right-hand side (UFuncTypeError, a TypeError)
            raise TypeError("Cannot cast ufunc 'add' output from dtype('complex128') to dtype('{}') with casting rule "
                            "'same_kind' (complex J: the reference's in-place D += J raises the same)".format(
                                "float64" if self.dtype == torch.float64 else "float32"))
        J = J.to(device=self.device, dtype=self.dtype)
        if J.dim() < 3:
            J = J.reshape(tuple(J.shape) + (1,) * (3 - J.dim())) if J.dim() > 0 else J
        return J.expand(self.grid_shape).contiguous()

    # ------------------------------------------------------------------ one time step
    def forward(self, Jx=None, Jy=None, Jz=None):
        """ one time step of FDTD (fdtd.py:74-144).  Returns the (same) `fields` dict holding nine
        FRESH tensors, like the reference hands out fresh arrays each step. """
        from . import autodiff
        plan = self._plan or self._ensure_plan()
        self.t_index += 1
        J = [self._as_J(j) for j in (Jx, Jy, Jz)]
        d_mask = (J[0] is not None) | ((J[1] is not None) << 1) | ((J[2] is not None) << 2)
        if d_mask & ~self._active:
            self._drive(d_mask)
        # the reference-style loop calls this thousands of times: the host work per call is kept small (the cheap
        # pre-check below decides whether the full differentiability test is needed at all)
        if (self._grad_state or autodiff.fwAD._current_level >= 0 or any(j is not None and j.requires_grad for j in J)) \
                and autodiff.needs_grad(self, J):
            self._H, self._D, self._E, self._pml = autodiff.step(self, J)
            self._grad_state = True
            self._st_cache = None
            self._publish()
            return self.fields
        if torch._C._are_functorch_transforms_active():     # nothing to differentiate in this step: run it below the transform
            with autodiff.plain():
                self._H, self._D = [_base(t) for t in self._H], [_base(t) for t in self._D]
                if self._pml is not None:
                    self._pml = {fam: [_base(t) for t in ts] for fam, ts in self._pml.items()}
                return self._step_plain(plan, [None if j is None else _base(j) for j in J])
        return self._step_plain(plan, J)

    def _step_plain(self, plan, J):
        # (no torch.cuda.device() context: every tensor names its device, the stream is asked for by device, and the C ABI
        # guards the device itself.)  The host work per call is what this path costs on small grids, so: one allocation
        # for the nine fresh arrays with their pointers by arithmetic, and the pointers of the state this very function
        # stored last time reused (`_ptr_cache` is keyed by the identity of the lists it stored).
        block = torch.empty((9,) + self.grid_shape, dtype=self.dtype, device=self.device)
        new = block.unbind(0)                                           # nine fresh arrays
        Hn, Dn, En = list(new[0:3]), list(new[3:6]), list(new[6:9])
        if self.N:
            base, n = block.data_ptr(), self.N * block.element_size()
            pH, pD, pE = (_lib.c_void_p3(base + (3 * f) * n, base + (3 * f + 1) * n, base + (3 * f + 2) * n) for f in range(3))
        else:
            pH = pD = pE = _lib.c_void_p3(None, None, None)
        lib, h, s = plan.lib, plan.handle, self._stream()
        st = self._st_cache
        if st is None:                  # 1/eps and PML pointers only change with eps_r / a reset
            st = self._st_cache = self._state()
        cache = self._ptr_cache
        if cache is not None and cache[0] is self._H and cache[1] is self._D:
            st.H, st.D = cache[2], cache[3]
        else:
            st.H, st.D = _ptr3(self._H), _ptr3(self._D)
        _lib.check(lib.cev_fdtd_step_H(h, C.byref(st), pH, 0, self.Nx, s))
        st.H = pH
        _lib.check(lib.cev_fdtd_step_D(h, C.byref(st), pD, pE, _ptr3(J), _ONES3, 0, self.Nx, s))
        self._H, self._D, self._E = Hn, Dn, En
        self._ptr_cache = (Hn, Dn, pH, pD)
        self._publish()
        return self.fields

    # ------------------------------------------------------------------ fused caller loop
    def _point_set(self, field_code, arr, keep, dense_ok):
        """profile / mask array -> cev_points (+ tensors kept alive in `keep`)."""
        if not torch.is_tensor(arr):
            arr = torch.as_tensor(np.asarray(arr, dtype=np.float64))
        arr = _base(arr).to(device=self.device, dtype=torch.float64)
        arr = reshape_to_ND(arr, 3).expand(self.grid_shape).reshape(-1)
        pts = _lib.cev_points()
        pts.field = field_code
        nz = torch.nonzero(arr).reshape(-1)
        if dense_ok and nz.numel() * 2 > arr.numel():
            w = arr.contiguous()
            pts.n, pts.idx, pts.cell0, pts.weight = w.numel(), None, 0, w.data_ptr()
            keep.append(w)
        else:
            w = arr[nz].contiguous()
            pts.n, pts.idx, pts.cell0, pts.weight = nz.numel(), _ptr(nz), 0, _ptr(w)
            keep += [nz, w]
        return pts

    @_host
    def set_sources(self, sources):
        """sources: [(component 'x'|'y'|'z', profile array)].  J(t) = sum_s profile_s * waveform[t, s]."""
        plan = self._ensure_plan()
        keep = []
        pts = (_lib.cev_points * max(1, len(sources)))()
        for s, (comp, profile) in enumerate(sources):
            pts[s] = self._point_set(3 + _COMP[comp], profile, keep, dense_ok=False)
        with torch.cuda.device(self.device):
            self._sync_for_upload()
            _lib.check(plan.lib.cev_fdtd_set_sources(plan.handle, len(sources), pts))
        self._n_sources = len(sources)
        self._source_mask = sum({1 << _COMP[comp] for comp, _ in sources})

    @_host
    def set_probes(self, probes):
        """probes: [(field key 'Ex'..'Hz', mask array)].  series[t, p] = sum(field_p * mask_p)."""
        plan = self._ensure_plan()
        keep = []
        pts = (_lib.cev_points * max(1, len(probes)))()
        for p, (key, mask) in enumerate(probes):
            pts[p] = self._point_set(_FIELD_CODE[key], mask, keep, dense_ok=True)
        n_slots = C.c_int64()
        with torch.cuda.device(self.device):
            self._sync_for_upload()
            _lib.check(plan.lib.cev_fdtd_set_probes(plan.handle, len(probes), pts, C.byref(n_slots)))
        owner = (C.c_int32 * max(1, n_slots.value))()
        _lib.check(plan.lib.cev_fdtd_probe_slots(plan.handle, owner))
        fold = torch.zeros((n_slots.value, len(probes)), dtype=torch.float64)
        for s in range(n_slots.value):
            fold[s, owner[s]] = 1.0
        self._slot_fold = fold.to(self.device)
        self._n_probes = len(probes)
        self._n_slots = n_slots.value

    @_host
    def set_monitors(self, monitors, freqs):
        """Running-DFT monitors: [(field key 'Ex'..'Hz', mask array)] (the non-zero cells of the mask are monitored)
        at the frequencies `freqs` (Hz).  While run() advances, F_m(f)[q] = sum_n field_q(n) exp(-2 pi i f n dt)
        is accumulated on the device (n = time-step index since initialize_fields(), i.e. the convention of
        np.fft.fft over the stored series, ceviche/utils.py:383-386, without storing the series).
        Read with monitor_values(); `monitor_points[m]` are the flat cell indices of monitor m."""
        plan = self._ensure_plan()
        freqs = np.atleast_1d(np.asarray(freqs, dtype=np.float64))
        keep = []
        pts = (_lib.cev_points * max(1, len(monitors)))()
        self.monitor_points, self._mon_counts = [], []
        for m, (key, mask) in enumerate(monitors):
            if not torch.is_tensor(mask):
                mask = torch.as_tensor(np.asarray(mask))
            mask = reshape_to_ND(_base(mask).to(self.device), 3).expand(self.grid_shape).reshape(-1)
            nz = torch.nonzero(mask).reshape(-1).contiguous()
            pts[m].field, pts[m].n, pts[m].idx, pts[m].cell0, pts[m].weight = _FIELD_CODE[key], nz.numel(), _ptr(nz), 0, None
            keep.append(nz)
            self.monitor_points.append(nz)
            self._mon_counts.append(int(nz.numel()))
        n_pts = C.c_int64()
        with torch.cuda.device(self.device):
            self._sync_for_upload()
            _lib.check(plan.lib.cev_fdtd_set_monitors(plan.handle, len(monitors), pts, len(freqs), C.byref(n_pts)))
        self._n_mon_pts = n_pts.value
        self._mon_freqs = freqs
        self._mon_acc = torch.zeros((self._n_mon_pts, len(freqs), 2), dtype=torch.float64, device=self.device)

    @_host
    def reset_monitors(self):
        if self._mon_acc is not None:
            self._mon_acc.zero_()

    def monitor_values(self):
        """[complex128 tensor (n_freq, n_points_m) per monitor]: the accumulated DFT sums."""
        if self._mon_acc is None:
            return []
        z = torch.view_as_complex(self._mon_acc)           # (points, freqs)
        out, off = [], 0
        for n in self._mon_counts:
            out.append(z[off:off + n].transpose(0, 1).contiguous())
            off += n
        return out

    def prepare(self, sources=(), probes=()):
        """Upload source profiles [(component, profile)] and probe masks [(field key, mask)] once;
        later `run(steps, waveforms=...)` calls with sources=None reuse them."""
        self.set_sources([(s[0], s[1]) for s in sources])
        self.set_probes(list(probes))

    @_host
    def _prepare_run(self, steps, sources, probes, waveforms):
        self._ensure_plan()
        if sources is None and probes is None and waveforms is None and self._n_sources == 0:
            sources = ()
        if sources is None:
            if waveforms is None:
                raise ValueError("run(): prepared sources need `waveforms` [steps, n_sources]")
            n_src = self._n_sources
        else:
            n_src = len(sources)
            self.set_sources([(s[0], s[1]) for s in sources])
        if probes is not None:
            self.set_probes(list(probes))
        elif self._slot_fold is None:
            self.set_probes([])
        if waveforms is None:
            if len(sources):
                waveforms = np.stack([np.asarray(s[2], dtype=np.float64)[:steps] for s in sources], axis=1)
            else:
                waveforms = np.zeros((steps, 0))
        if not torch.is_tensor(waveforms):
            waveforms = torch.as_tensor(np.ascontiguousarray(waveforms, dtype=np.float64))
        waveforms = _base(waveforms).to(device=self.device, dtype=torch.float64).contiguous()
        if tuple(waveforms.shape) != (steps, n_src):
            raise ValueError("waveforms must have shape (steps, n_sources) = {}".format((steps, n_src)))
        self._drive(self._source_mask)
        return waveforms

    def run(self, steps, sources=None, probes=None, waveforms=None, checkpoint_every=None, monitors=None, freqs=None):
        """`steps` fused time steps: the loop of ceviche/utils.py:325-331 on the device.

        sources: [(component, profile, waveform[steps])]  (or (component, profile) with
                 `waveforms` a [steps, n_sources] array/tensor); None = keep the prepared ones
        probes:  [(field key, mask)]; None = keep the prepared ones
        Returns series[steps, n_probes] (float64 tensor on the device).  State advances in place;
        `fields` is refreshed at the end.  If eps_r requires grad the series is differentiable
        (checkpointed adjoint FDTD, snapshots every `checkpoint_every` steps, default sqrt(steps))."""
        from . import autodiff
        steps = int(steps)
        waveforms = self._prepare_run(steps, sources, probes, waveforms)
        if monitors is not None:
            self.set_monitors(monitors, freqs)
        if autodiff.needs_grad(self, []):
            return autodiff.run(self, steps, waveforms, checkpoint_every)
        return self._run_raw(steps, waveforms)

    def jvp_run(self, steps, eps_tangents, sources=None, probes=None, waveforms=None):
        """Forward mode: the primal run plus a batch of tangents d/d(eps_r) along `eps_tangents`
        [B, Nx, Ny, Nz] in ONE sweep.  Returns (series [steps, P], dseries [B, steps, P])."""
        from . import autodiff
        steps = int(steps)
        waveforms = self._prepare_run(steps, sources, probes, waveforms)
        return autodiff.jvp_run(self, steps, waveforms, eps_tangents)

    @_host
    def _run_raw(self, steps, waveforms, refresh=True, fused=True):
        plan = self._ensure_plan()
        with torch.cuda.device(self.device):
            if self._published:   # the in-place loop must not touch tensors handed out earlier
                self._H = [t.clone() for t in self._H]
                self._D = [t.clone() for t in self._D]
                self._published = False
            partials = torch.zeros((steps, self._n_slots), dtype=torch.float64, device=self.device)
            st = self._state()
            monitoring = self._n_mon_pts > 0 and steps > 0
            if monitoring:      # phasors of this leg: exp(-2 pi i f n dt), n counted from initialize_fields()
                n = torch.arange(self.t_index, self.t_index + steps, dtype=torch.float64, device=self.device)
                ang = (-2 * np.pi * self.dt) * n[:, None] * torch.as_tensor(self._mon_freqs, device=self.device)[None, :]
                phasors = torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).contiguous()
                _lib.check(plan.lib.cev_fdtd_bind_monitors(plan.handle, _ptr(phasors), _ptr(self._mon_acc)))
            if fused and self._use_fused(steps) and not monitoring:
                # one kernel per time step (H and D half-steps fused): needs ping-pong scratch for H, D, ICE, IH
                sh = self._shadow_state()
                _lib.check(plan.lib.cev_fdtd_run_fused(plan.handle, C.byref(st), C.byref(sh), steps, _ptr(waveforms),
                                                       _ptr(partials), self._stream()))
            else:
                _lib.check(plan.lib.cev_fdtd_run(plan.handle, C.byref(st), steps, _ptr(waveforms), _ptr(partials),
                                                 self._stream()))
            if monitoring:      # detach: other callers of the C loop (the adjoint's recomputation) must not accumulate
                _lib.check(plan.lib.cev_fdtd_bind_monitors(plan.handle, None, None))
            self.t_index += steps
            if refresh:
                self._refresh_E()
            if self._n_probes == 0:
                return torch.zeros((steps, 0), dtype=torch.float64, device=self.device)
            return fold_probes(plan, partials, self._n_probes, self._stream())

    def _use_fused(self, steps):
        """The fused full-step kernel (step_v4.cuh) moves 15 instead of 21 words per cell but recomputes a halo.
        With PML code compiled in it needs 168-240 registers and is 15-40 % SLOWER than the two tuned half-step
        kernels; its PML-free instantiation (128 registers) is 20-25 % FASTER (profiles/README.md).  So: automatic on
        grids without any PML (periodic on all axes) of >= 2^21 cells, opt-in elsewhere (`kernel_variant` 4)."""
        kv = self._options.get("kernel_variant", 0)
        if kv not in (0, 4, 5) or steps < 2 or not self._fused_step:
            return False
        if self._options.get("active_components", 63) != 63:
            return False
        if kv in (4, 5):        # 5 = hybrid: lean fused kernel on the PML-free interior, general kernels on the shell
            return True
        return not any(int(p) for p in self.npml) and self.N >= (1 << 21) and min(self.grid_shape) >= 32

    @_host
    def _shadow_state(self):
        if self._shadow is None:
            self._shadow = ([torch.empty_like(t) for t in self._H], [torch.empty_like(t) for t in self._D],
                            [torch.empty_like(t) for t in self._pml["ICE"]], [torch.empty_like(t) for t in self._pml["IH"]])
        sh = _lib.cev_state()
        sh.H, sh.D, sh.ICE, sh.IH = (_ptr3(ts) for ts in self._shadow)
        return sh

    @_host
    def _refresh_E(self):
        plan = self._ensure_plan()
        En = [torch.empty_like(t) for t in self._D]
        st = self._state()
        _lib.check(plan.lib.cev_fdtd_compute_E(plan.handle, C.byref(st), None, _ptr3(En), self._stream()))
        self._E = En
        self._publish()
