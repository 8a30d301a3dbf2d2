"""ctypes binding of the C ABI in include/ceviche_b200.h.

There is no CPU fallback: if the CUDA library cannot be loaded (and cannot be built
because nvcc is absent) importing the compute path raises."""
import ctypes as C
import os
import threading

from . import build as _build

_lock = threading.Lock()
_lib = None

c_void_p3 = C.c_void_p * 3
c_double3 = C.c_double * 3
c_dptr3 = C.POINTER(C.c_double) * 3


class cev_state(C.Structure):
    _fields_ = [(n, c_void_p3) for n in
                ("H", "D", "inv_eps", "ICE", "IH", "ICH", "ID", "D_xhi", "inv_eps_xhi", "H_xlo")]


class cev_points(C.Structure):
    _fields_ = [("field", C.c_int32), ("n", C.c_int64), ("idx", C.c_void_p), ("cell0", C.c_int64),
                ("weight", C.c_void_p)]


class cev_tangent(C.Structure):
    _fields_ = [("d_inv_eps", c_void_p3), ("D_primal", c_void_p3)]


class cev_adjoint(C.Structure):
    _fields_ = [(n, c_void_p3) for n in ("lH", "lD", "lICE", "lIH", "lICH", "lID", "gC2", "G_mE")] + [("g_box", C.c_int64 * 6), ("gC", c_void_p3)]


class cev_halo_layout(C.Structure):
    _fields_ = [("bytes", C.c_size_t), ("plane_bytes", C.c_size_t), ("D_hi", C.c_size_t * 2), ("inv_eps_hi", C.c_size_t * 2),
                ("H_lo", C.c_size_t * 2), ("flag_D", C.c_size_t), ("flag_H", C.c_size_t), ("err", C.c_size_t)]


IPC_HANDLE_BYTES = 64


class CevicheB200Error(RuntimeError):
    pass


def _declare(lib):
    P = C.POINTER
    lib.cev_last_error.restype = C.c_char_p
    lib.cev_abi_version.restype = C.c_int
    sigs = {
        "cev_fdtd_create": [P(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64,
                            C.c_double, C.c_double, c_dptr3, c_dptr3],
        "cev_fdtd_destroy": [C.c_void_p],
        "cev_fdtd_pml_shapes": [C.c_void_p, P(C.c_int64 * 3 * 12)],
        "cev_fdtd_set_option": [C.c_void_p, C.c_char_p, C.c_int64],
        "cev_fdtd_step_H": [C.c_void_p, P(cev_state), C.c_void_p, C.c_int64, C.c_int64, C.c_void_p],
        "cev_fdtd_step_D": [C.c_void_p, P(cev_state), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                            C.c_int64, C.c_int64, C.c_void_p],
        "cev_fdtd_compute_E": [C.c_void_p, P(cev_state), P(cev_tangent), C.c_void_p, C.c_void_p],
        "cev_fdtd_step_H_ex": [C.c_void_p, P(cev_state), P(cev_tangent), C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                               C.c_void_p, C.c_void_p],
        "cev_fdtd_step_D_ex": [C.c_void_p, P(cev_state), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p],
        "cev_fdtd_sample_probes": [C.c_void_p, P(cev_state), P(cev_tangent), C.c_int, C.c_int64, C.c_void_p, C.c_void_p],
        "cev_fdtd_jvp_run": [C.c_void_p, P(cev_state), C.c_int, P(cev_state), P(cev_tangent), C.c_int64, C.c_void_p,
                             C.c_void_p, C.c_void_p, C.c_void_p],
        "cev_fdtd_adjoint_step": [C.c_void_p, P(cev_state), P(cev_adjoint), C.c_void_p],
        "cev_fdtd_adjoint_seed": [C.c_void_p, P(cev_state), P(cev_adjoint), C.c_void_p, C.c_void_p],
        "cev_fdtd_adjoint_run": [C.c_void_p, P(cev_state), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, P(cev_adjoint), C.c_void_p],
        "cev_fdtd_adjoint_graph_replays": [C.c_void_p],
        "cev_fdtd_adjoint_part": [C.c_void_p, C.c_int, P(cev_state), P(cev_adjoint), P(c_void_p3), C.c_void_p],
        "cev_fdtd_adjoint_boxed_supported": [C.c_void_p],
        "cev_fdtd_set_recorder": [C.c_void_p, P(C.c_int64), C.c_void_p, C.c_int64],
        "cev_fdtd_adjoint_run_boxed": [C.c_void_p, P(cev_state), C.c_int64, C.c_void_p, C.c_void_p, P(cev_adjoint), C.c_void_p],
        "cev_fdtd_set_sources": [C.c_void_p, C.c_int, P(cev_points)],
        "cev_fdtd_set_probes": [C.c_void_p, C.c_int, P(cev_points), P(C.c_int64)],
        "cev_fdtd_probe_slots": [C.c_void_p, P(C.c_int32)],
        "cev_fdtd_fold_probes": [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p],
        "cev_fdtd_set_monitors": [C.c_void_p, C.c_int, P(cev_points), C.c_int, P(C.c_int64)],
        "cev_fdtd_bind_monitors": [C.c_void_p, C.c_void_p, C.c_void_p],
        "cev_fdtd_run": [C.c_void_p, P(cev_state), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p],
        "cev_fdtd_run_fused": [C.c_void_p, P(cev_state), P(cev_state), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p],
        "cev_fdtd_halo_layout": [C.c_void_p, P(cev_halo_layout)],
        "cev_halo_alloc": [C.c_int, C.c_size_t, P(C.c_void_p), C.c_char_p],
        "cev_halo_open": [C.c_int, C.c_char_p, P(C.c_void_p)],
        "cev_halo_close": [C.c_int, C.c_void_p],
        "cev_halo_free": [C.c_int, C.c_void_p],
        "cev_fdtd_halo_attach": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
        "cev_fdtd_halo_push_static": [C.c_void_p, P(cev_state), C.c_void_p],
        "cev_fdtd_halo_reset": [C.c_void_p, C.c_void_p],
        "cev_fdtd_halo_error": [C.c_void_p, P(C.c_int)],
    }
    for name, argtypes in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.cev_fdtd_adjoint_graph_replays.restype = C.c_int64
    return sigs


EXPORTS = ("cev_last_error", "cev_abi_version", "cev_fdtd_create", "cev_fdtd_destroy", "cev_fdtd_pml_shapes",
           "cev_fdtd_set_option",
           "cev_fdtd_step_H", "cev_fdtd_step_D", "cev_fdtd_compute_E", "cev_fdtd_set_sources",
           "cev_fdtd_set_probes", "cev_fdtd_probe_slots", "cev_fdtd_fold_probes", "cev_fdtd_set_monitors", "cev_fdtd_bind_monitors", "cev_fdtd_run", "cev_fdtd_run_fused", "cev_fdtd_step_H_ex", "cev_fdtd_step_D_ex",
           "cev_fdtd_sample_probes", "cev_fdtd_jvp_run", "cev_fdtd_adjoint_step", "cev_fdtd_adjoint_seed", "cev_fdtd_adjoint_run",
           "cev_fdtd_adjoint_graph_replays", "cev_fdtd_adjoint_part", "cev_fdtd_adjoint_boxed_supported", "cev_fdtd_set_recorder", "cev_fdtd_adjoint_run_boxed",
           "cev_fdtd_halo_layout", "cev_halo_alloc", "cev_halo_open", "cev_halo_close", "cev_halo_free",
           "cev_fdtd_halo_attach", "cev_fdtd_halo_push_static", "cev_fdtd_halo_reset", "cev_fdtd_halo_error")


def load():
    """Load (building first if the in-tree .so is missing or older than its sources)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.LIB_PATH
        if _build.is_stale():
            if _build.find_nvcc() is not None:
                _build.build_library()
            elif not os.path.isfile(path):
                raise CevicheB200Error(
                    "ceviche_b200: %s is missing and nvcc is not available to build it; "
                    "there is no CPU fallback (run __graft_entry__.build())" % path)
        lib = C.CDLL(path)
        _declare(lib)
        _lib = lib
        return lib


def check(rc):
    if rc != 0:
        msg = load().cev_last_error()
        raise CevicheB200Error((msg or b"unknown error").decode())
