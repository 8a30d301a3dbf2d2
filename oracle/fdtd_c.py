"""ctypes front-end of oracle/fdtd_c.c (TEST INFRASTRUCTURE): the numpy oracle with its step replaced by the fused,
OpenMP-parallel C restatement.  Bit-identical to OracleFDTD (tests/test_oracle_c.py); used for the multi-core CPU
baseline of bench.py and for parity checks at sizes where ~102 numpy passes per step are too slow."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np

from .fdtd_numpy import OracleFDTD

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "fdtd_c.c")
LIB = os.path.join(HERE, "_build", "liboracle_fdtd.so")
_lib = None


def build(force=False):
    """gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC oracle/fdtd_c.c -> oracle/_build/liboracle_fdtd.so"""
    if not force and os.path.isfile(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    gcc = shutil.which("gcc")
    if gcc is None:
        raise RuntimeError("gcc not found: cannot build the C oracle")
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    tmp = LIB + ".tmp%d" % os.getpid()
    subprocess.run([gcc, "-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", tmp, SRC],
                   check=True)
    os.replace(tmp, LIB)
    return LIB


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        P = C.POINTER(C.c_double)
        lib.oracle_fdtd_step.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double] + [P * 3] * 7 + [P * 12, P * 12, P * 3, P * 3]
        lib.oracle_fdtd_step.restype = None
        lib.oracle_omp_threads.argtypes = [C.c_int]
        lib.oracle_omp_threads.restype = C.c_int
        _lib = lib
    return _lib


def set_threads(n=0):
    """Use n OpenMP threads (0 = leave as is); returns the number parallel regions will use."""
    return int(load().oracle_omp_threads(int(n)))


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class OracleFDTDC(OracleFDTD):
    """OracleFDTD whose step() is the C restatement (same state attributes, same results bit for bit)."""

    def __init__(self, eps_r, dL, npml):
        super().__init__(eps_r, dL, npml, materialize=True)
        self._lib = load()
        P = C.POINTER(C.c_double)
        self._mHc = [np.ascontiguousarray(m, dtype=np.float64) for c in range(3) for m in self.mH[c]]
        self._mDc = [np.ascontiguousarray(m, dtype=np.float64) for c in range(3) for m in self.mD[c]]
        self._mH12 = (P * 12)(*[_p(m) for m in self._mHc])
        self._mD12 = (P * 12)(*[_p(m) for m in self._mDc])

    def set_eps(self, eps_r):
        super().set_eps(eps_r)
        self.mE = [np.ascontiguousarray(m) for m in self.mE]

    def step(self, Jx=None, Jy=None, Jz=None):
        P = C.POINTER(C.c_double)
        self.t_index += 1
        Js = []
        for J in (Jx, Jy, Jz):
            Js.append(None if J is None else np.ascontiguousarray(np.broadcast_to(np.asarray(J, dtype=np.float64), self.shape)))
        p3 = lambda arrs: (P * 3)(*[_p(a) if a is not None else None for a in arrs])
        Nx, Ny, Nz = self.shape
        self._lib.oracle_fdtd_step(Nx, Ny, Nz, float(self.dL), p3(self.H), p3(self.D), p3(self.E), p3(self.ICE), p3(self.IH),
                                   p3(self.ICH), p3(self.ID), self._mH12, self._mD12, p3(self.mE), p3(Js))
        return self.fields()
