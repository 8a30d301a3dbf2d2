"""Kernel tuning sweep (GPU box): times the two half-step kernels for plan-option combinations.
usage: python scripts/tune.py N dtype 'opt=v,opt=v' ['opt=v,...' ...]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceviche_b200
from ceviche_b200 import _lib

N = int(sys.argv[1])
dtype = torch.float64 if sys.argv[2] == "f64" else torch.float32
w = 8 if dtype == torch.float64 else 4
shape = (N, N, N)
eps = 1 + np.random.default_rng(0).random(shape)
NPML = [int(v) for v in os.environ.get("TUNE_NPML", "20,20,20").split(",")]
NPML = NPML * 3 if len(NPML) == 1 else NPML
F = ceviche_b200.fdtd(eps, 5e-8, NPML, dtype=dtype)
for t in F._H + F._D:
    t.normal_()
plan = F._ensure_plan()
st = F._state()
s = F._stream()
cells = N ** 3


def timeit(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


RUN_STEPS = int(os.environ.get("TUNE_RUN", "0"))     # > 0: time the fused caller loop run() instead of single half-steps
for combo in sys.argv[3:]:
    for kv in combo.split(","):
        if kv:
            k, v = kv.split("=")
            F.set_option(k, int(v))
    if RUN_STEPS:
        F._active = 63
        F._apply_active(63)
        wf = torch.zeros((RUN_STEPS, 0), dtype=torch.float64, device="cuda")
        F.set_sources([])
        F.set_probes([])
        t = timeit(lambda: F._run_raw(RUN_STEPS, wf, refresh=False), reps=3) / RUN_STEPS
        print("%4d %s %-40s run(): %.4f ms/step  %.2f Gcell/s  21w: %5.0f GB/s  fused=%s" % (
            N, sys.argv[2], combo, t, cells / t / 1e6, cells * 21 * w / t / 1e6, F._use_fused(RUN_STEPS)), flush=True)
        continue
    h = timeit(lambda: _lib.check(plan.lib.cev_fdtd_step_H(plan.handle, C.byref(st), None, 0, N, s)))
    d = timeit(lambda: _lib.check(plan.lib.cev_fdtd_step_D(plan.handle, C.byref(st), None, None, None, None, 0, N, s)))
    print("%4d %s %-40s H %.4f ms %6.0f GB/s | D %.4f ms %6.0f GB/s | step %.2f Gcell/s %5.0f GB/s" % (
        N, sys.argv[2], combo, h, cells * 12 * w / h / 1e6, d, cells * 9 * w / d / 1e6,
        cells / (h + d) / 1e6, cells * 21 * w / (h + d) / 1e6), flush=True)
