"""2-D TM tuning sweep (GPU box): python scripts/tune2d.py N dtype 'opt=v,opt=v' ...  (primal run() and 4-tangent jvp_run())"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceviche_b200

N = int(sys.argv[1])
dtype = torch.float64 if sys.argv[2] == "f64" else torch.float32
shape = (N, N, 1)
eps = 1 + np.random.default_rng(0).random(shape)
steps = 200
prof = np.zeros(shape); prof[N // 2, N // 2, 0] = 1.0
mask = np.zeros(shape); mask[N // 3, :, 0] = 1.0
t = np.arange(steps)
wave = np.exp(-(t - 60) ** 2 / (2 * 20 ** 2)) * np.cos(0.2 * t)
B = int(os.environ.get('TUNE_B', '4'))
V = torch.as_tensor(np.random.default_rng(1).standard_normal((B,) + shape))


def timed(fn):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3


for combo in sys.argv[3:]:
    F = ceviche_b200.fdtd(eps, 5e-8, [20, 20, 0], dtype=dtype)
    for kv in combo.split(","):
        if kv:
            k, v = kv.split("=")
            F.set_option(k, int(v))

    def primal():
        F.initialize_fields()
        F.run(steps, [("z", prof, wave)], [("Ez", mask)])

    def jvp():
        F.initialize_fields()
        F.jvp_run(steps, V, [("z", prof, wave)], [("Ez", mask)])
    s1, s2 = timed(primal), timed(jvp)
    print("%5d %s %-34s primal %7.2f us/step %6.1f Gcell/s | jvp(B=%d) %7.2f us/step %6.1f Gcell/s incl. tangents" % (
        N, sys.argv[2], combo, s1 / steps * 1e6, N * N * steps / s1 / 1e9, B, s2 / steps * 1e6,
        N * N * steps * (1 + B) / s2 / 1e9), flush=True)
