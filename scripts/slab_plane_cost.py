"""Relative cost of an x-PML plane in the half-step kernels (for the load-balanced slab partition): a 128 x 1024 x 512 slab
with npml (20, 20, 20) -- 40 of its 128 planes in the x-PML -- against npml (0, 20, 20).   python scripts/slab_plane_cost.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceviche_b200  # noqa: E402
import bench  # noqa: E402

for dtype in (torch.float64, torch.float32):
    t = {}
    for npml in ([20, 20, 20], [0, 20, 20]):
        shape = (128, 1024, 512)
        F = ceviche_b200.fdtd(torch.ones(shape, dtype=torch.float64, device="cuda"), bench.DL, npml, dtype=dtype)
        F.run(3)
        h, d = bench.time_kernels(F, shape, reps=30)
        t[npml[0]] = (h, d)
        del F
        torch.cuda.empty_cache()
    for q, name in ((0, "H"), (1, "D")):
        a = (t[20][q] / t[0][q] - 1) * 128 / 40
        print("%s %s: %.4f ms with 40 x-PML planes, %.4f ms without -> alpha = %.3f" % ("f64" if dtype == torch.float64 else "f32", name, t[20][q], t[0][q], a))
    both = (sum(t[20]) / sum(t[0]) - 1) * 128 / 40
    print("   step: alpha = %.3f" % both, flush=True)
