"""Inverse-design iteration on a launch-bound 2-D grid (the reference's everyday use: 2-D, a few hundred cells a side,
reverse mode): forward + checkpointed adjoint with and without CUDA-graph replay.   python scripts/bench_small_adjoint.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceviche_b200  # noqa: E402

for N, steps in ((200, 2000), (400, 2000)):
    shape = (N, N, 1)
    rng = np.random.default_rng(0)
    eps_np = 1 + 3 * rng.random(shape)
    prof = np.zeros(shape); prof[N // 4, N // 2, 0] = 1.0
    mask = np.zeros(shape); mask[3 * N // 4, N // 2 - 5:N // 2 + 5, 0] = 1.0
    t = np.arange(steps)
    wave = np.exp(-(t - 300) ** 2 / (2 * 60.0 ** 2)) * np.cos(0.2 * t)
    for use_graph in (0, -1):
        eps = torch.as_tensor(eps_np).cuda().requires_grad_(True)
        F = ceviche_b200.fdtd(eps, 5e-8, [20, 20, 0])
        F.set_option("use_graph", use_graph)

        def iteration():
            F.eps_r = eps
            s = F.run(steps, [("z", prof, wave)], [("Ez", mask)])
            (g,) = torch.autograd.grad((s ** 2).sum(), eps)
            return g
        iteration()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            g = iteration()
        torch.cuda.synchronize()
        el = (time.perf_counter() - t0) / 3
        plan = F._ensure_plan()
        print("%dx%d 2-D TM, %d steps, forward + adjoint: use_graph=%2d  %.1f ms per iteration (%.1f us per time step), graph replays %d, |g| %.6e"
              % (N, N, steps, use_graph, el * 1e3, el / steps * 1e6, plan.lib.cev_fdtd_adjoint_graph_replays(plan.handle), float(g.norm())), flush=True)
