"""Host-side cost of the reference-style loop `for t: F.forward(Jz=...)` on config 1 (2-D 200x200): wall time per step and a
cProfile breakdown (GPU box).   python scripts/profile_forward_host.py [steps]"""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceviche_b200
from oracle import cases

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
case = cases.field_case("c1_tm")
F = ceviche_b200.fdtd(case["eps"], case["dL"], case["npml"])
prof = torch.as_tensor(case["sources"][0][1]).cuda()
J = prof * 1.0


def loop(n, fresh_J):
    for t in range(n):
        F.forward(Jz=prof * 0.5 if fresh_J else J)


loop(200, True)
for fresh in (True, False):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loop(steps, fresh)
    torch.cuda.synchronize()
    print("%s: %.2f us per step" % ("J = profile * s(t) built every step" if fresh else "constant J tensor", (time.perf_counter() - t0) / steps * 1e6))
pr = cProfile.Profile()
pr.enable()
loop(steps, False)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
