"""Timings of the other BASELINE configs (1, 4, 5) on one B200: python scripts/bench_configs.py [c1] [c4] [c5]
Prints one JSON line per measurement (recorded in profiles/)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceviche_b200
from oracle import cases

DL = 5e-8
which = sys.argv[1:] or ["c1", "c4", "c5"]
SPEC = os.environ.get("CEV_SPECIALISE", "1") != "0"    # skip provably-zero components (2-D TM: 10 instead of 21 words/cell)
_mk = ceviche_b200.fdtd


def _fdtd(*a, **k):
    F = _mk(*a, **k)
    F.specialise_components = SPEC
    return F


ceviche_b200.fdtd = _fdtd


def timed(fn, reps=1):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3, out


if "c1" in which:
    # config 1: 2-D 200x200, npml 20, point dipole Jz, 1000 steps, fp64 (tests/test_fields_fdtd.py path)
    case = cases.field_case("c1_tm")
    for dtype, name in ((torch.float64, "f64"), (torch.float32, "f32")):
        F = ceviche_b200.fdtd(case["eps"], case["dL"], case["npml"], dtype=dtype)
        F.prepare(case["sources"], case["probes"])
        wf = torch.as_tensor(np.stack([w for _, _, w in case["sources"]], 1)).cuda()
        F.run(1000, waveforms=wf)
        F.initialize_fields()
        s, _ = timed(lambda: F.run(1000, waveforms=wf), reps=5)
        print(json.dumps({"config": "c1 2-D 200x200x1 npml [20,20,0] 1000 steps fused run()", "dtype": name, "seconds": s,
                          "us_per_step": s / 1000 * 1e6, "gcell_per_s": 200 * 200 * 1000 / s / 1e9,
                          "active_components": F._options.get("active_components", 63)}), flush=True)
    F = ceviche_b200.fdtd(case["eps"], case["dL"], case["npml"])
    prof = torch.as_tensor(case["sources"][0][1]).cuda()
    wave = case["sources"][0][2]
    def loop():
        for t in range(1000):
            F.forward(Jz=prof * float(wave[t]))
    loop()
    s, _ = timed(loop)
    print(json.dumps({"config": "c1 per-step forward() API (reference caller loop), 1000 steps", "dtype": "f64", "seconds": s,
                      "us_per_step": s / 1000 * 1e6, "gcell_per_s": 200 * 200 * 1000 / s / 1e9}), flush=True)

if "c4" in which:
    # config 4: reverse-mode gradient of a probe objective w.r.t. eps_r on 256x256x128, checkpointed adjoint
    shape = (256, 256, 128)
    steps = int(os.environ.get("C4_STEPS", "200"))
    rng = np.random.default_rng(1)
    eps_np = np.ones(shape)
    eps_np[96:160, 96:160, 48:80] = 1 + 4.95 * rng.random((64, 64, 32))
    prof = np.zeros(shape); prof[30, 123:133, 61:67] = 1.0
    mask = np.zeros(shape); mask[226, 123:133, 61:67] = 1.0
    t = np.arange(steps)
    wave = 5 * np.exp(-(t - 60) ** 2 / (2 * 20 ** 2)) * np.cos(0.2 * t)
    box = ((96, 160), (96, 160), (48, 80))
    for dtype, name in ((torch.float64, "f64"), (torch.float32, "f32")):
      for region in (None, box):      # gradient w.r.t. every eps_r cell | w.r.t. the design box only
        eps = torch.as_tensor(eps_np).cuda().requires_grad_(True)
        F = ceviche_b200.fdtd(eps, DL, [20, 20, 20], dtype=dtype)
        for kv in os.environ.get("C4_OPTS", "").split(","):
            if kv:
                F.set_option(kv.split("=")[0], int(kv.split("=")[1]))
        F.design_region = region
        def fwd():
            F.eps_r = eps        # like the reference's objective (test_gradients_fdtd.py:73): new graph, fields reset
            return F.run(steps, [("z", prof, wave)], [("Ez", mask)])
        series = fwd()                                   # warm-up: allocator, point-set upload
        torch.autograd.grad((series ** 2).sum(), eps)
        torch.cuda.reset_peak_memory_stats()
        s_f, series = timed(fwd)
        s_f2, series = timed(fwd)            # (the first timed forward occasionally eats a caching-allocator refill of
        s_f = min(s_f, s_f2)                 #  the ~16 GB of checkpoints freed by the warm-up's backward: take the better)
        L = (series ** 2).sum()
        s_b, (g,) = timed(lambda: torch.autograd.grad(L, eps))
        cells = shape[0] * shape[1] * shape[2]
        mode = ("design box %s: D-box record, no recomputation" % (box,)) if region else "all cells: checkpoint every sqrt(steps) + recomputation"
        print(json.dumps({"config": "c4 reverse-mode gradient 256x256x128 npml 20, %d steps, %s" % (steps, mode),
                          "dtype": name, "forward_s": s_f, "backward_s": s_b,
                          "forward_gcell_per_s": cells * steps / s_f / 1e9,
                          "backward_gcell_per_s": cells * steps / s_b / 1e9,
                          "forward_plus_backward_gcell_per_s": cells * steps / (s_f + s_b) / 1e9,
                          "grad_l2_in_box": float(g[96:160, 96:160, 48:80].norm()),
                          "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}), flush=True)
        del F, eps, series, L, g
        torch.cuda.empty_cache()

if "c5" in which:
    # config 5: forward-mode JVP over 16 eps_r perturbations on a 2-D 2048x2048 grating-coupler grid (20 000 steps at its
    # stated size: C5_STEPS=20000).  Geometry and directions after examples/forwardmode_grating_coupler.py:33-98, 138-162
    # scaled to the grid (ceviche_b200.parametrization.grating_coupler): Si slab + teeth on SiO2, the teeth = sigmoid
    # projection of sin^2 around 1 - fill factor; the 16 directions are d eps_r / d(fill factor of tooth group g).
    from ceviche_b200.parametrization import grating_coupler
    shape = (2048, 2048, 1)
    steps = int(os.environ.get("C5_STEPS", "300"))
    B = 16
    G = grating_coupler(2048, 2048, DL, 20, groups=B)
    ff = torch.full((B,), 0.5, dtype=torch.float64, device="cuda")
    eps_t = G.eps_r(ff)
    Vt = G.fill_factor_directions(ff)                      # [16, 2048, 2048, 1] on the device
    prof = np.zeros(shape); prof[G.source_x, G.y_base[0]:G.y_teeth[1], 0] = 1.0          # modal-like sheet across the slab
    mask = np.zeros(shape); mask[G.x_grids[0]:G.x_grids[-1], G.probe_y, 0] = 1.0          # line probe above the grating
    t = np.arange(steps)
    wave = np.exp(-(t - 400) ** 2 / (2 * 120.0 ** 2)) * np.cos(2 * np.pi * 299792458.0 / 1550e-9 * (0.5 * DL / (np.sqrt(3) * 299792458.0)) * t)
    for dtype, name in ((torch.float64, "f64"), (torch.float32, "f32")):
        F = ceviche_b200.fdtd(eps_t, DL, [20, 20, 0], dtype=dtype)
        for kv in os.environ.get("C5_OPTS", "").split(","):
            if kv:
                F.set_option(kv.split("=")[0], int(kv.split("=")[1]))
        F.jvp_run(10, Vt, [("z", prof, wave[:10])], [("Ez", mask)])
        F.initialize_fields()
        s, (series, dseries) = timed(lambda: F.jvp_run(steps, Vt, [("z", prof, wave)], [("Ez", mask)]))
        cells = shape[0] * shape[1]
        F2 = ceviche_b200.fdtd(eps_t, DL, [20, 20, 0], dtype=dtype)
        F2.run(10, [("z", prof, wave[:10])], [("Ez", mask)])
        F2.initialize_fields()
        s1, _ = timed(lambda: F2.run(steps, [("z", prof, wave)], [("Ez", mask)]))
        w = 8 if name == "f64" else 4
        gc = cells * steps * (1 + B) / s / 1e9
        # 2-D TM byte model per cell and time step: primal 10 words; per tangent 8 (H half-step: dD, 1/eps, d(1/eps), D
        # primal, H x2 in, H x2 out -- 1/eps and the primal D shared by the batch: 6 own words) + 4 (D half-step)
        bytes_step = cells * w * (10 + B * 10)
        print(json.dumps({"config": "c5 batched JVP, 16 fill-factor tangents + primal, 2-D 2048x2048 grating coupler npml [20,20,0], %d steps" % steps,
                          "dtype": name, "seconds": s, "gcell_per_s_incl_tangents": gc,
                          "primal_only_seconds": s1, "primal_only_gcell_per_s": cells * steps / s1 / 1e9,
                          "byte_model_words_per_cell_step": 10 + B * 10,
                          "model_GBps": bytes_step * steps / s / 1e9,
                          "series_l2": float(series.norm()), "dseries_l2_per_direction": [float(dseries[b].norm()) for b in range(B)],
                          "teeth": int(G.num_teeth), "active_components": F._options.get("active_components", 63)}), flush=True)
        del F, F2
        torch.cuda.empty_cache()
