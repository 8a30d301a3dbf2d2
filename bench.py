#!/usr/bin/env python
"""Benchmark of the FDTD hot path (BASELINE.json: Gcell-updates/s + % HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype f64|f32] [--impl ours|reference]

Workload at N = 1: BASELINE config 2 -- 3-D 256^3 dielectric waveguide splitter, npml 20, Jz sheet
source, two arm probes.  One bench "step" = one batch of CHUNK (default 1000) FDTD time steps, so the
default K = 10 is the config's 10 000 steps.  `value` = cells * time-steps / device time with the
state resident in HBM; `e2e` = the same through the public API with HOST buffers (new eps_r uploaded,
sources/probes uploaded, probe series downloaded, every step).  N > 1: x-slab decomposition of a
(256*N) x 256 x 256 grid (weak scaling), NCCL halo exchange (see ceviche_b200/slab.py).

`--impl reference`: the reference's CPU algorithm (numpy port in oracle/, pinned bit-for-bit to the
reference) timed on the host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DL = 5e-8
NPML = [20, 20, 20]
B_ALG_WORDS = 21          # H sweep 12w (D,1/eps,H in; H out) + D sweep 9w (H,D in; D out)
H_KERNEL_WORDS = 12
D_KERNEL_WORDS = 9


# ----------------------------------------------------------------------------- workload
def splitter_eps(shape, dtype=np.float64):
    """Config 2 geometry: background 1.0, core 5.9536; a 10x6 (y x z) guide along x that
    splits linearly after the midpoint into two arms ending at y = Ny/2 +- 40*(Ny/256)."""
    Nx, Ny, Nz = shape
    eps = np.ones(shape, dtype=dtype)
    core = 5.9536
    cy, cz = Ny // 2, Nz // 2
    hy, hz = 5, 3
    off_max = 40 * Ny // 256
    for i in range(Nx):
        if i < Nx // 2:
            centres = [cy]
        else:
            f = min(1.0, (i - Nx // 2) / max(1, (Nx // 2 - Nx // 8)))
            d = int(round(off_max * f))
            centres = [cy - d, cy + d]
        for c in centres:
            eps[i, c - hy:c + hy, cz - hz:cz + hz] = core
    return eps


def workload(shape, chunk):
    Nx, Ny, Nz = shape
    from ceviche_b200.constants import C_0
    dt = 0.5 * DL / (np.sqrt(3) * C_0)
    eps = splitter_eps(shape)
    cy, cz = Ny // 2, Nz // 2
    prof = np.zeros(shape)
    prof[30 * Nx // 256, cy - 5:cy + 5, cz - 3:cz + 3] = 1.0
    t = np.arange(chunk)
    omega = 2 * np.pi * C_0 / 2e-6
    wave = 5 * np.exp(-(t - 2000) ** 2 / (2 * 100 ** 2)) * np.cos(omega * dt * t)
    off = 40 * Ny // 256
    probes = []
    for c in (cy - off, cy + off):
        m = np.zeros(shape)
        m[226 * Nx // 256, c - 5:c + 5, cz - 3:cz + 3] = 1.0
        probes.append(("Ez", m))
    return dict(eps=eps, sources=[("z", prof, wave)], probes=probes)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- reference arm
def cpu_reference(shape, n_steps, warm=1, multicore=False):
    """The reference's CPU algorithm (oracle port, full 3-D coefficient arrays like fdtd.py:265-316)
    stepped on the host on the config-2 workload.  Returns (Gcell/s, seconds per time step, sample str).
    multicore: the fused C / OpenMP restatement (oracle/fdtd_c.c, bit-identical) on all host cores instead of the
    reference-faithful single-threaded numpy passes."""
    if multicore:
        from oracle.fdtd_c import OracleFDTDC as OracleFDTD
    else:
        from oracle.fdtd_numpy import OracleFDTD
    wl = workload(shape, max(n_steps + warm, 8))
    sim = OracleFDTD(wl["eps"], DL, NPML) if multicore else OracleFDTD(wl["eps"], DL, NPML, materialize=True)
    comp, prof, wave = wl["sources"][0]
    for t in range(warm):
        sim.step(Jz=prof * wave[t])
    t0 = time.perf_counter()
    for t in range(warm, warm + n_steps):
        f = sim.step(Jz=prof * wave[t])
        for key, mask in wl["probes"]:
            np.sum(f[key] * mask)
    dt = time.perf_counter() - t0
    cells = shape[0] * shape[1] * shape[2]
    return cells * n_steps / dt / 1e9, dt / n_steps, "%dx%dx%d fp64, %d time steps after %d warm-up" % (*shape, n_steps, warm)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    shape = (256, 256, 256)
    try:
        from oracle.fdtd_numpy import OracleFDTD
        wl = workload(shape, 8 + args.steps + args.warmup)
        sim = OracleFDTD(wl["eps"], DL, NPML, materialize=True)
    except MemoryError:
        shape = (128, 128, 128)
        wl = workload(shape, 8 + args.steps + args.warmup)
        sim = OracleFDTD(wl["eps"], DL, NPML, materialize=True)
    comp, prof, wave = wl["sources"][0]

    def one(t):
        f = sim.step(Jz=prof * wave[t])
        for key, mask in wl["probes"]:
            np.sum(f[key] * mask)
    for t in range(args.warmup):
        one(t)
    t0 = time.perf_counter()
    for t in range(args.warmup, args.warmup + args.steps):
        one(t)
    el = time.perf_counter() - t0
    cells = shape[0] * shape[1] * shape[2]
    val = cells * args.steps / el / 1e9
    sample = "one FDTD time step of the %dx%dx%d config-2 grid per bench step (fp64 numpy port of ceviche/fdtd.py, 1 thread)" % shape
    line = {"impl": "reference", "metric": "Gcell-updates/s", "value": val, "unit": "Gcell/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config 2: 3-D %dx%dx%d waveguide splitter, npml 20" % shape, "sample": sample},
            "cpu_baseline": {"value": val, "unit": "Gcell/s", "cores": 1, "kind": "port", "sample": sample,
                             "host_cores": os.cpu_count()},
            "e2e": {"value": val, "unit": "Gcell/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    # next to the reference-faithful figure (numpy passes, one thread -- all the reference can use): the same algorithm
    # as a fused C / OpenMP pass on every host core (oracle/fdtd_c.c, bit-identical results)
    try:
        del sim
        v2, sec2, sample2 = cpu_reference(shape, 6, multicore=True)
        line["cpu_baseline"]["c_openmp_port"] = {"value": v2, "unit": "Gcell/s", "cores": os.cpu_count(), "kind": "port",
                                                 "sample": sample2 + " (oracle/fdtd_c.c, gcc -O2 -fopenmp)",
                                                 "s_per_time_step": sec2}
    except Exception as e:
        line["cpu_baseline"]["c_openmp_port"] = {"unavailable": str(e)[:200]}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import ceviche_b200
    from ceviche_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    w = 8 if args.dtype == "f64" else 4
    chunk = args.chunk
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    if world > 1:
        return run_slabs(args, dist, dev, local, rank, world, dtype, w, hbm_peak, peak_src)

    shape = tuple(args.grid)
    cells = shape[0] * shape[1] * shape[2]
    wl = workload(shape, chunk)
    F = ceviche_b200.fdtd(wl["eps"], DL, NPML, dtype=dtype, arith=args.arith)
    for kv in args.opt:
        k, v = kv.split("=")
        F.set_option(k, int(v))
    srcs, probes = wl["sources"], wl["probes"]

    def sync():
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput -------------------------------------------------
    wave_dev = torch.as_tensor(np.stack([s_[2] for s_ in srcs], 1)).to(dev)
    F.prepare(srcs, probes)        # profiles / masks uploaded once: inputs resident in HBM
    for _ in range(args.warmup):
        F.run(chunk, waveforms=wave_dev)
    sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        ev0.record()
        for _ in range(args.steps):
            F.run(chunk, waveforms=wave_dev)
        ev1.record()
        sync()
    ms = ev0.elapsed_time(ev1)
    ms_per_step = ms / args.steps
    value = cells * chunk * args.steps / (ms * 1e-3) / 1e9
    launches = args.steps * (chunk * 2 + 4)     # per run(): 2 half-step kernels per time step (sources and probes ride
                                                # inside them) + the trailing probe launch + 3 E-materialisation launches

    # ---- per-kernel timing of the two half-step kernels (events on the launch stream) ----
    import ctypes as C
    plan = F._ensure_plan()
    st = F._state()
    s = F._stream()
    reps = 50

    def time_kernel(fn):
        for _ in range(5):
            fn()
        sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        sync()
        return a.elapsed_time(b) / reps

    h_ms = time_kernel(lambda: _lib.check(plan.lib.cev_fdtd_step_H(plan.handle, C.byref(st), None, 0, shape[0], s)))
    d_ms = time_kernel(lambda: _lib.check(plan.lib.cev_fdtd_step_D(plan.handle, C.byref(st), None, None, None, None,
                                                                  0, shape[0], s)))
    h_gbs = cells * H_KERNEL_WORDS * w / (h_ms * 1e-3) / 1e9
    d_gbs = cells * D_KERNEL_WORDS * w / (d_ms * 1e-3) / 1e9
    step_gbs = value * B_ALG_WORDS * w
    traffic = None
    prof_json = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.isfile(prof_json):
        try:
            traffic = json.load(open(prof_json)).get("%s_%dx%dx%d" % ((args.dtype,) + shape), {}).get("step_H_bytes")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "step_H (H half-step: D, 1/eps, H in; H out = 12 words/cell)",
                "achieved": h_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": h_gbs / hbm_peak, "traffic": traffic,
                "peak_source": peak_src, "ms_per_launch": h_ms,
                "step_D": {"achieved": d_gbs, "frac": d_gbs / hbm_peak, "ms_per_launch": d_ms, "words_per_cell": D_KERNEL_WORDS},
                "whole_step": {"bytes_per_cell_update": B_ALG_WORDS * w, "achieved": step_gbs, "frac": step_gbs / hbm_peak,
                               "frac_of_nominal_8TBs": step_gbs / 8000.0}}

    # ---- end to end through the public API with host buffers ---------------------------
    eps_host = torch.as_tensor(wl["eps"]).to(dtype).pin_memory()
    wave_host = torch.as_tensor(np.stack([s_[2] for s_ in srcs], 1)).pin_memory()
    geo = [(c, p) for c, p, _ in srcs]
    h2d = eps_host.numel() * eps_host.element_size() + wave_host.numel() * 8 + sum(p.nbytes for _, p in geo) + sum(m.nbytes for _, m in probes)
    d2h = chunk * len(probes) * 8

    def e2e_step():
        F.eps_r = eps_host.to(dev, non_blocking=True)      # upload + Yee averaging + 1/eps + field reset
        series = F.run(chunk, geo, probes, waveforms=wave_host.to(dev, non_blocking=True))
        return series.cpu()
    if args.no_e2e:
        e2e_s, e2e_val = float("nan"), None
    else:
        for _ in range(max(1, args.warmup // 2)):
            e2e_step()
        sync()
        n_e2e = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            out = e2e_step()
        sync()
        e2e_s = (time.perf_counter() - t0) / n_e2e
        e2e_val = cells * chunk / e2e_s / 1e9

    # ---- CPU baseline: the reference algorithm on the host cores, bounded sample --------
    cpu = None
    if not args.no_cpu:
        try:
            v, sec, sample = cpu_reference((256, 256, 256) if cells >= 256 ** 3 else shape, 2)
        except MemoryError:
            v, sec, sample = cpu_reference((128, 128, 128), 4)
        cpu = {"value": v, "unit": "Gcell/s", "cores": 1, "kind": "port", "sample": sample,
               "host_cores": os.cpu_count(), "s_per_time_step": sec}
        try:      # the same algorithm as one fused C pass per half-step on every host core (bit-identical results)
            v2, sec2, sample2 = cpu_reference((256, 256, 256) if cells >= 256 ** 3 else shape, 6, multicore=True)
            cpu["c_openmp_port"] = {"value": v2, "unit": "Gcell/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": sample2 + " (oracle/fdtd_c.c, gcc -O2 -fopenmp)", "s_per_time_step": sec2}
        except Exception as e:      # no gcc / OpenMP on the box: the faithful numpy figure stands alone
            cpu["c_openmp_port"] = {"unavailable": str(e)[:200]}

    line = {"metric": "Gcell-updates/s", "value": value, "unit": "Gcell/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": "config 2: 3-D %dx%dx%d dielectric waveguide splitter, npml 20, Jz sheet source, 2 arm probes" % shape,
                       "time_steps_per_bench_step": chunk, "arith": "f64" if F.arith_f64 else "f32",
                       "l2": "state %.0f MB >> 126 MB L2 (inputs larger than L2, no flush needed)" % (cells * w * 9 / 1e6)},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": "Gcell/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_s * 1e3},
            "gpu_launches": int(launches), "clocks": clk.summary()}
    print(json.dumps(line))


def splitter_eps_slab(shape, lo, hi):
    """Planes lo-1 .. hi-1 (periodic) of the config-3 permittivity: the config-2 splitter stretched to
    the global grid, built slab by slab (the dense 1024x1024x512 array is never materialised)."""
    Nx, Ny, Nz = shape
    out = np.ones((hi - lo + 1, Ny, Nz))
    cy, cz, off_max = Ny // 2, Nz // 2, 40 * Ny // 256
    for q, i in enumerate(range(lo - 1, hi)):
        i %= Nx
        if i < Nx // 2:
            centres = [cy]
        else:
            d = int(round(off_max * min(1.0, (i - Nx // 2) / max(1, (Nx // 2 - Nx // 8)))))
            centres = [cy - d, cy + d]
        for c in centres:
            out[q, c - 5:c + 5, cz - 3:cz + 3] = 5.9536
    return out


def _box(i, j0, j1, k0, k1, Ny, Nz, val=1.0):
    jj, kk = np.meshgrid(np.arange(j0, j1), np.arange(k0, k1), indexing="ij")
    ijk = np.stack([np.full(jj.size, i), jj.ravel(), kk.ravel()], 1)
    return {"ijk": ijk, "w": np.full(jj.size, val), "Ny": Ny, "Nz": Nz}


def run_slabs(args, dist, dev, local, rank, world, dtype, w, hbm_peak, peak_src):
    """N > 1: BASELINE config 3 -- 1024 x 1024 x 512 (npml 20) cut into x-slabs, one per GPU, NCCL halo
    exchange.  Total work is fixed as N grows (strong scaling)."""
    import ctypes as C
    import torch
    from ceviche_b200 import _lib
    from ceviche_b200.constants import C_0
    from ceviche_b200.slab import SlabFDTD, partition
    shape = tuple(args.slab_grid)
    Nx, Ny, Nz = shape
    cells = Nx * Ny * Nz
    chunk = args.slab_chunk
    lo, hi = partition(Nx, world)[rank]
    cy, cz, off = Ny // 2, Nz // 2, 40 * Ny // 256
    sources = [("z", _box(30 * Nx // 256, cy - 5, cy + 5, cz - 3, cz + 3, Ny, Nz))]
    probes = [("Ez", _box(226 * Nx // 256, c - 5, c + 5, cz - 3, cz + 3, Ny, Nz)) for c in (cy - off, cy + off)]
    dt = 0.5 * DL / (np.sqrt(3) * C_0)
    t = np.arange(chunk)
    wave = (5 * np.exp(-(t - 2000) ** 2 / (2 * 100 ** 2)) * np.cos(2 * np.pi * C_0 / 2e-6 * dt * t))[:, None]
    eps_local = splitter_eps_slab(shape, lo, hi)
    sim = SlabFDTD(shape, eps_local, DL, NPML, dtype=dtype, device=dev)
    for kv in args.opt:
        k, v = kv.split("=")
        _lib.check(sim.be.plan.lib.cev_fdtd_set_option(sim.be.plan.handle, k.encode(), int(v)))
    sim.prepare(sources, probes)
    wave_dev = torch.as_tensor(wave).to(dev)

    def sync():
        torch.cuda.synchronize(dev)
        dist.barrier()

    for _ in range(args.warmup):
        sim.run(chunk, wave_dev)
    sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        ev0.record()
        for _ in range(args.steps):
            sim.run(chunk, wave_dev)
        ev1.record()
        sync()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    value = cells * chunk * args.steps / (ms * 1e-3) / 1e9

    # per-kernel timing of the local H half-step (whole local slab, halo in place)
    be = sim.be
    be.new_partials(1)
    reps = 20
    for _ in range(3):
        be.step_H(0, be.nx, -1)
    torch.cuda.synchronize(dev)
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        be.step_H(0, be.nx, -1)
    b_.record()
    torch.cuda.synchronize(dev)
    h_ms = a.elapsed_time(b_) / reps
    local_cells = be.nx * Ny * Nz
    h_gbs = local_cells * H_KERNEL_WORDS * w / (h_ms * 1e-3) / 1e9
    step_gbs = value * B_ALG_WORDS * w / world

    # end to end: host eps slab -> new simulator -> run -> series on the host
    eps_host = torch.as_tensor(eps_local).pin_memory()
    dist.barrier()
    t0 = time.perf_counter()
    sim2 = SlabFDTD(shape, eps_host.to(dev, non_blocking=True), DL, NPML, dtype=dtype, device=dev)
    sim2.prepare(sources, probes)
    out = sim2.run(chunk, torch.as_tensor(wave).pin_memory().to(dev, non_blocking=True)).cpu()
    sync()
    e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e)
    if rank == 0:
        line = {"metric": "Gcell-updates/s", "value": value, "unit": "Gcell/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                "config": {"workload": "config 3: 3-D %dx%dx%d splitter, npml 20, x-slabs over %d GPUs, NCCL halo exchange" % (shape + (world,)),
                           "time_steps_per_bench_step": chunk, "planes_per_gpu": hi - lo,
                           "l2": "per-GPU state %.0f MB >> 126 MB L2" % (local_cells * w * 9 / 1e6)},
                "roofline": {"bound": "hbm", "kernel": "step_H on the local slab (12 words/cell)", "achieved": h_gbs, "peak": hbm_peak,
                             "unit": "GB/s", "frac": h_gbs / hbm_peak, "traffic": None, "peak_source": peak_src, "ms_per_launch": h_ms,
                             "whole_step_per_gpu": {"achieved": step_gbs, "frac": step_gbs / hbm_peak}},
                "cpu_baseline": None,
                "e2e": {"value": cells * chunk / e2e_s / 1e9, "unit": "Gcell/s",
                        "h2d_bytes_per_step": int(eps_host.numel() * 8 * world + wave.size * 8 * world),
                        "d2h_bytes_per_step": int(chunk * len(probes) * 8 * world), "ms_per_step": e2e_s * 1e3},
                "gpu_launches": int(args.steps * (chunk * 4 + 1)), "clocks": clk.summary()}
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--arith", default=None, choices=[None, "f64", "f32"])
    ap.add_argument("--chunk", type=int, default=1000, help="FDTD time steps per bench step")
    ap.add_argument("--grid", type=int, nargs=3, default=[256, 256, 256])
    ap.add_argument("--slab-grid", type=int, nargs=3, default=[1024, 1024, 512], help="global grid for N > 1 (config 3)")
    ap.add_argument("--slab-chunk", type=int, default=200, help="FDTD time steps per bench step for N > 1")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (kernel tuning runs)")
    ap.add_argument("--opt", action="append", default=[], help="plan option name=value (repeatable)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    main()
