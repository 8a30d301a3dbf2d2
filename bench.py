#!/usr/bin/env python
"""Benchmark of the FDTD hot path (BASELINE.json: Gcell-updates/s + % HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype f64|f32] [--impl ours|reference]

N = 1: BASELINE config 2 -- 3-D 256^3 dielectric waveguide splitter, npml 20, Jz sheet source, two arm probes,
fp64.  One bench "step" = one chunk of CHUNK (default 1000) FDTD time steps of ONE continuous simulation: the
timed region starts from zero fields and advances the waveform chunk by chunk, so the default K = 10 is the
config's real 10 000-step run (pulse centred at t = 2000).  `value` = cells * time steps / device time with the
state resident in HBM; `e2e` = the same through the public API with HOST buffers (eps_r, profiles, masks and
waveform uploaded, probe series downloaded, every step).  The same JSON line also carries
  `parity_check`  the 256^3 grid against the CPU oracle on an early-pulse waveform (all nine fields, rel-L2),
  `other`         fp32 at 256^3, and the north_star target grid 512^3 in fp64 and fp32, each with its roofline fractions,
  `scale_anchor`  BASELINE config 3's grid (1024 x 1024 x 512) on ONE GPU: the same-workload anchor of the 1 -> 8 curve.
N > 1: config 3 cut into x-slabs, one per GPU (strong scaling); halo planes travel as direct stores into the
neighbour's peer-mapped halo buffer from inside the half-step kernels.  Before the timed region every run checks
the config-3 parity grid (256 x 128 x 64) on the N slabs against the same grid on one GPU, bit for bit, and exits
non-zero on a mismatch.

`--impl reference`: the reference's own ceviche/fdtd.py (unmodified copies under git-ignored oracle/_ref/, made by
__graft_entry__.build() where /root/reference exists; else the numpy port, pinned bit-for-bit to it) stepped on the
host on the config-2 grid.
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
NOMINAL_GBS = 8000.0      # the figure north_star quotes


# ----------------------------------------------------------------------------- workload
def _centres(i, Nx, Ny):
    cy, off_max = Ny // 2, 40 * Ny // 256
    if i < Nx // 2:
        return [cy]
    d = int(round(off_max * min(1.0, (i - Nx // 2) / max(1, (Nx // 2 - Nx // 8)))))
    return [cy - d, cy + d]


def splitter_eps(shape, dtype=np.float64):
    """Config 2 geometry: background 1.0, core 5.9536; a 10x6 (y x z) guide along x that
    splits linearly after the midpoint into two arms ending at y = Ny/2 +- 40*(Ny/256)."""
    Nx, Ny, Nz = shape
    eps = np.ones(shape, dtype=dtype)
    cz = Nz // 2
    for i in range(Nx):
        for c in _centres(i, Nx, Ny):
            eps[i, c - 5:c + 5, cz - 3:cz + 3] = 5.9536
    return eps


def splitter_eps_device(shape, device, lo=0, hi=None):
    """The same geometry built on the device (planes lo-1 .. hi-1, periodic, when a slab is asked for)."""
    import torch
    Nx, Ny, Nz = shape
    planes = range(0, Nx) if hi is None else range(lo - 1, hi)
    eps = torch.ones((len(planes), Ny, Nz), dtype=torch.float64, device=device)
    cz = Nz // 2
    for q, i in enumerate(planes):
        for c in _centres(i % Nx, Nx, Ny):
            eps[q, c - 5:c + 5, cz - 3:cz + 3] = 5.9536
    return eps


def pulse(n_steps, t0=2000.0, sigma=100.0):
    from ceviche_b200.constants import C_0
    dt = 0.5 * DL / (np.sqrt(3) * C_0)
    t = np.arange(n_steps)
    return 5 * np.exp(-(t - t0) ** 2 / (2 * sigma ** 2)) * np.cos(2 * np.pi * C_0 / 2e-6 * dt * t)


def workload(shape, n_steps, t0=2000.0, sigma=100.0):
    Nx, Ny, Nz = shape
    eps = splitter_eps(shape)
    cy, cz = Ny // 2, Nz // 2
    prof = np.zeros(shape)
    prof[30 * Nx // 256, cy - 5:cy + 5, cz - 3:cz + 3] = 1.0
    off = 40 * Ny // 256
    probes = []
    for c in (cy - off, cy + off):
        m = np.zeros(shape)
        m[226 * Nx // 256, c - 5:c + 5, cz - 3:cz + 3] = 1.0
        probes.append(("Ez", m))
    return dict(eps=eps, sources=[("z", prof, pulse(n_steps, t0, sigma))], probes=probes)


def _box(i, j0, j1, k0, k1, Ny, Nz, val=1.0):
    jj, kk = np.meshgrid(np.arange(j0, j1), np.arange(k0, k1), indexing="ij")
    ijk = np.stack([np.full(jj.size, i), jj.ravel(), kk.ravel()], 1)
    return {"ijk": ijk, "w": np.full(jj.size, val), "Ny": Ny, "Nz": Nz}


def sparse_points(shape):
    """Source sheet and the two arm probes of the splitter as point lists (grids too large for dense masks)."""
    Nx, Ny, Nz = shape
    cy, cz, off = Ny // 2, Nz // 2, 40 * Ny // 256
    sources = [("z", _box(30 * Nx // 256, cy - 5, cy + 5, cz - 3, cz + 3, Ny, Nz))]
    probes = [("Ez", _box(226 * Nx // 256, c - 5, c + 5, cz - 3, cz + 3, Ny, Nz)) for c in (cy - off, cy + off)]
    return sources, probes


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


# ----------------------------------------------------------------------------- reference arm / CPU legs
def reference_class():
    """The reference's own `fdtd` class if its files travelled (oracle/_ref/, or /root/reference in the build
    container), else None."""
    try:
        from oracle import ref_loader
        if ref_loader.available():
            return ref_loader.load().fdtd
    except Exception:
        pass
    return None


def make_cpu_sim(eps, kind):
    """kind: 'reference' (ceviche/fdtd.py itself) | 'port' (numpy restatement) | 'c' (fused C / OpenMP restatement).
    Returns (step(Jz) -> fields dict, label)."""
    if kind == "reference":
        F = reference_class()(eps, DL, NPML)
        return (lambda Jz: F.forward(Jz=Jz)), "ceviche/fdtd.py (unmodified reference, numpy, 1 thread)"
    if kind == "c":
        from oracle import fdtd_c
        from oracle.fdtd_c import OracleFDTDC
        fdtd_c.set_threads(os.cpu_count() or 1)      # (torchrun exports OMP_NUM_THREADS=1: use every host core anyway)
        sim = OracleFDTDC(eps, DL, NPML)
        return (lambda Jz: sim.step(Jz=Jz)), "oracle/fdtd_c.c (fused C restatement, gcc -O2 -fopenmp)"
    from oracle.fdtd_numpy import OracleFDTD
    sim = OracleFDTD(eps, DL, NPML, materialize=True)
    return (lambda Jz: sim.step(Jz=Jz)), "oracle/fdtd_numpy.py (numpy port of ceviche/fdtd.py, 1 thread)"


def cpu_leg(shape, n_steps, warm, kind, wl=None):
    """Time `n_steps` host time steps of the config-2 workload after `warm` untimed ones.
    Returns (Gcell/s, seconds per time step, sample description, last fields dict)."""
    wl = wl or workload(shape, n_steps + warm)
    step, label = make_cpu_sim(wl["eps"], kind)
    _, prof, wave = wl["sources"][0]
    f = None
    for t in range(warm):
        f = step(prof * wave[t])
    t0 = time.perf_counter()
    for t in range(warm, warm + n_steps):
        f = step(prof * wave[t])
        for key, mask in wl["probes"]:
            np.sum(f[key] * mask)
    el = time.perf_counter() - t0
    cells = shape[0] * shape[1] * shape[2]
    sample = "%dx%dx%d fp64, %d time steps after %d warm-up; %s" % (*shape, n_steps, warm, label)
    return cells * n_steps / el / 1e9, el / n_steps, sample, f


def config2_dict(shape, chunk, arith, w=8):
    """`config` of the N = 1 line, shared by both arms (the reference arm measures a bounded sample of the same workload)."""
    cells = shape[0] * shape[1] * shape[2]
    return {"workload": "config 2: 3-D %dx%dx%d dielectric waveguide splitter, npml 20, Jz sheet source, 2 arm probes" % tuple(shape),
            "time_steps_per_bench_step": chunk, "arith": arith,
            "simulation": "one continuous run of steps x time_steps_per_bench_step time steps from zero fields (pulse at t = 2000)",
            "l2": "state %.0f MB >> 126 MB L2 (inputs larger than L2, no flush needed)" % (cells * w * 9 / 1e6)}


PEER_TEXT = ("direct stores into the neighbour's peer-mapped halo block from inside the half-step kernels "
             "(CUDA IPC over NVLink), arrival counters, time loop in C")


def config3_dict(shape, world, chunk, path_text, w=8):
    """`config` of the N > 1 lines, shared by both arms."""
    planes = shape[0] // world
    local_cells = planes * shape[1] * shape[2]
    return {"workload": "config 3: 3-D %dx%dx%d splitter, npml 20, x-slabs over %d GPUs" % (tuple(shape) + (world,)),
            "halo_exchange": path_text, "time_steps_per_bench_step": chunk, "planes_per_gpu": planes,
            "l2": "per-GPU state %.0f MB >> 126 MB L2" % (local_cells * w * 9 / 1e6)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    shape = (256, 256, 256)
    kind = "reference" if reference_class() is not None else "port"
    try:
        val, sec, sample, _ = cpu_leg(shape, args.steps, args.warmup, kind)
    except MemoryError:
        shape = (128, 128, 128)
        val, sec, sample, _ = cpu_leg(shape, args.steps, args.warmup, kind)
    # the same config as our arm; each of the `steps` bench steps is a bounded sample of it: ONE time step measured,
    # ms_per_step = that time step scaled to the bench step's time_steps_per_bench_step (the value is a rate: unaffected)
    if args.gpus == 1:
        scale = float(args.chunk)
        sample = ("bounded sample: %d + %d single time steps of the config-2 run measured (ms_per_step extrapolates one to the "
                  "%d time steps of a bench step): " % (args.warmup, args.steps, args.chunk)) + sample
    else:     # config 3 (1024 x 1024 x 512) would need ~260 GB in the reference's 39 full-grid arrays: a 256^3 piece of the
        # same splitter stands for it (the numpy step is memory-bound: its per-cell rate does not depend on the extent)
        cells3 = float(args.slab_grid[0]) * args.slab_grid[1] * args.slab_grid[2]
        scale = args.slab_chunk * cells3 / (shape[0] * shape[1] * shape[2])
        sample = ("bounded sample of config 3: single time steps on a %dx%dx%d piece of the splitter (the reference needs "
                  "~260 GB at the full extent); ms_per_step extrapolates by cells and by the %d time steps of a bench step: "
                  % (shape + (args.slab_chunk,))) + sample
    line = {"impl": "reference", "metric": "Gcell-updates/s", "value": val, "unit": "Gcell/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3 * scale,
            "higher_is_better": True, "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": (config2_dict(shape, args.chunk, "f64") if args.gpus == 1 else
                       config3_dict(tuple(args.slab_grid), args.gpus, args.slab_chunk, PEER_TEXT)),
            "cpu_baseline": {"value": val, "unit": "Gcell/s", "cores": 1, "kind": kind, "sample": sample,
                             "host_cores": os.cpu_count(),
                             "note": "the reference is single-process numpy: roll / elementwise passes use 1 of the host cores"},
            "e2e": {"value": val, "unit": "Gcell/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    try:     # beside it: the same algorithm as one fused C pass per half-step on every host core (bit-identical results)
        v2, sec2, sample2, _ = cpu_leg(shape, 6, 1, "c")
        line["cpu_baseline"]["c_openmp_port"] = {"value": v2, "unit": "Gcell/s", "cores": os.cpu_count(), "kind": "port",
                                                 "sample": sample2, "s_per_time_step": sec2}
    except Exception as e:
        line["cpu_baseline"]["c_openmp_port"] = {"unavailable": str(e)[:200]}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- helpers of our arm
def _peaks():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        peaks = {}
    if "hbm_gbs" in peaks:
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def time_kernels(F, shape, reps=50):
    """CUDA-event time of the two half-step kernels alone (ms per launch), on the launch stream."""
    import ctypes as C
    import torch
    from ceviche_b200 import _lib
    plan = F._ensure_plan()
    st, s = F._state(), F._stream()

    def t(fn):
        for _ in range(5):
            fn()
        torch.cuda.synchronize(F.device)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize(F.device)
        return a.elapsed_time(b) / reps
    h = t(lambda: _lib.check(plan.lib.cev_fdtd_step_H(plan.handle, C.byref(st), None, 0, shape[0], s)))
    d = t(lambda: _lib.check(plan.lib.cev_fdtd_step_D(plan.handle, C.byref(st), None, None, None, None, 0, shape[0], s)))
    return h, d


def measure_grid(shape, dtype, n_steps, warm_steps, hbm_peak, opts=(), arith=None):
    """Sustained run() rate + kernel-alone rates of one grid (splitter geometry built on the device)."""
    import torch
    import ceviche_b200
    dev = torch.device("cuda", torch.cuda.current_device())
    w = 8 if dtype == torch.float64 else 4
    cells = shape[0] * shape[1] * shape[2]
    F = ceviche_b200.fdtd(splitter_eps_device(shape, dev), DL, NPML, dtype=dtype, arith=arith)
    for kv in opts:
        k, v = kv.split("=")
        F.set_option(k, int(v))
    srcs, probes = sparse_points(shape)
    # a third probe 8 cells behind the source sheet: the short runs below end before the pulse reaches the arm probes,
    # and a record whose series is identically zero could not notice a wrong answer
    cy, cz = shape[1] // 2, shape[2] // 2
    probes = probes + [("Ez", _box(30 * shape[0] // 256 + 8, cy - 5, cy + 5, cz - 3, cz + 3, shape[1], shape[2]))]
    from ceviche_b200.slab import localize_points
    plane = shape[1] * shape[2]

    def dense(p):      # flat-index point set -> (idx, w) tensors the single-GPU object accepts as a sparse mask
        idx, wts = localize_points(p, 0, shape[0], plane)
        m = torch.zeros(cells, dtype=torch.float64, device=dev)
        m[torch.as_tensor(idx, device=dev)] = torch.as_tensor(wts, device=dev)
        return m.reshape(shape)
    F.prepare([(c, dense(p)) for c, p in srcs], [(k, dense(p)) for k, p in probes])
    wave = torch.as_tensor(pulse(warm_steps + n_steps, t0=0.6 * warm_steps, sigma=warm_steps / 6.0)[:, None]).to(dev)
    F.run(warm_steps, waveforms=wave[:warm_steps])
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    series = F.run(n_steps, waveforms=wave[warm_steps:])
    b.record()
    torch.cuda.synchronize(dev)
    ms = a.elapsed_time(b) / n_steps
    h_ms, d_ms = time_kernels(F, shape, reps=20)
    gcell = cells / ms / 1e6
    out = {"grid": list(shape), "dtype": "f64" if w == 8 else "f32", "time_steps": n_steps, "value": gcell, "unit": "Gcell/s",
           "ms_per_time_step": ms, "series_l2": float(series.norm()),
           "whole_step": {"achieved": gcell * B_ALG_WORDS * w, "frac": gcell * B_ALG_WORDS * w / hbm_peak,
                          "frac_of_nominal_8TBs": gcell * B_ALG_WORDS * w / NOMINAL_GBS},
           "step_H": {"ms_per_launch": h_ms, "achieved": cells * H_KERNEL_WORDS * w / h_ms / 1e6,
                      "frac": cells * H_KERNEL_WORDS * w / h_ms / 1e6 / hbm_peak},
           "step_D": {"ms_per_launch": d_ms, "achieved": cells * D_KERNEL_WORDS * w / d_ms / 1e6,
                      "frac": cells * D_KERNEL_WORDS * w / d_ms / 1e6 / hbm_peak}}
    del F
    torch.cuda.empty_cache()
    return out


def oracle_parity(shape, dtype, n_steps, opts=()):
    """The bench grid against the CPU oracle (fused C restatement, bit-identical to the numpy port and to the
    reference) on an early pulse, so that the fields being compared are not zero: worst rel-L2 over the nine fields
    and over the probe series.  Also times the oracle: the multi-core CPU figure."""
    import torch
    import ceviche_b200
    from oracle.fdtd_numpy import FIELD_KEYS, rel_l2
    wl = workload(shape, n_steps, t0=n_steps / 3.0, sigma=n_steps / 10.0)
    # probes where the early pulse already is: two patches next to the source sheet
    Nx, Ny, Nz = shape
    cy, cz, ix = Ny // 2, Nz // 2, 30 * Nx // 256
    probes = []
    for di in (2, 5):
        m = np.zeros(shape)
        m[ix + di, cy - 5:cy + 5, cz - 3:cz + 3] = 1.0
        probes.append(("Ez", m))
    wl["probes"] = probes
    t0 = time.perf_counter()
    v, sec, sample, f_cpu = cpu_leg(shape, n_steps, 0, "c", wl=wl)
    # (cpu_leg does not keep the series: redo the probe sums on the final fields only; the GPU series is checked
    # against a second short oracle run below)
    F = ceviche_b200.fdtd(wl["eps"], DL, NPML, dtype=dtype)
    for kv in opts:
        k, vv = kv.split("=")
        F.set_option(k, int(vv))
    series = F.run(n_steps, wl["sources"], wl["probes"]).cpu().numpy()
    worst = 0.0
    for k in FIELD_KEYS:
        worst = max(worst, rel_l2(F.fields[k].cpu().numpy(), f_cpu[k]))
    last = [float(np.sum(f_cpu[key] * mask)) for key, mask in wl["probes"]]
    s_err = max(abs(series[-1, p] - last[p]) / (abs(last[p]) + 1e-300) for p in range(len(last)))
    tol = 1e-10 if dtype == torch.float64 else 1e-5
    res = {"vs": "CPU oracle (oracle/fdtd_c.c, bit-identical to the numpy port of ceviche/fdtd.py)", "grid": list(shape),
           "time_steps": n_steps, "worst_field_rel_l2": worst, "last_probe_sample_rel_err": s_err, "tolerance": tol,
           "field_l2": float(np.sqrt(sum(np.sum(f_cpu[k] ** 2) for k in ("Ex", "Ey", "Ez")))),
           "ok": bool(worst <= tol and s_err <= 100 * tol), "seconds": time.perf_counter() - t0}
    cpu = {"value": v, "unit": "Gcell/s", "cores": os.cpu_count(), "kind": "port", "sample": sample, "s_per_time_step": sec}
    del F
    torch.cuda.empty_cache()
    return res, cpu


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import ceviche_b200

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
    hbm_peak, peak_src = _peaks()

    if world > 1:
        return run_slabs(args, dist, dev, local, rank, world, dtype, w, hbm_peak, peak_src)

    shape = tuple(args.grid)
    cells = shape[0] * shape[1] * shape[2]
    n_total = chunk * args.steps
    wl = workload(shape, max(n_total, chunk * max(args.warmup, 1)))
    F = ceviche_b200.fdtd(wl["eps"], DL, NPML, dtype=dtype, arith=args.arith)
    for kv in args.opt:
        k, v = kv.split("=")
        F.set_option(k, int(v))
    srcs, probes = wl["sources"], wl["probes"]

    def sync():
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput: ONE simulation of steps * chunk time steps from zero fields -----------
    wave_dev = torch.as_tensor(np.stack([s_[2] for s_ in srcs], 1)).to(dev)
    F.prepare(srcs, probes)        # profiles / masks uploaded once: inputs resident in HBM
    for q in range(args.warmup):
        F.run(chunk, waveforms=wave_dev[q * chunk:(q + 1) * chunk])
    F.initialize_fields()          # the timed region is the config's run from t = 0
    sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    series = []
    with ClockSampler(local) as clk:
        ev0.record()
        for q in range(args.steps):
            series.append(F.run(chunk, waveforms=wave_dev[q * chunk:(q + 1) * chunk]))
        ev1.record()
        sync()
    ms = ev0.elapsed_time(ev1)
    ms_per_step = ms / args.steps
    value = cells * chunk * args.steps / (ms * 1e-3) / 1e9
    series = torch.cat(series)
    launches = args.steps * (chunk * 2 + 4)     # per run(): 2 half-step kernels per time step (sources and probes ride
                                                # inside them) + the trailing probe launch + 3 E-materialisation launches
    sim_check = {"time_steps": int(series.shape[0]), "probe_series_l2": float(series.norm()),
                 "probe_series_peak": float(series.abs().max()), "finite": bool(torch.isfinite(series).all()),
                 "peak_at_time_step": int(series.abs().amax(1).argmax())}

    # ---- per-kernel timing of the two half-step kernels (events on the launch stream) ----
    h_ms, d_ms = time_kernels(F, shape)
    h_gbs = cells * H_KERNEL_WORDS * w / (h_ms * 1e-3) / 1e9
    d_gbs = cells * D_KERNEL_WORDS * w / (d_ms * 1e-3) / 1e9
    step_gbs = value * B_ALG_WORDS * w
    traffic, traffic_src = None, None
    prof_json = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.isfile(prof_json):
        try:
            rec = json.load(open(prof_json)).get("%s_%dx%dx%d" % ((args.dtype,) + shape), {})
            traffic, traffic_src = rec.get("step_H_bytes"), "profiles/ncu_traffic.json (%s)" % rec.get("capture", "ncu --set full")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "step_H (H half-step: D, 1/eps, H in; H out = 12 words/cell)",
                "achieved": h_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": h_gbs / hbm_peak, "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peak_src, "ms_per_launch": h_ms,
                "step_D": {"achieved": d_gbs, "frac": d_gbs / hbm_peak, "ms_per_launch": d_ms, "words_per_cell": D_KERNEL_WORDS},
                "whole_step": {"bytes_per_cell_update": B_ALG_WORDS * w, "achieved": step_gbs, "frac": step_gbs / hbm_peak,
                               "frac_of_nominal_8TBs": step_gbs / NOMINAL_GBS}}

    # ---- end to end through the public API with host buffers ---------------------------
    eps_host = torch.as_tensor(wl["eps"]).to(dtype).pin_memory()
    wave_host = torch.as_tensor(np.stack([s_[2] for s_ in srcs], 1)[:chunk * 3]).pin_memory()
    geo = [(c, p) for c, p, _ in srcs]
    h2d = eps_host.numel() * eps_host.element_size() + chunk * 8 + sum(p.nbytes for _, p in geo) + sum(m.nbytes for _, m in probes)
    d2h = chunk * len(probes) * 8

    def e2e_step(q):
        F.eps_r = eps_host.to(dev, non_blocking=True)      # upload + Yee averaging + 1/eps + field reset
        out = F.run(chunk, geo, probes, waveforms=wave_host[q * chunk:(q + 1) * chunk].to(dev, non_blocking=True))
        return out.cpu()
    if args.no_e2e:
        e2e_s, e2e_val = float("nan"), None
    else:
        for _ in range(max(1, args.warmup // 2)):
            e2e_step(0)
        sync()
        n_e2e = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for q in range(n_e2e):
            e2e_step(q)
        sync()
        e2e_s = (time.perf_counter() - t0) / n_e2e
        e2e_val = cells * chunk / e2e_s / 1e9
    del F
    torch.cuda.empty_cache()

    # ---- parity at the bench grid + CPU baselines (the reference algorithm on the host cores, bounded samples) ----
    parity, cpu = None, None
    if not args.no_cpu:
        pshape = shape if cells <= 256 ** 3 else (256, 256, 256)
        try:
            parity, cpu_c = oracle_parity(pshape, dtype, 24, args.opt)
        except Exception as e:      # no gcc / OpenMP on the box: the numpy figure below stands alone
            parity, cpu_c = {"unavailable": str(e)[:200]}, {"unavailable": str(e)[:200]}
        kind = "reference" if reference_class() is not None else "port"
        try:
            v, sec, sample, _ = cpu_leg(pshape, 2, 1, kind)
        except MemoryError:
            v, sec, sample, _ = cpu_leg((128, 128, 128), 4, 1, kind)
        cpu = {"value": v, "unit": "Gcell/s", "cores": 1, "kind": kind, "sample": sample,
               "host_cores": os.cpu_count(), "s_per_time_step": sec, "c_openmp_port": cpu_c}

    # ---- the other dtype at this grid, the north_star target grid, and the same-workload anchor of the 1 -> 8 curve ----
    other, anchor = None, None
    if not args.no_extra:
        other = []
        odt = torch.float32 if dtype == torch.float64 else torch.float64
        for shp, dt_, n, wm in ((shape, odt, 1500, 300), ((512, 512, 512), torch.float64, 150, 40), ((512, 512, 512), torch.float32, 300, 60)):
            try:
                other.append(measure_grid(shp, dt_, n, wm, hbm_peak, args.opt))
            except Exception as e:
                other.append({"grid": list(shp), "unavailable": str(e)[:200]})
        try:
            anchor = measure_grid(tuple(args.slab_grid), dtype, 60, 20, hbm_peak, args.opt)
            anchor["note"] = "BASELINE config 3's grid on one GPU: divide the N-GPU values by this for same-workload scaling"
        except Exception as e:
            anchor = {"grid": list(args.slab_grid), "unavailable": str(e)[:200]}

    line = {"metric": "Gcell-updates/s", "value": value, "unit": "Gcell/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": config2_dict(shape, chunk, args.arith or args.dtype, w),
            "simulation_check": sim_check, "parity_check": parity,
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": "Gcell/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_s * 1e3},
            "other": other, "scale_anchor": anchor,
            "gpu_launches": int(launches), "clocks": clk.summary()}
    print(json.dumps(line))
    if parity is not None and parity.get("ok") is False:
        sys.exit(3)


def run_slabs(args, dist, dev, local, rank, world, dtype, w, hbm_peak, peak_src):
    """N > 1: BASELINE config 3 -- 1024 x 1024 x 512 (npml 20) cut into x-slabs, one per GPU.  Total work is fixed
    as N grows (strong scaling)."""
    import torch
    import ceviche_b200
    from ceviche_b200.slab import XPML_PLANE_COST, partition
    shape = tuple(args.slab_grid)
    Nx, Ny, Nz = shape
    cells = Nx * Ny * Nz
    chunk = args.slab_chunk
    devices = list(range(world))
    balance = XPML_PLANE_COST if args.balance < 0 else args.balance      # (default 0: equal plane counts)

    # ---- parity first: config 3's parity grid on the N slabs against one GPU, bit for bit --------------------
    from ceviche_b200.slab import XPML_PLANE_COST as _cost
    parity = slab_parity(dist, dev, rank, world, dtype, args.opt, _cost if args.balance < 0 else args.balance)

    parts = partition(Nx, world, NPML[0], balance)      # cost-balanced: the end ranks carry the x-PML and get fewer planes
    lo, hi = parts[rank]
    sources, probes = sparse_points(shape)
    n_total = chunk * (args.steps + args.warmup)
    wave = pulse(n_total, t0=0.5 * chunk, sigma=chunk / 8.0)[:, None]
    sim = ceviche_b200.fdtd(splitter_eps_device(shape, dev, lo, hi), DL, NPML, dtype=dtype, devices=devices, global_shape=shape,
                            balance=balance)
    for kv in args.opt:
        k, v = kv.split("=")
        sim.set_option(k, int(v))
    sim.prepare(sources, probes)
    wave_dev = torch.as_tensor(wave).to(dev)

    def sync():
        torch.cuda.synchronize(dev)
        dist.barrier()

    for q in range(args.warmup):
        sim.run(chunk, waveforms=wave_dev[q * chunk:(q + 1) * chunk])
    sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    series = []
    with ClockSampler(local) as clk:
        ev0.record()
        for q in range(args.warmup, args.warmup + args.steps):
            series.append(sim.run(chunk, waveforms=wave_dev[q * chunk:(q + 1) * chunk]))
        ev1.record()
        sync()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    value = cells * chunk * args.steps / (ms * 1e-3) / 1e9
    series = torch.cat(series)

    # per-kernel timing of the local H half-step (whole local slab, halo in place)
    h_ms = sim.time_local_step_H(reps=20)
    local_cells = (hi - lo) * Ny * Nz
    h_gbs = local_cells * H_KERNEL_WORDS * w / (h_ms * 1e-3) / 1e9
    step_gbs = value * B_ALG_WORDS * w / world
    path = sim.slab_path()

    # end to end: host eps slab -> new simulator -> run -> series on the host (averaged over 3 chunks)
    eps_host = splitter_eps_device(shape, "cpu", lo, hi).pin_memory()
    wave_host = torch.as_tensor(wave[:3 * chunk]).pin_memory()
    del sim
    torch.cuda.empty_cache()
    dist.barrier()
    t0 = time.perf_counter()
    sim2 = ceviche_b200.fdtd(eps_host.to(dev, non_blocking=True), DL, NPML, dtype=dtype, devices=devices, global_shape=shape,
                             balance=balance)
    sim2.prepare(sources, probes)
    for q in range(3):
        sim2.run(chunk, waveforms=wave_host[q * chunk:(q + 1) * chunk].to(dev, non_blocking=True)).cpu()
    sync()
    e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e) / 3
    if rank == 0:
        line = {"metric": "Gcell-updates/s", "value": value, "unit": "Gcell/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                "config": config3_dict(shape, world, chunk, path, w),
                "parity_check": parity,
                "simulation_check": {"time_steps": int(series.shape[0]), "probe_series_l2": float(series.norm()),
                                     "finite": bool(torch.isfinite(series).all())},
                "roofline": {"bound": "hbm", "kernel": "step_H on the local slab (12 words/cell)", "achieved": h_gbs, "peak": hbm_peak,
                             "unit": "GB/s", "frac": h_gbs / hbm_peak, "traffic": None, "peak_source": peak_src, "ms_per_launch": h_ms,
                             "whole_step_per_gpu": {"achieved": step_gbs, "frac": step_gbs / hbm_peak,
                                                    "frac_of_nominal_8TBs": step_gbs / NOMINAL_GBS}},
                "cpu_baseline": None,
                "e2e": {"value": cells * chunk / e2e_s / 1e9, "unit": "Gcell/s",
                        "h2d_bytes_per_step": int(eps_host.numel() * 8 * world / 3 + chunk * 8 * world),
                        "d2h_bytes_per_step": int(chunk * len(probes) * 8 * world), "ms_per_step": e2e_s * 1e3,
                        "note": "3 chunks after constructing the simulator from host buffers (construction inside the timed region)"},
                "gpu_launches": int(args.steps * chunk * 2), "clocks": clk.summary()}
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
    if not parity["bitwise_equal"]:
        sys.exit(3)


def slab_parity(dist, dev, rank, world, dtype, opts=(), balance=0.0):
    """BASELINE config 3's parity grid (256 x 128 x 64, npml 20) for 30 steps on the N slabs and on one GPU
    (every rank computes the single-GPU answer itself): all nine fields and the probe series must agree bit for bit."""
    import torch
    import ceviche_b200
    from ceviche_b200.slab import partition
    # (fp32: Nz = 128, so that the grid takes the same peer-memory halo path as the timed one)
    shape, steps = ((256, 128, 64) if dtype == torch.float64 else (256, 128, 128)), 30
    rng = np.random.default_rng(11)
    eps = 1 + 2 * rng.random(shape)
    prof = rng.random(shape) * (rng.random(shape) < 0.02)
    one = np.zeros(shape)
    one[0, 1, 2] = 1.0
    t = np.arange(steps)
    wf = np.stack([np.exp(-(t - 10.0) ** 2 / 18.0) * np.cos(0.7 * t), np.exp(-(t - 7.0) ** 2 / 8.0)], 1)
    sources = [("z", prof), ("y", one)]
    probes = [("Ez", rng.random(shape)), ("Hy", (rng.random(shape) < 0.01) * 1.0), ("Dx", rng.random(shape))]
    keys = ("Ex", "Ey", "Ez", "Dx", "Dy", "Dz", "Hx", "Hy", "Hz")

    one_gpu = ceviche_b200.fdtd(eps, DL, NPML, dtype=dtype, device=dev)
    for kv in opts:
        k, v = kv.split("=")
        one_gpu.set_option(k, int(v))
    s1 = one_gpu.run(steps, sources, probes, waveforms=wf)
    f1 = {k: one_gpu.fields[k].clone() for k in keys}
    del one_gpu

    parts = partition(shape[0], world, NPML[0], balance)       # the same (cost-balanced, uneven) cut as the timed run
    lo, hi = parts[rank]
    eps_local = np.concatenate([eps[(lo - 1) % shape[0]][None], eps[lo:hi]], 0)
    sim = ceviche_b200.fdtd(eps_local, DL, NPML, dtype=dtype, devices=list(range(world)), global_shape=shape, balance=balance)
    for kv in opts:
        k, v = kv.split("=")
        sim.set_option(k, int(v))
    half = steps // 2
    sN = torch.cat([sim.run(half, sources, probes, waveforms=wf[:half]), sim.run(steps - half, waveforms=wf[half:])])
    ok = True
    for k in keys:       # every rank checks its own slab of every field
        ok = ok and bool(torch.equal(sim.local_fields[k], f1[k][lo:hi]))
    series_err = float((sN - s1).abs().max() / s1.abs().max())
    flag = torch.tensor([1.0 if (ok and series_err <= 1e-11) else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res = {"ranks": world, "grid": list(shape), "planes_per_rank": [b - a for a, b in parts], "time_steps": steps,
           "halo_exchange": sim.slab_path(),
           "bitwise_equal": bool(flag.item() == 1.0), "fields_compared": 9, "series_max_rel_diff": series_err,
           "vs": "the same grid stepped on one GPU (itself <= 1e-10 from the CPU oracle: tests/test_gpu_slab.py)"}
    del sim
    torch.cuda.empty_cache()
    return res


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
    ap.add_argument("--balance", type=float, default=0.0,
                    help="x-PML plane cost of the slab partition (0: equal plane counts, the default; -1: the measured plane cost "
                         "0.25 -- cost-balanced slabs, measured neutral on 8 B200)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline + oracle parity legs")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (kernel tuning runs)")
    ap.add_argument("--no-extra", action="store_true", help="skip the other-dtype / 512^3 / scale-anchor measurements")
    ap.add_argument("--opt", action="append", default=[], help="plan option name=value (repeatable)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    main()
