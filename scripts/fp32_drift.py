"""fp32 drift of the shipped kernels against the fp64 GPU run: rel-L2 of the nine fields and of
the probe series at 1k / 2k / 5k / 10k time steps, fp32 storage with fp32 arithmetic and with fp64 arithmetic
(arith='f64').  Grids: config 2 scaled to 96^3 (tests/golden c2_96 geometry, pulse at t0 = 300) and the real config 2
(256^3, pulse at t0 = 2000: bench.py's workload).      python scripts/fp32_drift.py > gpurun_out/fp32_drift.json"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ceviche_b200  # noqa: E402
import bench  # noqa: E402
from oracle import cases  # noqa: E402

KEYS = ("Ex", "Ey", "Ez", "Dx", "Dy", "Dz", "Hx", "Hy", "Hz")
MILESTONES = (1000, 2000, 5000, 10000)


def run(eps, sources, probes, wave, dtype, arith, npml=(20, 20, 20)):
    F = ceviche_b200.fdtd(eps, cases.DL, list(npml), dtype=dtype, arith=arith)
    F.prepare(sources, probes)
    wave = torch.as_tensor(wave).cuda()
    snaps, series, t = {}, [], 0
    for m in MILESTONES:
        series.append(F.run(m - t, waveforms=wave[t:m]))
        t = m
        snaps[m] = {k: F.fields[k].double().clone() for k in KEYS}
    return snaps, torch.cat(series).double()


def rel(a, b):
    n = float(torch.linalg.vector_norm(b))
    return float(torch.linalg.vector_norm(a - b)) / n if n > 0 else float(torch.linalg.vector_norm(a))


def table(name, eps, sources, probes, wave, npml=(20, 20, 20)):
    ref_snaps, ref_series = run(eps, sources, probes, wave, torch.float64, None, npml)
    out = {"grid": list(eps.shape), "case": name, "rows": []}
    for arith in ("f32", "f64"):
        snaps, series = run(eps, sources, probes, wave, torch.float32, arith, npml)
        for m in MILESTONES:
            allf = torch.cat([snaps[m][k].ravel() for k in KEYS]), torch.cat([ref_snaps[m][k].ravel() for k in KEYS])
            e_only = torch.cat([snaps[m][k].ravel() for k in KEYS[:3]]), torch.cat([ref_snaps[m][k].ravel() for k in KEYS[:3]])
            out["rows"].append({"storage": "f32", "arith": arith, "time_steps": m,
                                "fields_rel_l2": rel(*allf), "E_rel_l2": rel(*e_only),
                                "worst_field_rel_l2": max(rel(snaps[m][k], ref_snaps[m][k]) for k in KEYS),
                                "series_rel_l2_up_to_here": rel(series[:m], ref_series[:m]),
                                "series_rel_l2_per_probe": [rel(series[:m, q], ref_series[:m, q]) for q in range(series.shape[1])],
                                "E_energy_vs_peak": float(sum(ref_snaps[m][k].pow(2).sum() for k in KEYS[:3])) /
                                max(float(sum(ref_snaps[q][k].pow(2).sum() for k in KEYS[:3])) for q in MILESTONES)})
        del snaps
        torch.cuda.empty_cache()
    return out


def main():
    res = []
    case = cases.scaled_case("c2_96")
    srcs = [(c, p) for c, p, _ in case["sources"]]
    w = np.stack([np.concatenate([wv, np.zeros(MILESTONES[-1] - len(wv))]) for _, _, wv in case["sources"]], 1)
    res.append(table("config 2 scaled to 96^3 (pulse t0=300, sigma=60; zero drive after step 2000)", case["eps"], srcs, case["probes"], w))
    res.append(table("the same with a 12-cell PML (the golden fixture's)", case["eps"], srcs, case["probes"], w, npml=[12, 12, 12]))
    shape = (256, 256, 256)
    wl = bench.workload(shape, MILESTONES[-1])
    res.append(table("config 2: 256^3 splitter, pulse t0=2000 sigma=100 (bench.py workload)", wl["eps"],
                     [(c, p) for c, p, _ in wl["sources"]], wl["probes"], np.stack([s[2] for s in wl["sources"]], 1)))
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
