"""Summarise .ncu-rep captures (ncu --set full) as JSON: per kernel launch, the metrics the roofline discussion uses.
    python scripts/ncu_summary.py gpurun_out/a.ncu-rep [b.ncu-rep ...] > profiles/<name>.json"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "time_ns",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__occupancy_limit_registers": "occ_limit_regs",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct": "issue_active_pct",
    "sm__inst_executed.sum": "inst_executed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard_per_issue",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier_per_issue",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard_per_issue",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_throttle_per_issue",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait_per_issue",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle_per_issue",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio": "stall_membar_per_issue",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
}


def summarise(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6,
             "Gbyte": 1e9, "Tbyte": 1e12}          # times -> ns, sizes -> bytes
    res = []
    for r in rows[2:]:
        rec = {"kernel": r[head.index("Kernel Name")][:120]}
        for m, name in WANT.items():
            if m in head:
                v = r[head.index(m)].replace(",", "")
                try:
                    rec[name] = float(v) * scale.get(units[head.index(m)], 1.0)
                except ValueError:
                    rec[name] = v
        if "time_ns" in rec and "dram_read_bytes" in rec:
            rec["dram_bytes"] = rec["dram_read_bytes"] + rec["dram_write_bytes"]
            rec["dram_GBps"] = rec["dram_bytes"] / rec["time_ns"]
        res.append(rec)
    return res


if __name__ == "__main__":
    print(json.dumps({p: summarise(p) for p in sys.argv[1:]}, indent=1))
