"""Summarise `ncu --set full` reports into profiles/ncu_summary.json entries.

  python tools/ncu_summarize.py KEY=path.ncu-rep:nodes:steps_per_launch [...]  > entries.json

Run where `ncu` is on PATH (the build container: reading a report needs no GPU).  For every report the first
profiled launch is used; `nodes` is the number of lattice nodes one launch updates, `steps_per_launch` how many
time steps it advances them (1 for k_lbm, 2 for k_lbm2)."""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "gpu_time_ms_under_ncu",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct_of_ncu_peak",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct_of_ncu_peak",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_active_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio_throttle_per_issue",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard_per_issue",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier_per_issue",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "msecond": 1.0, "usecond": 1e-3, "second": 1e3, "ms": 1.0, "us": 1e-3}


def summarise(path, nodes, steps):
    if path.endswith(".csv"):  # a raw page already exported (profiles/*_ncu_full_*.csv)
        raw = open(path).read()
    else:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, first = rows[0], rows[1], rows[2]
    out = {"kernel": first[head.index("Kernel Name")], "nodes_per_launch": nodes, "steps_per_launch": steps}
    for k, u, v in zip(head, units, first):
        if k in WANT:
            x = float(v.replace(",", ""))
            out[WANT[k]] = x * SCALE.get(u, 1.0)
    out["dram_bytes_per_lup"] = (out["dram_bytes_read"] + out["dram_bytes_write"]) / (nodes * steps)
    return out


if __name__ == "__main__":
    res = {}
    for arg in sys.argv[1:]:
        key, rest = arg.split("=", 1)
        path, nodes, steps = rest.rsplit(":", 2)
        res[key] = summarise(path, int(nodes), int(steps))
        res[key]["source"] = path
    json.dump(res, sys.stdout, indent=1)
    print()
