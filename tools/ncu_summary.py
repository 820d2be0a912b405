#!/usr/bin/env python
"""Turn ncu output (gpurun_out/) into the small text summaries committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_q70.csv profiles/r01_launches_q70.txt
    python tools/ncu_summary.py full gpurun_out/prof_mma_q70.ncu-rep profiles/r01_ncu_mma_q70.txt [kernel-regex]

`launches`: per-kernel launch count / mean / total of gpu__time_duration.sum for librir kernels (rir::*), and their
share of one step.  `full`: the metrics B200_PROFILING.md names, from `ncu -i <rep> --page raw --csv`.
"""
from __future__ import annotations

import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.max",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum.per_second",
    "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "sm__inst_executed_pipe_uniform.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__cluster_size",
    "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_barrier",
    "smsp__pcsamp_warps_issue_stalled_membar",
]


def launches(src: str, dst: str) -> None:
    rows = [r for r in csv.reader(open(src, newline="")) if len(r) > 5]
    hdr = next(i for i, r in enumerate(rows) if r[0] == "ID")
    h = rows[hdr]
    ki, vi, gi, bi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
    data = rows[hdr + 1:]
    agg: "OrderedDict[str, list]" = OrderedDict()
    for r in data:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
        if "rir::" not in r[ki]:
            name = "(torch / other)"
        us = float(r[vi].replace(",", "")) / 1e3
        a = agg.setdefault(name, [0, 0.0, r[gi], r[bi], []])
        a[0] += 1
        a[1] += us
        a[4].append(us)
    ours = {k: v for k, v in agg.items() if k != "(torch / other)"}
    tot = sum(v[1] for v in ours.values())
    with open(dst, "w") as f:
        f.write(f"# source: {src} (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)\n")
        f.write(f"# librir kernels only; share = kernel total / all librir kernel time ({tot:.1f} us)\n")
        f.write(f"{'kernel':48s} {'launches':>8s} {'mean_us':>9s} {'min_us':>9s} {'max_us':>9s} {'total_us':>10s} {'share':>7s}  grid block\n")
        for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:48]:48s} {v[0]:8d} {v[1] / v[0]:9.1f} {min(v[4]):9.1f} {max(v[4]):9.1f} {v[1]:10.1f} {v[1] / tot:7.3f}  {v[2]} {v[3]}\n")
        if "(torch / other)" in agg:
            v = agg["(torch / other)"]
            f.write(f"# torch / other kernels in the capture (data generation, copies): {v[0]} launches, {v[1]:.1f} us\n")
    print(open(dst).read())


def full(src: str, dst: str, kernel_re: str = ".") -> None:
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    ki = h.index("Kernel Name")
    with open(dst, "w") as f:
        f.write(f"# source: {src} (ncu --set full --clock-control none --import-source on)\n")
        for r in rows[2:]:
            if not re.search(kernel_re, r[ki]):
                continue
            f.write(f"\n## {r[ki][:120]}\n")
            for key in KEYS:
                for i, c in enumerate(h):
                    if c == key or (key.startswith("smsp__pcsamp") and c.startswith(key)):
                        f.write(f"{c:86s} {r[i]:>20s} {units[i]}\n")
    print(open(dst).read())


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else ".")
