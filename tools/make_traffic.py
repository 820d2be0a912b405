#!/usr/bin/env python
"""profiles/traffic.json from the `ncu --set full` captures of the scan kernel (gpurun_out/prof_mma_q*.ncu-rep):
dram__bytes_read.sum + dram__bytes_write.sum of ONE launch, stamped with the HEAD the captures were taken at.
bench.py reports it as `roofline.traffic` for the matching configuration."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def metric(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", name], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    col = [i for i, h in enumerate(hdr) if h == name]
    if not col:
        return None
    units = rows[1][col[0]]
    v = float(rows[2][col[0]].replace(",", ""))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units, 1.0)
    return v * scale


def main():
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    out = {"captured_at": f"HEAD {head}, ncu --set full --clock-control none, one launch of sim_mma_kernel, 1 GPU, "
                          "1,007,323 x 2048 bf16 (tools/gpu_profiles.sh)"}
    for nq in (1, 70, 1024, 4096):
        rep = os.path.join(ROOT, "gpurun_out", f"prof_mma_q{nq}.ncu-rep")
        if not os.path.exists(rep):
            continue
        r, w = metric(rep, "dram__bytes_read.sum"), metric(rep, "dram__bytes_write.sum")
        if r is None or w is None:
            continue
        out[f"nq{nq}"] = {"dram_bytes_read": r, "dram_bytes_write": w, "dram_bytes_per_launch": r + w,
                          "source": f"gpurun_out/prof_mma_q{nq}.ncu-rep -> profiles/r02_ncu_mma_q{nq}.txt"}
    json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    sys.exit(main())
