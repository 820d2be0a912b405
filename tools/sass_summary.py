#!/usr/bin/env python
"""Opcode histogram per kernel of librir.so (cuobjdump -sass): the evidence that the hot kernels are tcgen05 / TMEM /
TMA code (B200_PROFILING.md "What proves a Blackwell-native kernel").  Runs on the CPU box.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "research_image_retrieval_b200", "lib", "librir.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCATOM",
         "SYNCS", "HMMA", "HGMMA", "LDGSTS", "LDG", "STG", "LDS", "STS", "ATOM", "RED", "ATOMS", "BAR", "SHFL", "FFMA",
         "MUFU", "ACQBULK", "CCTL", "ERRBAR", "MEMBAR", "ELECT", "NANOSLEEP"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            kernels[cur][op.split(".")[0]] += 1
            if op.startswith(("UTC", "UTMA", "LDTM", "UBLKCP")):
                kernels[cur]["full:" + op] += 1
    res = {}   # registers / static shared / local (spill) bytes per kernel: cuobjdump --dump-resource-usage
    ru = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True, check=True).stdout
    fn = None
    for line in ru.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and fn:
            res[fn] = tuple(int(x) for x in m.groups())
            fn = None
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a), sources at HEAD {head} (+ working tree)")
    print("# per kernel: instruction count, then the opcodes that matter (tcgen05.mma = UTC*MMA, tcgen05.ld = LDTM,")
    print("# TMA = UTMALDG / UBLKCP, tcgen05.commit = UTCBAR; HMMA would be a legacy mma.sync path: none expected);")
    print("# regs / stack / static shared / local bytes from cuobjdump --dump-resource-usage (local > 0 = spills)")
    for (name, c), dm in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", dm)
        ops = " ".join(f"{k}={c[k]}" for k in WATCH if c[k])
        r = res.get(name)
        usage = f"  regs={r[0]} stack={r[1]} smem_static={r[2]} local={r[3]}" if r else ""
        print(f"\n{short}  [{c['_total']} instr]{usage}\n    {ops}")
        full = sorted((k[5:], v) for k, v in c.items() if k.startswith("full:"))
        if full:
            print("    " + "  ".join(f"{k} x{v}" for k, v in full))
    tot = collections.Counter()
    for c in kernels.values():
        tot.update({k: v for k, v in c.items() if not k.startswith(("full:", "_"))})
    print("\n# whole library: " + " ".join(f"{k}={tot[k]}" for k in WATCH if tot[k]))
    print(f"# legacy tensor-core opcodes (HMMA / HGMMA / IMMA): {tot['HMMA'] + tot['HGMMA'] + tot['IMMA']}")


if __name__ == "__main__":
    sys.exit(main())
