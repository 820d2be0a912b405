#!/usr/bin/env python
"""Event timeline of one fused tcgen05 scan (development tool; rir_profile_timeline in include/rir.h).

    python tools/timeline.py [--nq 70] [--n 125916] [--d 2048] [--k 100]

Prints, per event kind and round, min / median / max over the CTAs of the time since the first CTA started (us)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import research_image_retrieval_b200 as rir  # noqa: E402

NAMES = {0: "cta start", 1: "first db TMA", 2: "mma start", 3: "mma issued", 4: "acc ready", 5: "epilogue done",
         6: "threshold begin", 7: "grid barrier passed", 8: "thresholds published", 9: "cta done"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nq", type=int, default=70)
    ap.add_argument("--n", type=int, default=125916)
    ap.add_argument("--d", type=int, default=2048)
    ap.add_argument("--k", type=int, default=100)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(1)
    X = torch.randn(a.n, a.d, generator=gen, device=dev)
    X = (X / X.norm(dim=1, keepdim=True)).to(torch.bfloat16)
    Q = torch.randn(a.nq, a.d, generator=gen, device=dev)
    Q = (Q / Q.norm(dim=1, keepdim=True)).to(torch.bfloat16)
    db = rir.Database(X, None, "bf16")
    for _ in range(5):
        db.search(Q, None, a.k)
    cap = 1 << 16
    buf = torch.zeros(1 + 2 * cap, dtype=torch.int64, device=dev)
    lib = rir.load()
    lib.rir_profile_timeline(buf.data_ptr(), cap)
    db.search(Q, None, a.k)
    torch.cuda.synchronize()
    lib.rir_profile_timeline(None, 0)
    h = buf.cpu().numpy().view(np.uint64)
    n = int(min(h[0], cap))
    meta, t = h[1:1 + 2 * n:2], h[2:2 + 2 * n:2].astype(np.int64)
    cta, ev, rd = (meta >> np.uint64(32)).astype(np.int64), ((meta >> np.uint64(16)) & np.uint64(0xffff)).astype(np.int64), \
        (meta & np.uint64(0xffff)).astype(np.int64)
    t0 = t[ev == 0].min()
    print(f"{n} events, {len(np.unique(cta))} CTAs, nq={a.nq} n={a.n} d={a.d}; times in us since the first CTA start")
    print(f"{'event':24s} {'round':>5s} {'ctas':>5s} {'min':>8s} {'median':>8s} {'max':>8s}")
    for e in sorted(NAMES):
        for r in sorted(np.unique(rd[ev == e])):
            sel = (ev == e) & (rd == r)
            us = (t[sel] - t0) / 1e3
            print(f"{NAMES[e]:24s} {r:5d} {sel.sum():5d} {us.min():8.1f} {np.median(us):8.1f} {us.max():8.1f}")


if __name__ == "__main__":
    main()
