#!/usr/bin/env python
"""Where does the host-buffer (e2e) step spend its time?  Times, per step with a stream sync each step:
   (a) resident search only, (b) H2D + pack only, (c) the full rir_search_host call."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import research_image_retrieval_b200 as rir  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=125916)
ap.add_argument("--steps", type=int, default=300)
a = ap.parse_args()
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1)
d, nq, k = 2048, 70, 100
X = torch.empty(a.n, d, device=dev, dtype=torch.bfloat16)
for lo in range(0, a.n, 65536):
    blk = torch.randn(min(65536, a.n - lo), d, generator=gen, device=dev)
    X[lo:lo + blk.shape[0]] = (blk / blk.norm(dim=1, keepdim=True)).to(torch.bfloat16)
qh = torch.randn(nq, d)
qh = (qh / qh.norm(dim=1, keepdim=True)).pin_memory()
db = rir.Database(X, None, "bf16")
qr, qs = db.pack_queries(qh.to(dev))
out = (torch.empty((nq, k), dtype=torch.float32).pin_memory(), torch.empty((nq, k), dtype=torch.int32).pin_memory())
st = torch.cuda.current_stream()


def run(name, fn, sync_each):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(a.steps):
        fn()
        if sync_each:
            st.synchronize()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"{name:44s} device {e0.elapsed_time(e1) / a.steps * 1e3:8.1f} us/step   host {(t1 - t0) / a.steps * 1e6:8.1f} us/step")


run("resident search, no sync", lambda: db.search(qr, qs, k), False)
run("resident search, sync each step", lambda: db.search(qr, qs, k), True)
run("H2D + pack, sync each step", lambda: db.pack_queries(qh.to(dev, non_blocking=True)), True)
run("query_host (one C call), sync inside", lambda: db.query_host(qh, k, out=out), False)
