#!/usr/bin/env python
"""Does ENABLED peer access change the cost of a kernel boundary?  Times the single-GPU search step (scan -> select,
PDL-chained) on a 1/8-size shard before and after this process maps a peer GPU's memory (2+ GPUs needed)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import research_image_retrieval_b200 as rir  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
n, d, nq, k = 125916, 2048, 70, 100
gen = torch.Generator(device=dev).manual_seed(1)
X = torch.empty(n, d, device=dev, dtype=torch.bfloat16)
for lo in range(0, n, 65536):
    blk = torch.randn(min(65536, n - lo), d, generator=gen, device=dev)
    X[lo:lo + blk.shape[0]] = (blk / blk.norm(dim=1, keepdim=True)).to(torch.bfloat16)
Q = torch.randn(nq, d, generator=gen, device=dev)
Q = (Q / Q.norm(dim=1, keepdim=True)).to(torch.bfloat16)
db = rir.Database(X, None, "bf16")
out = (torch.empty((nq, k), dtype=torch.float32, device=dev), torch.empty((nq, k), dtype=torch.int32, device=dev))


def timed(label, steps=400):
    for _ in range(20):
        db.search(Q, None, k, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        db.search(Q, None, k, out=out)
    e1.record()
    torch.cuda.synchronize()
    print(f"{label:46s} {e0.elapsed_time(e1) / steps * 1e3:8.1f} us/step")


timed("no peer access")
if torch.cuda.device_count() > 1:
    a = torch.zeros(1 << 20, device="cuda:0")
    b = torch.zeros(1 << 20, device="cuda:1")
    b.copy_(a)                      # torch enables peer access both ways for the copy
    torch.cuda.synchronize()
    timed("peer access enabled (torch P2P copy done)")
    import ctypes
    lib = rir.load()
    p = ctypes.c_void_p()
    lib.rir_peer_alloc(1 << 20, ctypes.byref(p))
    h = ctypes.create_string_buffer(64)
    lib.rir_peer_export(p, h)
    timed("+ an IPC-exportable allocation on this GPU")
