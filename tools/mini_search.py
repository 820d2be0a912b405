#!/usr/bin/env python
"""Smallest fused-scan search (one fused tcgen05 scan + select), for ncu / launch-attribute probing on the GPU box."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import research_image_retrieval_b200 as rir  # noqa: E402

dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1)
n, d, nq, k = 90000, 256, 6, 50
X = torch.randn(n, d, generator=gen, device=dev)
X = (X / X.norm(dim=1, keepdim=True)).to(torch.bfloat16)
Q = torch.randn(nq, d, generator=gen, device=dev)
Q = (Q / Q.norm(dim=1, keepdim=True)).to(torch.bfloat16)
db = rir.Database(X, None, "bf16")
for _ in range(3):
    sc, ix = db.search(Q, None, k)
sc2, ix2 = db.search(Q, None, k, path="exact")
torch.cuda.synchronize()
assert torch.equal(ix, ix2), "fused scan != exact path"
print("mini_search OK", float(sc[0, 0]))
