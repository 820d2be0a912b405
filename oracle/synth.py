"""Seeded synthetic inputs of the BASELINE.json shapes (scaled down for CI).  TEST INFRASTRUCTURE ONLY.

Recipe: torch.Generator().manual_seed(seed); randn -> L2-normalise rows in fp32; per query `n_pos`
database rows are replaced by L2(q + sigma * randn) so there are realistic near-duplicates; ground truth per query =
random disjoint easy / hard / junk id lists, with a few queries given an empty `easy` list to exercise the
`inf` / excluded-query path of compute_map (utils/evaluate.py:65-68).
"""
from __future__ import annotations

import numpy as np
import torch


def unit_rows(n: int, d: int, gen: torch.Generator) -> torch.Tensor:
    x = torch.randn(n, d, generator=gen)
    return x / x.norm(dim=1, keepdim=True).clamp_min(1e-12)


def retrieval_set(nq: int, n: int, d: int, seed: int, n_pos: int = 8, sigma: float = 0.5):
    """Returns (Q [nq,d], X [n,d]) fp32 unit rows with planted positives, and the planted ids [nq, n_pos]."""
    gen = torch.Generator().manual_seed(seed)
    X = unit_rows(n, d, gen)
    Q = unit_rows(nq, d, gen)
    n_pos = min(n_pos, max(n // max(nq, 1), 0))
    planted = np.zeros((nq, n_pos), dtype=np.int64)
    if n_pos:
        perm = torch.randperm(n, generator=gen)[: nq * n_pos].reshape(nq, n_pos)
        for i in range(nq):
            noise = torch.randn(n_pos, d, generator=gen) * (sigma / d ** 0.5)
            rows = Q[i][None, :] + noise
            X[perm[i]] = rows / rows.norm(dim=1, keepdim=True)
        planted = perm.numpy()
    return Q, X, planted


def revisited_gnd(nq: int, n: int, seed: int, n_empty_easy: int = 2):
    """Random disjoint easy (U[5,60]) / hard (U[5,80]) / junk (U[0,40]) id lists per query (scaled to n)."""
    rng = np.random.RandomState(seed)
    gnd = []
    for i in range(nq):
        ne, nh, nj = rng.randint(5, 61), rng.randint(5, 81), rng.randint(0, 41)
        tot = min(ne + nh + nj, n)
        ids = rng.choice(n, size=tot, replace=False)
        ne = min(ne, tot)
        nh = min(nh, tot - ne)
        g = {"easy": np.sort(ids[:ne]), "hard": np.sort(ids[ne:ne + nh]), "junk": np.sort(ids[ne + nh:])}
        if i < n_empty_easy:
            g["junk"] = np.sort(np.concatenate([g["junk"], g["easy"]]))
            g["easy"] = np.zeros(0, dtype=ids.dtype)
        gnd.append(g)
    return gnd


def feature_maps(B: int, C: int, H: int, W: int, seed: int) -> torch.Tensor:
    """relu(randn) * 2 — non-negative like post-ReLU conv5 activations."""
    gen = torch.Generator().manual_seed(seed)
    return torch.relu(torch.randn(B, C, H, W, generator=gen)) * 2.0
