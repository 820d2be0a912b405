"""CPU oracle for similarity + ranking + sharded merge + alpha query expansion.  TEST INFRASTRUCTURE ONLY.

Restates
  similarity + full ranking  /root/reference/src/benchmark/iris_evaluate.py:379-386
                             (F.normalize, torch.mm(q, g.t()), np.argsort(-sim, axis=1))
  cosine + top-k             /root/reference/reference/manus/7_AdaptiveHybridModel/modified/
                             adaptive_hybrid_retrieval_complete.py:11-16, 428 (torch.topk)
  query expansion skeleton   /root/reference/reference/manus/1_SPARSE/sparse_model.py:374-405
                             (search -> gather top-k rows -> renormalise -> search again)
The arithmetic lives in torch / numpy (pinned 2.7.1 / 2.3.1 by the reference's pyproject.toml:16-20; 2.11 / 2.3.5 here).

NOT pinned by the reference (it has no such code or tests): the tie rule (descending score, ties -> ascending index;
np.argsort's default order among equals is unspecified), top-k truncation, alpha-QE weights max(s,0)^alpha, and the
sharded merge.  Those follow SURVEY.md §7.3 / §8 a10 / §8e — "parity unpinned".

Comparison rule for quantised runs: oracle scores are fp32 torch.mm on the DE-QUANTISED tensors;
indices must match except where the oracle scores differ by <= eps*|s| (eps = 1e-3 bf16, 5e-3 fp8 vs an fp32 rescore).
"""
from __future__ import annotations

import numpy as np
import torch


def similarity(q: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    return torch.mm(q.float(), x.float().t())


def order_desc_stable(scores: np.ndarray) -> np.ndarray:
    """argsort of -scores with ties -> ascending index (stable sort on the negated scores)."""
    return np.argsort(-scores, axis=1, kind="stable")


def full_rank(q: torch.Tensor, x: torch.Tensor) -> np.ndarray:
    """ranks [nq, n] (the reference's np.argsort(-sim, axis=1) with a defined tie order)."""
    return order_desc_stable(similarity(q, x).numpy())


def topk(q: torch.Tensor, x: torch.Tensor, k: int, idx_offset: int = 0, chunk: int = 262144):
    """Exact top-k (scores [nq,k] fp32, idx [nq,k] int64), ties -> lower index; chunked so 1M-row sets fit in RAM."""
    nq, n = q.shape[0], x.shape[0]
    k = min(k, n)
    best_s = np.full((nq, 0), 0, dtype=np.float32)
    best_i = np.full((nq, 0), 0, dtype=np.int64)
    qf = q.float()
    for lo in range(0, n, chunk):
        s = torch.mm(qf, x[lo:lo + chunk].float().t()).numpy()
        kk = min(k, s.shape[1])
        # candidates of this chunk: everything >= the kk-th best value (keeps all ties), then exact ordering below
        part = np.partition(s, s.shape[1] - kk, axis=1)[:, s.shape[1] - kk][:, None]
        rows, cols = np.nonzero(s >= part)
        cs, ci = [], []
        for r in range(nq):
            sel = cols[rows == r]
            cs.append(s[r, sel])
            ci.append(sel.astype(np.int64) + lo)
        width = max(len(c) for c in cs) + best_s.shape[1]
        ms = np.full((nq, width), -np.inf, dtype=np.float32)
        mi = np.full((nq, width), np.iinfo(np.int64).max, dtype=np.int64)
        for r in range(nq):
            a = np.concatenate([best_s[r], cs[r]])
            b = np.concatenate([best_i[r], ci[r]])
            ms[r, : a.size] = a
            mi[r, : b.size] = b
        order = np.lexsort((mi, -ms.astype(np.float64)), axis=1)[:, :k]
        best_s = np.take_along_axis(ms, order, 1)
        best_i = np.take_along_axis(mi, order, 1)
    return best_s, best_i + idx_offset


def merge_shards(scores_list, idx_list, k: int):
    """k-way merge of per-shard (scores [nq,k], idx [nq,k] global) lists; -1 indices are padding."""
    s = np.concatenate(scores_list, axis=1)
    i = np.concatenate(idx_list, axis=1).astype(np.int64)
    s = np.where(i < 0, -np.inf, s)
    key_i = np.where(i < 0, np.iinfo(np.int64).max, i)
    order = np.lexsort((key_i, -s.astype(np.float64)), axis=1)[:, :k]
    return np.take_along_axis(s, order, 1), np.take_along_axis(i, order, 1)


def alpha_qe(q: torch.Tensor, x: torch.Tensor, scores: np.ndarray, idx: np.ndarray, kq: int = 10, alpha: float = 3.0):
    """q' = L2(q + sum_{j<kq} max(s_j,0)^alpha * x[idx_j])  (fp32)."""
    q = q.float()
    out = torch.empty_like(q)
    for r in range(q.shape[0]):
        acc = q[r].clone()
        for j in range(min(kq, idx.shape[1])):
            if idx[r, j] < 0:
                continue
            w = max(float(scores[r, j]), 0.0) ** alpha
            acc = acc + np.float32(w) * x[int(idx[r, j])].float()
        out[r] = acc / acc.norm().clamp_min(1e-12)
    return out


def indices_match_up_to_ties(got_idx, got_sc, ref_idx, ref_sc, rel_eps: float, abs_eps: float = 2e-6):
    """Bit-exact indices except inside near-tie groups: position j may differ only if the ORACLE scores of the two
    candidates (got_sc = oracle score of the returned row) differ by <= rel_eps * |score| (north_star's bar), with an
    absolute floor of abs_eps = 2e-6 for near-zero cosines in full rankings (fp32 accumulation-order noise of a
    2048-term dot product of unit vectors).  Returns (ok, message)."""
    got_idx = np.asarray(got_idx).astype(np.int64)
    ref_idx = np.asarray(ref_idx).astype(np.int64)
    if got_idx.shape != ref_idx.shape:
        return False, f"shape {got_idx.shape} vs {ref_idx.shape}"
    bad = np.argwhere(got_idx != ref_idx)
    for r, j in bad:
        a, b = float(got_sc[r, j]), float(ref_sc[r, j])
        tol = max(rel_eps * max(abs(a), abs(b)), abs_eps)
        if abs(a - b) > tol:
            return False, f"query {r} position {j}: got idx {got_idx[r, j]} (score {a}) vs oracle {ref_idx[r, j]} (score {b})"
    return True, f"{len(bad)} tie swaps"


def count_outranking(shard_scores: np.ndarray, idx_offset: int, ids, id_scores) -> np.ndarray:
    """Shard-local part of the count-based position (restates csrc/rank_positions.cu): given one query's
    scores of THIS shard's rows and the GLOBAL scores of its ground-truth ids, the number of shard rows that outrank
    each id in the order "score descending, ties -> lower global index".  Summed over the shards this is the id's
    position in the full ranking, i.e. what `np.arange(N)[np.in1d(ranks[:, i], ids)]` (utils/evaluate.py:76-80) reads
    off a complete np.argsort (iris_evaluate.py:386).  Row and id scores must come from the SAME arithmetic (a row
    must score exactly like its own threshold, or it would count itself)."""
    s = np.asarray(shard_scores)
    rows = np.arange(s.shape[0], dtype=np.int64) + int(idx_offset)
    out = np.zeros(len(ids), dtype=np.int64)
    for j, (p, sp) in enumerate(zip(ids, id_scores)):
        out[j] = int(np.sum((s > sp) | ((s == sp) & (rows < int(p)))))
    return out
