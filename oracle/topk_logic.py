"""CPU emulation of the ranking-key and fused-threshold LOGIC of the tcgen05 scan.  TEST INFRASTRUCTURE ONLY.

No reference counterpart (the reference sorts the whole score matrix, iris_evaluate.py:386): this restates, in numpy,
the rules csrc/rir_common.cuh (keys), csrc/sim_topk_mma.cu (first-phase keys, threshold) and csrc/sim_topk_select.cu
(merge of the first-phase keys, redo rule) implement, so the exactness argument can be property-tested without a GPU:

  * key = ordered(score) << 32 | ~row : larger key <=> (higher score, then lower row);
  * tau = lower edge of the 24-bit bucket of the k-th best KEPT key  =>  at least k kept rows score >= tau;
  * rows of filter tiles are kept iff score >= tau; first-phase tiles contribute their kept keys >= tau, and a tile
    whose LAST kept key still reaches tau is re-scored completely;
  * the top-k of that candidate set equals the top-k of all rows.
"""
from __future__ import annotations

import numpy as np

TILE = 256
TOPT = 8


def float_to_ordered(s: np.ndarray) -> np.ndarray:
    s = np.asarray(s, dtype=np.float32) + np.float32(0.0)          # -0 -> +0
    u = s.view(np.uint32)
    return np.where(u & np.uint32(0x80000000), ~u, u | np.uint32(0x80000000)).astype(np.uint32)


def ordered_to_float(o: np.ndarray) -> np.ndarray:
    o = np.asarray(o, dtype=np.uint32)
    u = np.where(o & np.uint32(0x80000000), o & np.uint32(0x7FFFFFFF), ~o).astype(np.uint32)
    return u.view(np.float32)


def make_key(score: np.ndarray, row: np.ndarray) -> np.ndarray:
    return (float_to_ordered(score).astype(np.uint64) << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - np.asarray(row, np.uint64))


def key_score(key: np.ndarray) -> np.ndarray:
    return ordered_to_float((np.asarray(key, np.uint64) >> np.uint64(32)).astype(np.uint32))


def key_row(key: np.ndarray) -> np.ndarray:
    return (np.uint64(0xFFFFFFFF) - (np.asarray(key, np.uint64) & np.uint64(0xFFFFFFFF))).astype(np.int64)


def fused_tau(kept: np.ndarray, k: int) -> np.float32:
    """cta_fused_tau: 3 radix passes over the top 24 bits of the non-zero kept keys."""
    kept = np.asarray(kept, dtype=np.uint64)
    nz = kept[kept != 0]
    if nz.size < k:
        return np.float32(-np.inf)
    prefix, need = 0, k
    for p in range(3):
        shift = 56 - 8 * p
        sel = nz if p == 0 else nz[(nz >> np.uint64(shift + 8)) == np.uint64(prefix)]
        digits = ((sel >> np.uint64(shift)) & np.uint64(255)).astype(np.int64)
        hist = np.bincount(digits, minlength=256)
        cum = 0
        for dgt in range(255, -1, -1):
            if cum < need <= cum + hist[dgt]:
                prefix, need = (prefix << 8) | dgt, need - cum
                break
            cum += hist[dgt]
    return ordered_to_float(np.array([(prefix << 8) & 0xFFFFFFFF], dtype=np.uint32))[0]


def fused_scan_topk(scores: np.ndarray, k: int, first_tiles, topt: int = TOPT):
    """One query.  scores[n] fp32; first_tiles = physical tile indices of the first phase.  Returns (rows of the exact
    top-k by key order, number of candidates, number of re-scored tiles)."""
    n = scores.shape[0]
    keys = make_key(scores, np.arange(n))
    ntiles = -(-n // TILE)
    first = set(int(t) for t in first_tiles)
    kept_per_tile = {}
    for t in first:
        seg = keys[t * TILE:(t + 1) * TILE]
        kept = np.sort(seg)[::-1][:topt]
        kept_per_tile[t] = np.concatenate([kept, np.zeros(topt - kept.size, np.uint64)])
    tau = fused_tau(np.concatenate(list(kept_per_tile.values())), k)
    cand, redo = [], 0
    for t in range(ntiles):
        seg_k, seg_s = keys[t * TILE:(t + 1) * TILE], scores[t * TILE:(t + 1) * TILE]
        if t in first:
            kept = kept_per_tile[t]
            last = kept[topt - 1]
            if last != 0 and key_score(np.array([last]))[0] >= tau:      # may hide more rows >= tau: re-score the tile
                redo += 1
                cand.append(seg_k[seg_s >= tau])
            else:
                kk = kept[kept != 0]
                cand.append(kk[key_score(kk) >= tau])
        else:
            cand.append(seg_k[seg_s >= tau])
    cand = np.concatenate(cand)
    top = np.sort(cand)[::-1][:k]
    return key_row(top), cand.size, redo
