"""CPU oracle for the revisited-protocol evaluation.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, in numpy + Python floats, the arithmetic of
  compute_ap            /root/reference/src/benchmark/utils/evaluate.py:4-34
  compute_map           /root/reference/src/benchmark/utils/evaluate.py:37-150
  compute_map_and_print /root/reference/src/benchmark/utils/evaluate.py:153-194
It is written independently of both the reference text and the CUDA kernel: positions are found with a set
membership mask, the junk shift with a cumulative sum, and the AP terms are added in rank order with Python floats
(IEEE double, same operation order as the reference: ((p0 + p1) * recall_step) / 2.0 added left to right).
Pinned against the reference's own functions in tests/test_oracle_vs_reference.py and tests/golden/map_*.npz.
"""
from __future__ import annotations

import numpy as np


def ap_from_adjusted_ranks(adj_ranks, nres: int) -> float:
    """AP by the trapezoid rule over the zero-based junk-adjusted ranks of the positives (evaluate.py:4-34)."""
    recall_step = 1.0 / nres
    ap = 0.0
    for i, r in enumerate(adj_ranks):
        r = int(r)
        p0 = 1.0 if r == 0 else float(i) / r
        p1 = float(i + 1) / (r + 1)
        ap += (p0 + p1) * recall_step / 2.0
    return ap


def _positions(ranked, ids):
    """positions j (ascending) of `ranked` whose id is in `ids` — np.in1d semantics (evaluate.py:76-80)."""
    ranked = np.asarray(ranked).reshape(-1)
    ids = np.asarray(ids).reshape(-1)
    if ids.size == 0 or ranked.size == 0:
        return np.zeros(0, dtype=np.int64)
    mask = np.isin(ranked, ids)
    return np.nonzero(mask)[0].astype(np.int64)


def query_ap(ranked, ok, junk, keeps=None):
    """One query: returns (ap, prs or None).  Raises ValueError like the reference when `keeps` is given and no
    positive was retrieved (evaluate.py:101: max() of an empty sequence)."""
    pos = _positions(ranked, ok)
    jnk = _positions(ranked, junk)
    # number of junk positions strictly before each positive (evaluate.py:82-91)
    shift = np.searchsorted(jnk, pos, side="left") if jnk.size else np.zeros_like(pos)
    adj = pos - shift
    ap = ap_from_adjusted_ranks(adj, len(np.asarray(ok).reshape(-1)))
    prs = None
    if keeps:
        one_based = adj + 1
        if one_based.size == 0:
            raise ValueError("max() iterable argument is empty")
        prs = np.zeros(len(keeps))
        for j, kappa in enumerate(keeps):
            kp = min(int(one_based.max()), kappa)
            prs[j] = (one_based <= kp).sum() / kp
    return ap, prs


def compute_map(ranks, gnd, keeps=None, li=False):
    """evaluate.py:37-150.  ranks [L, nq] (or list of lists with li=True); gnd list of {'ok', 'junk'}."""
    nq = len(gnd)
    aps = np.zeros(nq)
    prs = np.zeros((nq, len(keeps))) if keeps else None
    pr = np.zeros(len(keeps)) if keeps else None
    total = 0.0
    empty = 0
    for i in range(nq):
        ok = np.array(gnd[i]["ok"])
        if ok.shape[0] == 0:
            aps[i] = float("inf")
            if keeps:
                prs[i, :] = float("inf")
            empty += 1
            continue
        try:
            junk = np.array(gnd[i]["junk"])
        except Exception:
            junk = np.empty(0)
        ranked = np.asarray(ranks[i]) if li else np.asarray(ranks)[:, i]
        ap, p = query_ap(ranked, ok, junk, keeps)
        total += ap
        aps[i] = ap
        if keeps:
            prs[i, :] = p
            pr += prs[i, :]
    m = total / (nq - empty)
    if keeps:
        return m, aps, pr / (nq - empty), prs
    return m, aps


def revisited_gnd(gnd):
    """Easy / Medium / Hard {'ok','junk'} lists from {'easy','hard','junk'} (evaluate.py:163-185)."""
    e, m, h = [], [], []
    for g in gnd:
        e.append({"ok": np.concatenate([g["easy"]]), "junk": np.concatenate([g["junk"], g["hard"]])})
        m.append({"ok": np.concatenate([g["easy"], g["hard"]]), "junk": np.concatenate([g["junk"]])})
        h.append({"ok": np.concatenate([g["hard"]]), "junk": np.concatenate([g["junk"], g["easy"]])})
    return e, m, h


def compute_map_revisited(ranks, gnd, kappas=(1, 5, 10), li=False):
    """Returns ((mapE, apsE, mprE, prsE), (M...), (H...)) unrounded."""
    return tuple(compute_map(ranks, g, list(kappas), li=li) for g in revisited_gnd(gnd))


def compute_map_and_print_values(ranks, gnd, kappas=(1, 5, 10), li=False):
    """The 2-dp rounded (mapE, mapM, mapH) the reference returns (evaluate.py:194)."""
    e, m, h = compute_map_revisited(ranks, gnd, kappas, li)
    return tuple(np.around(x[0] * 100, decimals=2) for x in (e, m, h))
