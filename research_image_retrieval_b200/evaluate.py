"""Revisited-protocol evaluation on the GPU — drop-in for the reference's `utils/evaluate.py`.

Same call signatures, return arity, printed text, `inf` conventions, exceptions and 2-dp rounding as

  compute_ap            utils/evaluate.py:4-34
  compute_map           utils/evaluate.py:37-150
  compute_map_and_print utils/evaluate.py:153-194          (duplicate copies at iris_evaluate.py:11-265)

The arithmetic runs in `rir_compute_map` (csrc/evaluate_map.cu) in fp64 with the reference's operation order, so the
returned numbers equal the reference's Python floats.  Host code here only converts the ground-truth dictionaries to
sorted CSR id lists and formats the results.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib

PROTO_OK_A_JUNK_C = 0x41       # compute_map: ok = list A, junk = list C
PROTO_EASY = 0x61              # ok = easy(A)        junk = junk(C) | hard(B)     utils/evaluate.py:163-169
PROTO_MEDIUM = 0x43            # ok = easy | hard    junk = junk                  utils/evaluate.py:171-177
PROTO_HARD = 0x52              # ok = hard(B)        junk = junk(C) | easy(A)     utils/evaluate.py:179-185


# ----------------------------------------------------------------------------------------------
# host-side format conversion (no arithmetic)
# ----------------------------------------------------------------------------------------------
def ids_to_csr(lists):
    """[array-like of ids per query] -> (ids int32 sorted within each query, off int32[nq+1]).

    Duplicates are kept: the reference uses len(ok) (duplicates included) as the number of positives.
    One concatenate + one lexsort for the whole query set (no per-query numpy calls)."""
    nq = len(lists)
    arrs = [np.asarray(l).reshape(-1) for l in lists]
    lens = np.fromiter((a.size for a in arrs), dtype=np.int64, count=nq)
    off = np.zeros(nq + 1, dtype=np.int32)
    np.cumsum(lens, out=off[1:])
    if int(off[-1]) == 0:
        return np.empty(0, dtype=np.int32), off
    flat = np.concatenate([a.astype(np.int64, copy=False) for a in arrs if a.size])
    owner = np.repeat(np.arange(nq, dtype=np.int64), lens)
    order = np.lexsort((flat, owner))            # by query, then by id
    return np.ascontiguousarray(flat[order], dtype=np.int32), off


def ranks_to_rows(ranks, li: bool, nq: int):
    """Reference layouts -> row-per-query int32 [nq, L] on the host or device; -1 pads ragged lists.

    li=False: `ranks` is [L, nq] (column per query, utils/evaluate.py:49,79); li=True: list of nq ranked lists."""
    if li:
        lens = [len(r) for r in ranks]
        L = max(lens) if lens else 0
        rows = np.full((nq, max(L, 1)), -1, dtype=np.int32)
        for i in range(nq):
            if lens[i]:
                rows[i, : lens[i]] = np.asarray(ranks[i], dtype=np.int64)
        return rows, L
    if isinstance(ranks, torch.Tensor):
        if ranks.dim() != 2:
            raise ValueError("ranks must be 2-D [L, nq]")
        return ranks.t().to(torch.int32).contiguous(), ranks.shape[0]
    r = np.asarray(ranks)
    if r.ndim != 2:
        raise ValueError("ranks must be 2-D [L, nq]")
    return np.ascontiguousarray(r.T, dtype=np.int32), r.shape[0]  # one transposing + narrowing pass


def _dev(a, device):
    if isinstance(a, torch.Tensor):
        return a.to(device)
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def _run_map(rows, L, nq, lists_abc, protos, keeps, device=None, positions=None):
    """Launch rir_compute_map; returns host numpy (map[P], aps[P,nq], mpr[P,nk], prs[P,nq,nk], status[P,nq]).

    Two host->device copies (ranked lists; all id lists + offsets packed into one int32 array) and ONE device->host
    copy (every output carved out of one buffer): at ROxford size the copies and their synchronisations, not the
    kernel, are the cost."""
    lib = _lib.load()
    if device is None:
        device = rows.device if isinstance(rows, torch.Tensor) and rows.is_cuda else torch.device("cuda", torch.cuda.current_device())
    rows_d = _dev(rows, device)
    ld = rows_d.shape[1]
    # pack [ids_A | off_A | ids_B | off_B | ids_C | off_C] (absent lists contribute nothing)
    pieces, where, pos = [], [], 0
    for lst in lists_abc:
        if lst is None:
            where.append(None)
            continue
        ids, off = lst
        ids = ids if ids.size else np.zeros(1, dtype=np.int32)
        where.append((pos, pos + ids.size))
        pieces += [np.ascontiguousarray(ids, dtype=np.int32), np.ascontiguousarray(off, dtype=np.int32)]
        pos += ids.size + off.size
    meta_d = _dev(np.concatenate(pieces) if pieces else np.zeros(1, dtype=np.int32), device)
    ptrs = []
    for w in where:
        ptrs += [None, None] if w is None else [meta_d.data_ptr() + 4 * w[0], meta_d.data_ptr() + 4 * w[1]]
    P = len(protos)
    keeps = list(keeps) if keeps else []
    nk = len(keeps)
    nk1 = max(nk, 1)
    n_f64 = P + P * nq + P * nk1 + P * nq * nk1
    out = torch.empty(n_f64 * 8 + P * nq * 4, dtype=torch.uint8, device=device)   # fp64 block, then int32 status
    base = out.data_ptr()
    o_map, o_aps, o_mpr, o_prs = 0, P, P + P * nq, P + P * nq + P * nk1
    proto_arr = (ctypes.c_int32 * P)(*protos)
    kappa_arr = (ctypes.c_int32 * nk1)(*([int(k) for k in keeps] or [0]))
    with torch.cuda.device(device):
        if positions is None:
            _lib.check(lib.rir_compute_map(rows_d.data_ptr(), nq, int(L), int(ld), ptrs[0], ptrs[1], ptrs[2], ptrs[3],
                                           ptrs[4], ptrs[5], proto_arr, P, kappa_arr, nk, base + 8 * o_map,
                                           base + 8 * o_aps, base + 8 * o_mpr, base + 8 * o_prs, base + 8 * n_f64,
                                           _lib.stream_ptr()))
        else:  # compact lists of ground-truth ids + their positions in the full ranking (rir_rank_count)
            pos_d = _dev(positions, device)
            assert tuple(pos_d.shape) == tuple(rows_d.shape) and pos_d.dtype == torch.int32 and pos_d.is_contiguous()
            _lib.check(lib.rir_compute_map_at(rows_d.data_ptr(), pos_d.data_ptr(), nq, int(L), int(ld), ptrs[0], ptrs[1],
                                              ptrs[2], ptrs[3], ptrs[4], ptrs[5], proto_arr, P, kappa_arr, nk,
                                              base + 8 * o_map, base + 8 * o_aps, base + 8 * o_mpr, base + 8 * o_prs,
                                              base + 8 * n_f64, _lib.stream_ptr()))
    host = out.cpu().numpy()
    f64 = host[: n_f64 * 8].view(np.float64)
    status = host[n_f64 * 8:].view(np.int32).reshape(P, nq)
    res = (f64[o_map:o_aps].copy(), f64[o_aps:o_mpr].reshape(P, nq).copy(),
           f64[o_mpr:o_prs].reshape(P, nk1)[:, :nk].copy(), f64[o_prs:].reshape(P, nq, nk1)[:, :, :nk].copy(), status.copy())
    del meta_d, rows_d
    return res


def _finish(map_p, aps_p, mpr_p, prs_p, status_p, keeps):
    """Apply the reference's Python-level conventions to one protocol's raw outputs."""
    nq = aps_p.shape[0]
    if keeps and np.any(status_p == _lib.RIR_MAP_NO_POS_RETRIEVED):
        # utils/evaluate.py:101 — max(pos) over an empty array
        raise ValueError("max() iterable argument is empty")
    if np.all(status_p == _lib.RIR_MAP_EMPTY_OK):
        # utils/evaluate.py:105 — Python float divided by the int 0
        raise ZeroDivisionError("float division by zero")
    if keeps:
        return np.float64(map_p), aps_p.copy(), mpr_p.copy(), prs_p.reshape(nq, len(keeps)).copy()
    return np.float64(map_p), aps_p.copy()


# ----------------------------------------------------------------------------------------------
# reference API
# ----------------------------------------------------------------------------------------------
def compute_ap(ranks, nres):
    """Average precision from the zero-based, junk-adjusted ranks of the positives (utils/evaluate.py:4-34).

    Evaluated by the same GPU kernel as compute_map: the ranks become a ranked list with the positive id 1 at the
    given positions and `nres` copies of id 1 as the ok list (len(ok) == nres)."""
    r = np.asarray(ranks, dtype=np.int64).reshape(-1)
    if r.size == 0:
        return 0.0
    if np.any(np.diff(r) <= 0) or r[0] < 0:
        raise ValueError("compute_ap expects strictly increasing zero-based ranks")
    L = int(r[-1]) + 1
    rows = np.zeros((1, L), dtype=np.int32)
    rows[0, r] = 1
    ok = (np.ones(int(nres), dtype=np.int32), np.array([0, int(nres)], dtype=np.int32))
    m, aps, _, _, _ = _run_map(rows, L, 1, [ok, None, None], [PROTO_OK_A_JUNK_C], None)
    return float(aps[0, 0])


def compute_map(ranks, gnd, keeps=None, li=False):
    """mAP (and mP@k) of ranked lists against {'ok', 'junk'} ground truth — utils/evaluate.py:37-150.

    ranks: int array [L, nq] (column per query; L may be a truncated top-k) or, with li=True, a list of nq ranked lists
    of possibly different lengths.  Returns (mAP, aps) or, with `keeps`, (mAP, aps, pr, prs)."""
    nq = len(gnd)
    ok = ids_to_csr([g["ok"] for g in gnd])
    junk_lists = []
    for g in gnd:
        try:
            junk_lists.append(g["junk"])
        except Exception:  # the reference tolerates a missing 'junk' key (utils/evaluate.py:70-73)
            junk_lists.append([])
    junk = ids_to_csr(junk_lists)
    rows, L = ranks_to_rows(ranks, li, nq)
    m, aps, mpr, prs, status = _run_map(rows, L, nq, [ok, None, junk], [PROTO_OK_A_JUNK_C], keeps)
    return _finish(m[0], aps[0], mpr[0], prs[0], status[0], keeps)


def revisited_map(ranks, gnd, kappas=(1, 5, 10), li=False):
    """Easy / Medium / Hard in ONE launch.  Returns [(mAP, aps, mpr, prs)] * 3 (E, M, H)."""
    nq = len(gnd)
    easy = ids_to_csr([g["easy"] for g in gnd])
    hard = ids_to_csr([g["hard"] for g in gnd])
    junk = ids_to_csr([g["junk"] for g in gnd])
    rows, L = ranks_to_rows(ranks, li, nq)
    kappas = list(kappas)
    m, aps, mpr, prs, status = _run_map(rows, L, nq, [easy, hard, junk], [PROTO_EASY, PROTO_MEDIUM, PROTO_HARD], kappas)
    return [_finish(m[p], aps[p], mpr[p], prs[p], status[p], kappas) for p in range(3)]


# ----------------------------------------------------------------------------------------------
# full-protocol evaluation without a full ranking (R1M: +1M distractors)
# ----------------------------------------------------------------------------------------------
def gnd_positions(db, q_rows, q_scale, id_lists):
    """Positions of the given database ids in each query's FULL ranking (score descending, ties -> lower index) over a
    Database or ShardedDatabase — what `np.arange(N)[np.in1d(ranks[:, i], ids)]` reads off the reference's complete
    np.argsort (utils/evaluate.py:76-80, iris_evaluate.py:386), computed as counts: position(p) = number of rows
    that outrank p.  Each shard counts over its own rows; two small all-reduces (scores, counts) join the shards.

    id_lists: per query an array of GLOBAL row ids.  Returns (ranked_ids [nq, m_pad] int32 CUDA — the query's ids in
    rank order, -1 padded —, positions [nq, m_pad] int32 CUDA)."""
    import torch.distributed as dist

    from .search import _DTYPES, ShardedDatabase
    local = db.local if isinstance(db, ShardedDatabase) else db
    sharded = isinstance(db, ShardedDatabase) and db.world > 1
    nq = q_rows.shape[0]
    if len(id_lists) != nq:
        raise ValueError("one id list per query")
    uniq = [np.unique(np.asarray(l).reshape(-1).astype(np.int64)) for l in id_lists]
    ids, off = ids_to_csr(uniq)
    longest = max((u.size for u in uniq), default=0)
    m_pad = 32
    while m_pad < longest:
        m_pad *= 2
    if m_pad > 4096:
        raise ValueError(f"at most 4096 ground-truth ids per query (got {longest})")
    dev = local.rows.device
    lib = _lib.load()
    dt = _DTYPES[local.dtype]
    ids_d = _dev(ids if ids.size else np.zeros(1, dtype=np.int32), dev)
    off_d = _dev(off, dev)
    scores = torch.zeros(max(int(off[-1]), 1), dtype=torch.float32, device=dev)
    qs_ptr = None if q_scale is None else q_scale.data_ptr()
    xs_ptr = None if local.scale is None else local.scale.data_ptr()
    q_rows = q_rows.contiguous()
    with torch.cuda.device(dev):
        _lib.check(lib.rir_gnd_scores(q_rows.data_ptr(), local.rows.data_ptr(), dt, qs_ptr, xs_ptr, nq, local.n, local.d,
                                      local.idx_offset, ids_d.data_ptr(), off_d.data_ptr(), int(off[-1]),
                                      scores.data_ptr(), _lib.stream_ptr()))
    if sharded:
        dist.all_reduce(scores, group=db.group)       # every id is scored by exactly one shard
    keys = torch.empty((nq, m_pad), dtype=torch.int64, device=dev)   # uint64 ranking keys
    counts = torch.empty((nq, m_pad), dtype=torch.int32, device=dev)
    ws = torch.empty(nq * m_pad * 4, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.rir_rank_thresholds(scores.data_ptr(), ids_d.data_ptr(), off_d.data_ptr(), nq, m_pad,
                                           keys.data_ptr(), _lib.stream_ptr()))
        _lib.check(lib.rir_rank_count(q_rows.data_ptr(), local.rows.data_ptr(), dt, qs_ptr, xs_ptr, nq, local.n, local.d,
                                      local.idx_offset, keys.data_ptr(), m_pad, counts.data_ptr(), ws.data_ptr(),
                                      ws.numel(), _lib.stream_ptr()))
    if sharded:
        dist.all_reduce(counts, group=db.group)
    ranked = (0xFFFFFFFF - (keys & 0xFFFFFFFF)).to(torch.int32)
    ranked = torch.where(keys == 0, torch.full_like(ranked, -1), ranked)
    return ranked.contiguous(), counts


def revisited_map_full(db, q_rows, q_scale, gnd, kappas=(1, 5, 10)):
    """Easy / Medium / Hard over the FULL ranking of a (sharded) database that is never sorted or materialised:
    [(mAP, aps, mpr, prs)] * 3, equal to revisited_map(full_ranks, gnd, kappas) on the complete ranked lists."""
    nq = len(gnd)
    lists = [np.concatenate([np.asarray(g["easy"]).reshape(-1), np.asarray(g["hard"]).reshape(-1),
                             np.asarray(g["junk"]).reshape(-1)]) for g in gnd]
    ranked, pos = gnd_positions(db, q_rows, q_scale, lists)
    easy = ids_to_csr([g["easy"] for g in gnd])
    hard = ids_to_csr([g["hard"] for g in gnd])
    junk = ids_to_csr([g["junk"] for g in gnd])
    kappas = list(kappas)
    m, aps, mpr, prs, status = _run_map(ranked, ranked.shape[1], nq, [easy, hard, junk],
                                        [PROTO_EASY, PROTO_MEDIUM, PROTO_HARD], kappas, positions=pos)
    return [_finish(m[p], aps[p], mpr[p], prs[p], status[p], kappas) for p in range(3)]


def compute_map_full(db, q_rows, q_scale, gnd, keeps=None):
    """compute_map (utils/evaluate.py:37-150) over the full ranking of a (sharded) database, via gnd_positions."""
    nq = len(gnd)
    junk_lists = [g["junk"] if "junk" in g else [] for g in gnd]
    lists = [np.concatenate([np.asarray(g["ok"]).reshape(-1), np.asarray(j).reshape(-1)]) for g, j in zip(gnd, junk_lists)]
    ranked, pos = gnd_positions(db, q_rows, q_scale, lists)
    ok = ids_to_csr([g["ok"] for g in gnd])
    junk = ids_to_csr(junk_lists)
    m, aps, mpr, prs, status = _run_map(ranked, ranked.shape[1], nq, [ok, None, junk], [PROTO_OK_A_JUNK_C], keeps,
                                        positions=pos)
    return _finish(m[0], aps[0], mpr[0], prs[0], status[0], keeps)


def compute_map_and_print(dataset, featuretype, mode, ranks, gnd, kappas=[1, 5, 10], verbose=False, li=False):
    """utils/evaluate.py:153-194 — same prints, same 2-dp rounded (mapE, mapM, mapH) return."""
    # old evaluation protocol: the reference unpacks 4 values from a 2-tuple and raises (utils/evaluate.py:157)
    if dataset.startswith('oxford5k') or dataset.startswith('paris6k'):
        map, aps, _, _ = compute_map(ranks, gnd)
        print('>> {}: mAP {:.2f}'.format(dataset, np.around(map * 100, decimals=2)))

    # new evaluation protocol
    elif dataset.startswith('roxford5k') or dataset.startswith('rparis6k'):
        (mapE, apsE, mprE, prsE), (mapM, apsM, mprM, prsM), (mapH, apsH, mprH, prsH) = \
            revisited_map(ranks, gnd, kappas, li=li)

        print('>> Test Dataset: {} *** Feature Type: {} >>'.format(dataset, featuretype))
        print('>> mAP Eeay: {}, Medium: {}, Hard: {}'.format(np.around(mapE * 100, decimals=2), np.around(mapM * 100, decimals=2), np.around(mapH * 100, decimals=2)))
        print('>> mP@k{} Easy: {}, Medium: {}, Hard: {}'.format(kappas, np.around(mprE * 100, decimals=2), np.around(mprM * 100, decimals=2), np.around(mprH * 100, decimals=2)))

        if verbose:
            print('>> Query aps: >>\nEeay: {}\nMedium: {}\nHard: {}'.format(np.around(apsE * 100, decimals=2), np.around(apsM * 100, decimals=2), np.around(apsH * 100, decimals=2)))

        return np.around(mapE * 100, decimals=2), np.around(mapM * 100, decimals=2), np.around(mapH * 100, decimals=2)
