"""GPU parity: similarity + top-k (+ merge, alpha-QE) through the C ABI vs the fp32 oracle on de-quantised inputs.

Bar (BASELINE.json north_star): indices bit-exact except ties within 1e-3 relative (bf16) / 5e-3 relative (fp8,
against an fp32 rescore).  fp32 descriptors: indices exact except fp32-rounding ties (1e-5)."""
import numpy as np
import pytest
import torch

import research_image_retrieval_b200 as rir
from conftest import load_golden
from oracle import descriptor_oracle as D
from oracle import search_oracle as S
from oracle import synth

pytestmark = pytest.mark.gpu
EPS = {"fp32": 1e-5, "bf16": 1e-3, "fp8": 5e-3}


def _dequant(rows, scale, dtype):
    r = rows.cpu()
    if dtype == "fp8":
        r = r.view(torch.float8_e4m3fn).float() * scale.cpu()[:, None]
    return r.float()


def _check(got_sc, got_ix, Qf, Xf, k, eps, idx_offset=0):
    """got vs oracle on the same (de-quantised) fp32 inputs."""
    ref_sc, ref_ix = S.topk(Qf, Xf, k, idx_offset=idx_offset)
    got_ix = got_ix.cpu().numpy().astype(np.int64)
    got_sc = got_sc.cpu().numpy()
    assert got_ix.shape == ref_ix.shape
    # the scores we return must be the fp32 dot products of the rows we return
    sim_of_got = np.stack([(Xf[torch.from_numpy(got_ix[r] - idx_offset)] @ Qf[r]).numpy() for r in range(Qf.shape[0])])
    np.testing.assert_allclose(got_sc, sim_of_got, rtol=2e-4, atol=2e-6)
    ok, msg = S.indices_match_up_to_ties(got_ix, sim_of_got, ref_ix, ref_sc, eps)
    assert ok, msg
    # rows come out in non-increasing score order, and each index appears once
    assert np.all(np.diff(got_sc, axis=1) <= 0)
    for r in range(got_ix.shape[0]):
        assert len(set(got_ix[r].tolist())) == k
    return msg


CASES = [
    # nq, n, d, k
    (1, 5000, 2048, 100),      # scan-all (n <= 16384), 1 query
    (3, 20011, 512, 100),      # filtered scan, n not a multiple of 256
    (4, 70000, 256, 10),
    (8, 33000, 128, 17),
    (70, 4993, 2048, 4993),    # cfg-1 shape: full ranking
    (70, 50000, 1024, 100),    # cfg-2 shape scaled down
    (130, 40000, 256, 50),     # two query blocks on the MMA path
    (5, 18000, 72, 5),         # d not a multiple of the 128-byte K chunk
    (2, 300, 64, 300),         # tiny database, full ranking
]


@pytest.mark.parametrize("dtype", ["bf16", "fp8", "fp32"])
@pytest.mark.parametrize("path", ["stream", "mma", "exact"])
@pytest.mark.parametrize("nq,n,d,k", CASES)
def test_sim_topk_paths(cuda_device, dtype, path, nq, n, d, k):
    if dtype == "fp32" and path == "mma":
        pytest.skip("fp32 descriptors run on the CUDA-core paths only")
    if path == "exact" and nq * n * d > 4e9:
        pytest.skip("exact path is the slow fallback; covered at smaller sizes")
    Q, X, planted = synth.retrieval_set(nq, n, d, seed=n + d + nq)
    db = rir.Database.from_descriptors(X.to(cuda_device), dtype)
    qr, qs = db.pack_queries(Q.to(cuda_device))
    sc, ix = db.search(qr, qs, k, path=path)
    torch.cuda.synchronize()
    Xf = _dequant(db.rows, db.scale, dtype)[:, :d]
    Qf = _dequant(qr, qs, dtype)[:, :d]
    _check(sc, ix, Qf, Xf, k, EPS[dtype] if dtype != "fp8" else 1e-3)  # vs the de-quantised oracle: accumulate-order only
    if planted.shape[1] and k >= planted.shape[1] and dtype != "fp8":
        got = ix.cpu().numpy()
        for r in range(nq):
            assert set(planted[r].tolist()) <= set(got[r, : max(k, planted.shape[1])].tolist())


FUSED_CASES = [
    # nq, n, d, k, dtype   (n >= 2 * 148 * 256 rows: the tcgen05 path runs as ONE fused launch)
    (70, 100003, 256, 100, "bf16"),   # one query block, query box trimmed to 80 rows, pair shares the query chunk
    (5, 90000, 128, 10, "bf16"),      # tiny batch on the tensor path (box trimmed to 16 rows)
    (130, 80000, 64, 50, "fp8"),      # two query blocks in one CTA (512 TMEM columns)
    (300, 90000, 128, 100, "bf16"),   # two super-blocks: pair shares the database chunk, one dummy query block
    (600, 76000, 64, 20, "bf16"),     # three super-blocks (odd): dummy super-block in the last pair
]


@pytest.mark.parametrize("nq,n,d,k,dtype", FUSED_CASES)
def test_fused_scan(cuda_device, nq, n, d, k, dtype):
    """Fused tcgen05 scan (first round = sample, in-kernel threshold, grid barrier) vs the oracle."""
    Q, X, planted = synth.retrieval_set(nq, n, d, seed=n + nq)
    db = rir.Database.from_descriptors(X.to(cuda_device), dtype)
    qr, qs = db.pack_queries(Q.to(cuda_device))
    sc, ix = db.search(qr, qs, k, path="mma")
    torch.cuda.synchronize()
    Xf = _dequant(db.rows, db.scale, dtype)[:, :d]
    Qf = _dequant(qr, qs, dtype)[:, :d]
    _check(sc, ix, Qf, Xf, k, 1e-3)
    # the exact path is an independent implementation: same indices up to ties
    sc2, ix2 = db.search(qr[:4].contiguous(), None if qs is None else qs[:4].contiguous(), k, path="exact")
    ok, msg = S.indices_match_up_to_ties(ix[:4].cpu().numpy().astype(np.int64), sc[:4].cpu().numpy(),
                                         ix2.cpu().numpy().astype(np.int64), sc2.cpu().numpy(), 1e-3)
    assert ok, msg


def test_fused_scan_first_phase_redo(cuda_device):
    """Clustered database: tile 0 (always a first-phase tile) holds 40 near-duplicates of query 0 — more than the 8 keys
    the first phase keeps per tile — so the select kernel must re-score that tile.  Another 30 sit in the last tile."""
    nq, n, d, k = 6, 120000, 128, 64
    Q, X, _ = synth.retrieval_set(nq, n, d, seed=4242, n_pos=0)
    gen = torch.Generator().manual_seed(6)
    near = Q[0][None, :] + torch.randn(70, d, generator=gen) * (0.3 / d ** 0.5)
    near = near / near.norm(dim=1, keepdim=True)
    rows = np.concatenate([np.arange(3, 43), np.arange(n - 30, n)])
    X[torch.from_numpy(rows)] = near
    db = rir.Database.from_descriptors(X.to(cuda_device), "bf16")
    qr, qs = db.pack_queries(Q.to(cuda_device))
    sc, ix = db.search(qr, qs, k, path="mma")
    Xf, Qf = db.rows.float().cpu(), qr.float().cpu()
    _check(sc, ix, Qf, Xf, k, 1e-3)
    assert set(ix[0].cpu().tolist()) <= set(rows.tolist())


def test_fp8_against_fp32_rescore(cuda_device):
    """fp8 bar (north_star): indices exact except ties within 5e-3 relative, judged on TRUE fp32 scores.  The fp8 scan
    over-fetches 2k+16 candidates and a bf16 re-score of those rows decides the final order."""
    nq, n, d, k = 16, 30000, 512, 8
    Q, X, planted = synth.retrieval_set(nq, n, d, seed=123, n_pos=8)
    db = rir.Database.from_descriptors(X.to(cuda_device), "fp8", rescore=True)
    sc, ix = db.query(Q.to(cuda_device), k)
    ref_sc, ref_ix = S.topk(Q, X, k)
    got = ix.cpu().numpy().astype(np.int64)
    true_of_got = np.stack([(X[torch.from_numpy(got[r])] @ Q[r]).numpy() for r in range(nq)])
    ok, msg = S.indices_match_up_to_ties(got, true_of_got, ref_ix, ref_sc, 5e-3)
    assert ok, msg
    np.testing.assert_allclose(sc.cpu().numpy(), true_of_got, rtol=5e-3)
    for r in range(nq):
        assert set(got[r].tolist()) == set(planted[r].tolist())
    # random (non-planted) neighbours, k=100: same bar
    sc, ix = db.query(Q.to(cuda_device), 100)
    ref_sc, ref_ix = S.topk(Q, X, 100)
    got = ix.cpu().numpy().astype(np.int64)
    true_of_got = np.stack([(X[torch.from_numpy(got[r])] @ Q[r]).numpy() for r in range(nq)])
    ok, msg = S.indices_match_up_to_ties(got, true_of_got, ref_ix, ref_sc, 5e-3)
    assert ok, msg


def test_auto_path_and_grouping(cuda_device):
    Q, X, _ = synth.retrieval_set(9, 21000, 64, seed=3)
    db = rir.Database.from_descriptors(X.to(cuda_device), "bf16")
    qr, qs = db.pack_queries(Q.to(cuda_device))
    a = db.search(qr, qs, 20, path="auto")
    b = db.search(qr, qs, 20, path="stream")  # 9 queries -> two stream launches of 8 + 1
    c = db.search(qr[:2].contiguous(), None, 20, path="auto")  # <= 2 queries on a small shard -> stream
    assert torch.equal(a[1], b[1]) and torch.equal(a[1][:2], c[1])
    np.testing.assert_allclose(a[0].cpu().numpy(), b[0].cpu().numpy(), rtol=1e-5, atol=1e-6)


def test_exact_ties_ascending_index(cuda_device):
    """Duplicate rows have bit-identical scores on every path: they must come out in ascending index order."""
    Q, X, _ = synth.retrieval_set(2, 17000, 128, seed=8)
    dup = [16999, 5, 9000, 777]
    for i in dup:
        X[i] = Q[0]
    for dtype, paths in [("bf16", ["stream", "mma", "exact"]), ("fp32", ["stream", "exact"])]:
        db = rir.Database.from_descriptors(X.to(cuda_device), dtype)
        qr, qs = db.pack_queries(Q.to(cuda_device))
        for path in paths:
            sc, ix = db.search(qr, qs, 6, path=path)
            assert ix[0, :4].cpu().tolist() == sorted(dup), (dtype, path)
            assert float(sc[0, 0]) == float(sc[0, 3])


def _sampled_blocks(n):
    nblk = -(-n // 256)
    sblk = min(max(32, int(0.02 * nblk + 0.5)), 148)
    sblk = min(sblk, nblk - 1)
    return {(j * nblk) // sblk for j in range(sblk)}, nblk


@pytest.mark.parametrize("path", ["stream", "mma"])
def test_candidate_overflow_falls_back_to_exact_scan(cuda_device, path):
    """Adversarial row order: thousands of rows beating everything in the strided sample sit in un-sampled blocks, so
    the candidate list overflows and the one-CTA-per-query fallback must take over.  Result must still be exact."""
    n, d, k = 200000, 64, 10
    Q, X, _ = synth.retrieval_set(2, n, d, seed=99, n_pos=0)
    sampled, nblk = _sampled_blocks(n)
    free = [b for b in range(nblk - 1) if b not in sampled]
    rows = np.concatenate([np.arange(b * 256, b * 256 + 256) for b in free[:24]])  # 6144 rows > cap (2048)
    gen = torch.Generator().manual_seed(5)
    near = Q[0][None, :] + torch.randn(len(rows), d, generator=gen) * (0.3 / d ** 0.5)
    X[torch.from_numpy(rows)] = near / near.norm(dim=1, keepdim=True)
    db = rir.Database.from_descriptors(X.to(cuda_device), "bf16")
    qr, qs = db.pack_queries(Q.to(cuda_device))
    sc, ix = db.search(qr, qs, k, path=path)
    Xf, Qf = db.rows.float().cpu(), qr.float().cpu()
    _check(sc, ix, Qf, Xf, k, 1e-3)
    assert set(ix[0].cpu().tolist()) <= set(rows.tolist())


def test_rank_dropin_golden(cuda_device):
    """rank() against the reference call site run verbatim (iris_evaluate.py:379-386) on the golden inputs."""
    g = load_golden("ranking")
    ranks = rir.rank(g["q"], g["g"], normalize=True)            # fp32, full ranking, [N, nq] int64
    assert ranks.dtype == np.int64 and ranks.shape == (300, 6)
    ours = np.take_along_axis(g["sim"], ranks.T, 1)
    theirs = np.take_along_axis(g["sim"], g["ranks"], 1)
    np.testing.assert_allclose(ours, theirs, rtol=0, atol=2e-7)  # same order up to fp32-rounding ties
    top = rir.rank(g["q"], g["g"], k=10, normalize=True)
    np.testing.assert_array_equal(top.T, g["topk_idx"])
    r2, s2 = rir.rank(g["q"], g["g"], k=10, normalize=True, return_scores=True)
    np.testing.assert_allclose(s2.T, g["topk_scores"], rtol=1e-5)


def test_merge_and_single_gpu_shard_emulation(cuda_device):
    """Row-sharded search emulated on one GPU: 3 shards with idx_offset -> merge_topk == unsharded search."""
    nq, n, d, k = 7, 60000, 128, 100
    Q, X, _ = synth.retrieval_set(nq, n, d, seed=17)
    Xd, Qd = X.to(cuda_device), Q.to(cuda_device)
    whole = rir.Database.from_descriptors(Xd, "bf16")
    qr, qs = whole.pack_queries(Qd)
    want_sc, want_ix = whole.search(qr, qs, k)
    parts_s, parts_i = [], []
    for r in range(3):
        lo, hi = rir.shard_bounds(n, 3, r)
        sh = rir.Database.from_descriptors(Xd[lo:hi], "bf16", idx_offset=lo)
        s, i = sh.search(qr, qs, k)
        parts_s.append(s)
        parts_i.append(i)
    ms, mi = rir.merge_topk(torch.stack(parts_s), torch.stack(parts_i))
    assert torch.equal(mi, want_ix)
    assert torch.equal(ms, want_sc)
    # padding entries (-inf, -1) from a short shard are ignored
    parts_s[2][:, 50:] = float("-inf")
    parts_i[2][:, 50:] = -1
    ms2, mi2 = rir.merge_topk(torch.stack(parts_s), torch.stack(parts_i))
    os_, oi = S.merge_shards([p.cpu().numpy() for p in parts_s], [p.cpu().numpy() for p in parts_i], k)
    np.testing.assert_array_equal(mi2.cpu().numpy(), oi)


@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_alpha_query_expansion(cuda_device, dtype):
    nq, n, d, k, kq, alpha = 12, 25000, 256, 50, 10, 3.0
    Q, X, _ = synth.retrieval_set(nq, n, d, seed=31)
    db = rir.Database.from_descriptors(X.to(cuda_device), dtype)
    qr, qs = db.pack_queries(Q.to(cuda_device))
    sc, ix = db.search(qr, qs, k)
    q2, q2s, q2f = rir.alpha_query_expansion(db, qr, qs, sc, ix, kq=kq, alpha=alpha)
    Xf, Qf = db.rows.float().cpu(), qr.float().cpu()
    want = S.alpha_qe(Qf, Xf, sc.cpu().numpy(), ix.cpu().numpy(), kq, alpha)
    np.testing.assert_allclose(q2f.cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-6)
    sc2, ix2, _ = rir.search_with_aqe(db, qr, qs, k=k, kq=kq, alpha=alpha)
    _check(sc2, ix2, q2.float().cpu(), Xf, k, EPS[dtype])
    # sharded accumulate: two half-shards, partial sums added by hand == unsharded
    lo, hi = rir.shard_bounds(n, 2, 1)
    a = rir.Database(db.rows[:lo], None, dtype, 0)
    b = rir.Database(db.rows[lo:], None, dtype, lo)
    lib = rir.load()
    from research_image_retrieval_b200 import _lib
    acc = torch.zeros(nq, d, device=cuda_device)
    for part in (a, b):
        _lib.check(lib.rir_aqe_accumulate(part.rows.data_ptr(), _lib.RIR_BF16 if dtype == "bf16" else _lib.RIR_F32, None,
                                          part.n, part.idx_offset, d, sc.data_ptr(), ix.data_ptr(), nq, k, kq, alpha,
                                          acc.data_ptr(), None))
    torch.cuda.synchronize()
    v = Qf.to(cuda_device) + acc
    np.testing.assert_allclose((v / v.norm(dim=1, keepdim=True)).cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-6)


def test_argument_errors(cuda_device):
    X = torch.zeros(100, 64, device=cuda_device, dtype=torch.bfloat16)
    Q = torch.zeros(2, 64, device=cuda_device, dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        rir.sim_topk(Q, X, 0)
    with pytest.raises(ValueError):
        rir.sim_topk(Q, X, 101)
    with pytest.raises(ValueError):
        rir.sim_topk(Q[:, :32].contiguous(), X, 5)
    with pytest.raises(rir.RirError):  # 68 bf16 = 136 B rows: not a multiple of 16
        rir.sim_topk(torch.zeros(2, 68, device=cuda_device, dtype=torch.bfloat16),
                     torch.zeros(100, 68, device=cuda_device, dtype=torch.bfloat16), 5)
    with pytest.raises(rir.RirError):  # workspace too small
        rir.sim_topk(Q, X, 5, workspace=torch.zeros(256, dtype=torch.uint8, device=cuda_device))
    sc, ix = rir.sim_topk(Q[:0].contiguous(), X, 5)  # empty query batch is a no-op
    assert tuple(sc.shape) == (0, 5)


@pytest.mark.parametrize("nq,path", [(1, "stream"), (70, "mma")])
def test_full_size_properties(cuda_device, nq, path):
    """BASELINE cfg-2 size (1,007,323 x 2048 bf16, top-100): size-independent properties + a GPU fp32 rescore."""
    n, d, k = 1007323, 2048, 100
    gen = torch.Generator(device=cuda_device).manual_seed(1002)
    X = torch.empty(n, d, device=cuda_device, dtype=torch.bfloat16)
    for lo in range(0, n, 65536):
        blk = torch.randn(min(65536, n - lo), d, generator=gen, device=cuda_device)
        X[lo:lo + blk.shape[0]] = (blk / blk.norm(dim=1, keepdim=True)).to(torch.bfloat16)
    Q = torch.randn(nq, d, generator=gen, device=cuda_device)
    Q = Q / Q.norm(dim=1, keepdim=True)
    planted = torch.randperm(n, generator=gen, device=cuda_device)[: nq * 5].reshape(nq, 5)
    for r in range(nq):
        rows = Q[r][None] + torch.randn(5, d, generator=gen, device=cuda_device) * (0.5 / d ** 0.5)
        X[planted[r]] = (rows / rows.norm(dim=1, keepdim=True)).to(torch.bfloat16)
    db = rir.Database(X, None, "bf16")
    qr = Q.to(torch.bfloat16)
    sc, ix = db.search(qr, None, k, path=path)
    torch.cuda.synchronize()
    ixl = ix.long()
    # (1) planted near-duplicates lead the list; (2) sorted; (3) unique; (4) returned scores == fp32 rescore
    for r in range(nq):
        assert set(planted[r].tolist()) == set(ixl[r, :5].tolist())
        assert len(set(ixl[r].tolist())) == k
    assert bool((sc[:, 1:] <= sc[:, :-1]).all())
    rescored = torch.einsum("qkd,qd->qk", X[ixl].float(), qr.float())
    assert torch.allclose(sc, rescored, rtol=2e-4, atol=2e-6)
    # (5) exactness: nothing outside the list beats the k-th score by more than the bf16 tie epsilon
    _nothing_better_outside(qr.float(), X, sc, ixl, 1e-3)


def _unit_rows_gpu(n, d, gen, dev, dtype=torch.float32, chunk=131072):
    out = torch.empty(n, d, device=dev, dtype=dtype)
    for lo in range(0, n, chunk):
        blk = torch.randn(min(chunk, n - lo), d, generator=gen, device=dev)
        out[lo:lo + blk.shape[0]] = (blk / blk.norm(dim=1, keepdim=True)).to(dtype)
    return out


def _nothing_better_outside(Qf, X, sc, ixl, eps, chunk=131072):
    """Exactness at full size: every row whose fp32 score beats the k-th returned score by more than eps (relative) IS
    in the returned list (membership is checked per row, so fp32 accumulation-order noise at the bar cannot flip it)."""
    kth = sc[:, -1]
    bar = kth + eps * kth.abs()
    for lo in range(0, X.shape[0], chunk):
        s = Qf @ X[lo:lo + chunk].float().t()
        qi, ci = torch.nonzero(s > bar[:, None], as_tuple=True)
        if qi.numel():
            present = (ixl[qi] == (ci + lo)[:, None]).any(1)
            assert bool(present.all()), f"rows in [{lo}, {lo + chunk}) beat the k-th score but are not in the list"


def test_cfg3_gldv2_shape_with_alpha_qe(cuda_device):
    """BASELINE cfg-3 size: 1,129 queries x 761,757 x 512-d bf16, top-100, alpha-QE (k=10, alpha=3), re-query.
    Size-independent properties + GPU fp32 restatement of the QE formula."""
    nq, n, d, k, kq, alpha = 1129, 761757, 512, 100, 10, 3.0
    gen = torch.Generator(device=cuda_device).manual_seed(1003)
    X = _unit_rows_gpu(n, d, gen, cuda_device, torch.bfloat16)
    Q = _unit_rows_gpu(nq, d, gen, cuda_device)
    planted = torch.randperm(n, generator=gen, device=cuda_device)[: nq * 4].reshape(nq, 4)
    for r in range(0, nq, 64):  # 4 near-duplicates per query
        rows = Q[r:r + 64, None, :] + torch.randn(min(64, nq - r), 4, d, generator=gen, device=cuda_device) * (0.5 / d ** 0.5)
        X[planted[r:r + 64].reshape(-1)] = (rows / rows.norm(dim=2, keepdim=True)).reshape(-1, d).to(torch.bfloat16)
    db = rir.Database(X, None, "bf16")
    qr = Q.to(torch.bfloat16)
    sc, ix = db.search(qr, None, k)
    ixl = ix.long()
    assert bool((sc[:, 1:] <= sc[:, :-1]).all())
    assert all(len(set(ixl[r].tolist())) == k for r in range(0, nq, 37))
    assert bool((torch.sort(ixl[:, :4], 1).values == torch.sort(planted, 1).values).all())
    rescored = torch.einsum("qkd,qd->qk", X[ixl].float(), qr.float())
    assert torch.allclose(sc, rescored, rtol=2e-4, atol=2e-6)
    _nothing_better_outside(qr.float(), X, sc, ixl, 1e-3)
    # alpha-QE: q' = L2(q + sum_j max(s_j,0)^alpha x_j) on the bf16-dequantised operands, fp32
    q2, q2s, q2f = rir.alpha_query_expansion(db, qr, None, sc, ix, kq=kq, alpha=alpha)
    w = sc[:, :kq].clamp(min=0).pow(alpha)
    want = qr.float() + torch.einsum("qk,qkd->qd", w, X[ixl[:, :kq]].float())
    want = want / want.norm(dim=1, keepdim=True)
    assert torch.allclose(q2f, want, rtol=1e-4, atol=1e-6)
    gathered = X[ixl[:8, :kq].reshape(-1)].float().cpu()  # the CPU oracle on the 80 rows the first 8 queries expand with
    host = S.alpha_qe(qr[:8].float().cpu(), gathered, sc[:8].cpu().numpy(), np.arange(8 * kq).reshape(8, kq), kq, alpha)
    np.testing.assert_allclose(q2f[:8].cpu().numpy(), host.numpy(), rtol=1e-4, atol=1e-6)
    sc2, ix2 = db.search(q2, q2s, k)
    assert bool((sc2[:, 1:] <= sc2[:, :-1]).all())
    _nothing_better_outside(q2.float(), X, sc2, ix2.long(), 1e-3)
    # expansion pulls the planted neighbours even closer: they stay on top
    assert bool((torch.sort(ix2.long()[:, :4], 1).values == torch.sort(planted, 1).values).all())


def test_cfg5_fp8_shape_subsample(cuda_device):
    """BASELINE cfg-5 shape on one GPU: 1,580,470 x 2048 fp8 (+ per-row scales), top-10, 2,048-query subsample.
    Bar: indices exact except ties within 5e-3 relative, judged on fp32 scores of the UNQUANTISED rows."""
    nq, n, d, k = 2048, 1580470, 2048, 10
    gen = torch.Generator(device=cuda_device).manual_seed(1005)
    rows8 = torch.empty(n, d, device=cuda_device, dtype=torch.uint8)
    scale = torch.empty(n, device=cuda_device)
    rows16 = torch.empty(n, d, device=cuda_device, dtype=torch.bfloat16)
    Q = _unit_rows_gpu(nq, d, gen, cuda_device)
    planted = torch.randperm(n, generator=gen, device=cuda_device)[: nq * 3].reshape(nq, 3)
    inv = torch.full((n,), -1, device=cuda_device, dtype=torch.long)
    inv[planted.reshape(-1)] = torch.arange(nq * 3, device=cuda_device)
    for lo in range(0, n, 131072):
        blk = torch.randn(min(131072, n - lo), d, generator=gen, device=cuda_device)
        hit = inv[lo:lo + blk.shape[0]]
        sel = hit >= 0
        if bool(sel.any()):
            src = Q[hit[sel] // 3] + torch.randn(int(sel.sum()), d, generator=gen, device=cuda_device) * (0.5 / d ** 0.5)
            blk[sel] = src
        blk = blk / blk.norm(dim=1, keepdim=True)
        r8, s8 = rir.pack_descriptors(blk, "fp8")
        rows8[lo:lo + blk.shape[0]], scale[lo:lo + blk.shape[0]] = r8, s8
        rows16[lo:lo + blk.shape[0]] = blk.to(torch.bfloat16)
    db = rir.Database(rows8, scale, "fp8", rescore_rows=rows16)
    sc, ix = db.query(Q, k)                      # fp8 scan for 2k+16 candidates, bf16 re-score
    ixl = ix.long()
    assert bool((sc[:, 1:] <= sc[:, :-1]).all())
    assert bool((torch.sort(ixl[:, :3], 1).values == torch.sort(planted, 1).values).all())
    # fp32 truth for a 256-query subsample, chunked
    sub = torch.arange(0, nq, 8, device=cuda_device)
    best_s = torch.full((sub.numel(), k), -1e30, device=cuda_device)
    best_i = torch.zeros((sub.numel(), k), device=cuda_device, dtype=torch.long)
    for lo in range(0, n, 262144):
        s = Q[sub] @ rows16[lo:lo + 262144].float().t()
        cs, ci = torch.cat([best_s, s], 1).topk(k, dim=1)
        allidx = torch.cat([best_i, torch.arange(lo, lo + s.shape[1], device=cuda_device).expand(sub.numel(), -1)], 1)
        best_s, best_i = cs, torch.gather(allidx, 1, ci)
    got = ixl[sub].cpu().numpy()
    true_of_got = torch.einsum("qkd,qd->qk", rows16[ixl[sub]].float(), Q[sub]).cpu().numpy()
    ok, msg = S.indices_match_up_to_ties(got, true_of_got, best_i.cpu().numpy(), best_s.cpu().numpy(), 5e-3)
    assert ok, msg


def test_descriptor_store_to_sharded_search(cuda_device, tmp_path):
    """shard files on disk -> per-rank resident shards -> merged result == unsharded search."""
    nq, n, d, k = 5, 30000, 64, 20
    Q, X, _ = synth.retrieval_set(nq, n, d, seed=77)
    rows, _ = rir.pack_descriptors(X.to(cuda_device), "bf16")
    st = rir.DescriptorStore.create(str(tmp_path / "store"), rows.shape[1], "bf16")
    for a, b in [(0, 9000), (9000, 9001), (9001, 30000)]:
        st.append(rows[a:b])
    whole = rir.Database(rows, None, "bf16")
    qr, qs = whole.pack_queries(Q.to(cuda_device))
    want_sc, want_ix = whole.search(qr, qs, k)
    parts = [rir.DescriptorStore(str(tmp_path / "store")).load_database(4, r, cuda_device) for r in range(4)]
    assert [p.idx_offset for p in parts] == [0, 7500, 15000, 22500] and sum(p.n for p in parts) == n
    res = [p.search(qr, qs, k) for p in parts]
    ms, mi = rir.merge_topk(torch.stack([r[0] for r in res]), torch.stack([r[1] for r in res]))
    assert torch.equal(mi, want_ix) and torch.equal(ms, want_sc)


@pytest.mark.parametrize("dtype", ["bf16", "fp8", "fp32"])
def test_query_host_single_call(cuda_device, dtype):
    """rir_search_host: host fp32 queries in, host top-k out, one C-ABI call == pack + search + copies by hand."""
    nq, n, d, k = 9, 80000, 128, 30
    Q, X, _ = synth.retrieval_set(nq, n, d, seed=55)
    db = rir.Database.from_descriptors(X.to(cuda_device), dtype)
    qr, qs = db.pack_queries(Q.to(cuda_device))
    want_sc, want_ix = db.search(qr, qs, k)
    qh = Q.contiguous().pin_memory()
    sc, ix = db.query_host(qh, k)
    assert sc.device.type == "cpu" and ix.dtype == torch.int32
    assert torch.equal(ix, want_ix.cpu()) and torch.equal(sc, want_sc.cpu())
    out = (torch.empty(nq, k).pin_memory(), torch.empty(nq, k, dtype=torch.int32).pin_memory())
    sc2, ix2 = db.query_host(qh, k, out=out)      # result buffers re-used
    assert sc2 is out[0] and torch.equal(ix2, want_ix.cpu())
    with pytest.raises(TypeError):
        db.query_host(Q.to(cuda_device), k)


def test_clustered_database_r1m_layout(cuda_device):
    """R1M-style layout: the landmark images (every query's positives, contiguous per landmark) lead the database,
    distractors follow.  Tile 0 — always a first-phase tile of the fused scan — then holds ~80 rows per query that beat
    every distractor, far more than the 8 keys kept per tile: the select kernel must re-score it (and the tiles the
    permutation happens to pick).  Result must still be exact."""
    nq, per, n, d, k = 24, 80, 300000, 128, 100
    gen = torch.Generator(device=cuda_device).manual_seed(77)
    X = _unit_rows_gpu(n, d, gen, cuda_device)
    Q = _unit_rows_gpu(nq, d, gen, cuda_device)
    for i in range(nq):  # cluster i: rows [i*per, (i+1)*per)
        rows = Q[i][None] + torch.randn(per, d, generator=gen, device=cuda_device) * (0.6 / d ** 0.5)
        X[i * per:(i + 1) * per] = rows / rows.norm(dim=1, keepdim=True)
    db = rir.Database.from_descriptors(X, "bf16")
    qr, qs = db.pack_queries(Q)
    sc, ix = db.search(qr, qs, k, path="mma")
    ixl = ix.long()
    for i in range(nq):
        inside = ((ixl[i] >= i * per) & (ixl[i] < (i + 1) * per)).sum().item()
        assert inside == per, f"query {i}: only {inside} of its {per} cluster rows came back"
    assert bool((sc[:, 1:] <= sc[:, :-1]).all())
    Xb, Qb = db.rows.float(), qr.float()
    rescored = torch.einsum("qkd,qd->qk", Xb[ixl], Qb)
    assert torch.allclose(sc, rescored, rtol=2e-4, atol=2e-6)
    _nothing_better_outside(Qb, db.rows, sc, ixl, 1e-3)
    sc_e, ix_e = db.search(qr[:3].contiguous(), None, k, path="exact")   # independent implementation
    ok, msg = S.indices_match_up_to_ties(ix[:3].cpu().numpy().astype(np.int64), sc[:3].cpu().numpy(),
                                         ix_e.cpu().numpy().astype(np.int64), sc_e.cpu().numpy(), 1e-3)
    assert ok, msg


BOUNDARY_CASES = [
    # nq, n, d, k, dtype — mode / tile / block boundaries of the tcgen05 path
    (128, 75776, 64, 10, "bf16"),     # exactly one full query block; n == 2 * 148 tiles exactly (smallest fused shard)
    (129, 75775, 64, 10, "bf16"),     # one row short of the fused route (three-launch route), two query blocks
    (256, 80000, 64, 7, "bf16"),      # two full blocks -> one CTA pair
    (257, 80001, 64, 7, "fp8"),       # third block with a single query, n % 256 == 129
    (1, 76000, 2048, 1, "bf16"),      # k = 1, one query, full-width rows
    (3, 76000, 64, 592, "bf16"),      # largest k of the fused route (148 * 8 >= 2k)
    (3, 76000, 64, 593, "bf16"),      # first k of the three-launch route
    (2, 40000, 64, 4096, "bf16"),     # large k: dense sample; overflow redo still inside the select kernel
    (2, 40000, 64, 8192, "bf16"),     # largest k: the separate exact-scan fallback launch
    (70, 80000, 4096, 20, "bf16"),    # 8 KB rows: 64 K-chunks per tile
    (16, 77000, 16, 5, "fp8"),        # a single 16-byte chunk per row
]


@pytest.mark.parametrize("nq,n,d,k,dtype", BOUNDARY_CASES)
def test_tensor_path_boundaries(cuda_device, nq, n, d, k, dtype):
    Q, X, _ = synth.retrieval_set(nq, n, d, seed=7 * n + nq + k, n_pos=min(4, k))
    db = rir.Database.from_descriptors(X.to(cuda_device), dtype)
    qr, qs = db.pack_queries(Q.to(cuda_device))
    sc, ix = db.search(qr, qs, k, path="mma")
    torch.cuda.synchronize()
    Xf = _dequant(db.rows, db.scale, dtype)[:, :d]
    Qf = _dequant(qr, qs, dtype)[:, :d]
    _check(sc, ix, Qf, Xf, k, 1e-3)


def test_rerank_hook(cuda_device):
    """a pairwise re-ranker re-orders the head of every list; the tail stays put."""
    nq, n, d, k = 5, 30000, 64, 40
    Q, X, _ = synth.retrieval_set(nq, n, d, seed=5)
    db = rir.Database.from_descriptors(X.to(cuda_device), "bf16")
    qr, qs = db.pack_queries(Q.to(cuda_device))
    sc, ix = db.search(qr, qs, k)
    Xf = X.to(cuda_device)

    def fp32_pair_scores(q_ids, cand):                     # a stand-in "model": exact fp32 cosine of the pair
        return torch.einsum("qrd,qd->qr", Xf[cand.long()], Q.to(cuda_device)[q_ids])

    rs, ri = rir.rerank_topk(sc, ix, fp32_pair_scores, top_r=25)
    assert torch.equal(ri[:, 25:], ix[:, 25:]) and torch.equal(rs[:, 25:], sc[:, 25:])
    assert torch.equal(torch.sort(ri[:, :25], 1).values, torch.sort(ix[:, :25], 1).values)   # same candidates
    assert bool((rs[:, 1:25] <= rs[:, :24]).all())                                           # new order
    want = fp32_pair_scores(torch.arange(nq, device=cuda_device), ri[:, :25])
    assert torch.allclose(rs[:, :25], want, rtol=1e-6)
    # reversing scorer: the head comes out reversed
    rs2, ri2 = rir.rerank_topk(sc, ix, lambda q, c: -sc[:, :c.shape[1]], top_r=k)
    assert torch.equal(ri2, torch.flip(ix, dims=[1]))
    with pytest.raises(ValueError):
        rir.rerank_topk(sc, ix, lambda q, c: torch.zeros(1, device=cuda_device))


def test_workspace_reuse_across_shapes_stays_exact(cuda_device):
    """One Database, one (initialised, self-cleaning) workspace, many shapes in a row — fused scan, three-launch route
    (k > 592), a failing call, 1 query, several query groups: no memset launches in between, every answer exact."""
    n, d = 90000, 64
    Q, X, _ = synth.retrieval_set(300, n, d, seed=5150)
    db = rir.Database.from_descriptors(X.to(cuda_device), "bf16")
    qr_all, _ = db.pack_queries(Q.to(cuda_device))
    Xf, Qf = db.rows.float().cpu(), qr_all.float().cpu()
    seq = [(70, 100, "auto"), (300, 20, "mma"), (5, 700, "mma"), (1, 10, "auto"), (70, 100, "auto"), (3, 10, "stream"),
           (130, 50, "auto")]
    for i, (nq, k, path) in enumerate(seq):
        q = qr_all[:nq].contiguous()
        sc, ix = db.search(q, None, k, path=path)
        _check(sc, ix, Qf[:nq], Xf, k, 1e-3)
        if i == 2:  # a call that fails after the workspace was handed over must not poison the next one
            with pytest.raises((ValueError, rir.RirError)):
                db.search(q, None, 9000, path=path)
    # several internal query groups (nq > 4096) through the same workspace, then the small batch again
    Qb, Xb, _ = synth.retrieval_set(4500, 80000, 64, seed=77)
    dbb = rir.Database.from_descriptors(Xb.to(cuda_device), "bf16")
    qb, _ = dbb.pack_queries(Qb.to(cuda_device))
    sc, ix = dbb.search(qb, None, 10)
    a = dbb.search(qb[:64].contiguous(), None, 10)
    assert torch.equal(ix[:64], a[1]) and torch.equal(sc[:64], a[0])
    _check(sc[4090:4110], ix[4090:4110], qb.float().cpu()[4090:4110], dbb.rows.float().cpu(), 10, 1e-3)


def test_plain_abi_call_on_dirty_workspace(cuda_device):
    """rir_sim_topk WITHOUT RIR_WS_CLEAN must not depend on the workspace contents (it initialises the header itself)."""
    nq, n, d, k = 70, 100003, 128, 100
    Q, X, _ = synth.retrieval_set(nq, n, d, seed=31337)
    db = rir.Database.from_descriptors(X.to(cuda_device), "bf16")
    qr, _ = db.pack_queries(Q.to(cuda_device))
    want = db.search(qr, None, k)
    need = rir.load().rir_sim_topk_workspace(nq, n, d, k, 1)
    ws = torch.randint(0, 255, (need,), dtype=torch.uint8, device=cuda_device)   # garbage
    for _ in range(2):
        got = rir.sim_topk(qr, db.rows, k, dtype="bf16", workspace=ws)
        assert torch.equal(got[1], want[1]) and torch.equal(got[0], want[0])
        ws[: 1 << 16].random_(0, 255)  # dirty the header again between calls


def test_concurrent_searches_on_two_streams(cuda_device):
    """The fused scan spins on a grid barrier: it is a cooperative launch, so two searches in flight on two streams
    (each needs every SM) serialise at the launch instead of dead-locking.  Results must be exact on both."""
    nq, n, d, k = 70, 160000, 256, 100
    Q, X, _ = synth.retrieval_set(2 * nq, n, d, seed=2024)
    db = rir.Database.from_descriptors(X.to(cuda_device), "bf16")
    qr, _ = db.pack_queries(Q.to(cuda_device))
    qa, qb = qr[:nq].contiguous(), qr[nq:].contiguous()
    want_a, want_b = db.search(qa, None, k), db.search(qb, None, k)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for rep in range(6):
        with torch.cuda.stream(s1):
            ra = db.search(qa, None, k)
        with torch.cuda.stream(s2):
            rb = db.search(qb, None, k)
        outs.append((ra, rb))
    torch.cuda.synchronize()
    for ra, rb in outs:
        assert torch.equal(ra[1], want_a[1]) and torch.equal(ra[0], want_a[0])
        assert torch.equal(rb[1], want_b[1]) and torch.equal(rb[0], want_b[0])


def test_merge_topk_with_fewer_valid_entries_than_k(cuda_device):
    """Mostly-padding lists: fewer than k real entries among G*k > k inputs (radix select ends at prefix 0) — the real
    ones must all come out, in order, followed by (-inf, -1)."""
    for G, nq, k, valid in [(4, 3, 100, 7), (8, 2, 1000, 3), (2, 5, 32, 0), (3, 2, 400, 150)]:
        gen = torch.Generator().manual_seed(G * 1000 + k)
        sc = torch.full((G, nq, k), float("-inf"))
        ix = torch.full((G, nq, k), -1, dtype=torch.int32)
        for g in range(G):
            v = torch.sort(torch.rand(nq, valid, generator=gen), dim=1, descending=True).values
            sc[g, :, :valid] = v
            ix[g, :, :valid] = torch.arange(valid, dtype=torch.int32)[None, :] + g * 100000
        ms, mi = rir.merge_topk(sc.to(cuda_device), ix.to(cuda_device))
        ws, wi = S.merge_shards([sc[g].numpy() for g in range(G)], [ix[g].numpy() for g in range(G)], k)
        np.testing.assert_array_equal(mi.cpu().numpy(), wi)
        tot = min(k, G * valid)
        np.testing.assert_array_equal(ms.cpu().numpy()[:, :tot], ws[:, :tot])
        assert bool((mi[:, tot:] == -1).all()) and bool(torch.isinf(ms[:, tot:]).all())


def test_profile_scan_hook_counts_every_query_group(cuda_device):
    """rir_profile_scan_begin/end: one event pair per scan launch — a 4500-query search (two internal groups) records
    two launches per call, with plausible durations."""
    import ctypes
    Q, X, _ = synth.retrieval_set(4500, 80000, 64, seed=78)
    db = rir.Database.from_descriptors(X.to(cuda_device), "bf16")
    qr, _ = db.pack_queries(Q.to(cuda_device))
    db.search(qr, None, 10)
    lib = rir.load()
    lib.rir_profile_scan_begin()
    for _ in range(3):
        db.search(qr, None, 10)
    buf, n = (ctypes.c_float * 16)(), ctypes.c_int(0)
    assert lib.rir_profile_scan_end(buf, 16, ctypes.byref(n)) == 0
    assert n.value == 6
    ms = [buf[i] for i in range(6)]
    assert all(0.0 < m < 50.0 for m in ms)
    lib.rir_profile_scan_begin()       # disarmed after end: a fresh begin/end with no search records nothing
    assert lib.rir_profile_scan_end(buf, 16, ctypes.byref(n)) == 0 and n.value == 0


def test_host_query_pipeline_matches_sync_call(cuda_device):
    """HostQueryPipeline (copy stream + two batches in flight) returns exactly what the synchronous host call returns,
    batch after batch, with re-used staging slots."""
    nq, n, d, k = 70, 100003, 128, 100
    Q, X, _ = synth.retrieval_set(4 * nq, n, d, seed=606)
    db = rir.Database.from_descriptors(X.to(cuda_device), "bf16")
    batches = [Q[i * nq:(i + 1) * nq].contiguous().pin_memory() for i in range(4)]
    want = [tuple(t.clone() for t in db.query_host(b, k)) for b in batches]
    pipe = rir.HostQueryPipeline(db, nq, k, depth=2)
    pending, got = [], []
    for rep in range(3):
        for i, b in enumerate(batches):
            pending.append((i, pipe.submit(b)))
            if len(pending) == 2:
                j, h = pending.pop(0)
                sc, ix = h.result()
                got.append((j, sc.clone(), ix.clone()))
    for j, h in pending:
        sc, ix = h.result()
        got.append((j, sc.clone(), ix.clone()))
    assert len(got) == 12
    for j, sc, ix in got:
        assert torch.equal(ix, want[j][1]) and torch.equal(sc, want[j][0])


@pytest.mark.parametrize("n", [125916, 76000, 100003, 251831])
def test_fused_scan_range_mode_idle_tail_ctas(cuda_device, n):
    """Range mode gives every CTA ceil(n / 148) rows rounded up to 32: for these sizes the last CTA(s) own NO real row.
    Their first-phase slots must read as empty — after a search on another database left real keys in the workspace."""
    nq, d, k = 70, 64, 100
    Q, X, _ = synth.retrieval_set(nq, n, d, seed=n)
    db = rir.Database.from_descriptors(X.to(cuda_device), "bf16")
    qr, _ = db.pack_queries(Q.to(cuda_device))
    # poison the workspace: first search a LARGER database through the same workspace whose extra rows (exactly the
    # ranges of the CTAs that are idle for `db`) are copies of every query — their first-phase slots then hold keys
    # with score ~1.0, far above anything in `db`
    per = -(-n // 148)
    R = -(-per // 32) * 32
    if R % 256 and R % 256 < 64:
        R += 64 - R % 256
    n_big = R * 148
    assert n_big - n >= R, "this size leaves no idle CTA: pick another n"
    extra = qr[torch.arange(n_big - n, device=cuda_device) % nq]
    big = rir.Database(torch.cat([db.rows, extra]), None, "bf16")
    big.search(qr, None, k, path="mma")
    db._ws = big._ws
    sc, ix = db.search(qr, None, k, path="mma")
    _check(sc, ix, qr.float().cpu(), db.rows.float().cpu(), k, 1e-3)


def test_rank_fp32_on_a_large_gallery(cuda_device):
    """rank(dtype='fp32') beyond 16,384 rows: bf16 candidates + fp32 re-score — fp32 scores, reference order; a full
    ranking of such a gallery is refused with a pointer to revisited_map_full."""
    nq, n, d, k = 70, 60000, 256, 100
    Q, X, _ = synth.retrieval_set(nq, n, d, seed=77077)
    ranks, sc = rir.rank(Q, X, k=k, return_scores=True)
    assert ranks.shape == (k, nq) and ranks.dtype == np.int64
    ref_sc, ref_ix = S.topk(Q, X, k)
    got = ranks.T
    true_of_got = np.stack([(X[torch.from_numpy(got[r])] @ Q[r]).numpy() for r in range(nq)])
    np.testing.assert_allclose(sc.T, true_of_got, rtol=1e-5, atol=1e-6)      # fp32 scores
    ok, msg = S.indices_match_up_to_ties(got, true_of_got, ref_ix, ref_sc, 1e-5)
    assert ok, msg
    with pytest.raises(ValueError, match="revisited_map_full"):
        rir.rank(Q, X)
