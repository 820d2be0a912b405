"""CPU, world_size 2, gloo: the multi-rank plumbing of the sharded search (shard bounds -> local top-k ->
all-gather of packed candidates -> k-way merge).  The local search here is the oracle (there is no CPU product path);
on GPUs the same gather feeds rir_merge_topk."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import search_oracle as S
from oracle import synth
from research_image_retrieval_b200 import search


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Q, X, _ = synth.retrieval_set(6, n, 32, seed=77)
        lo, hi = search.shard_bounds(n, world, rank)
        sc, ix = S.topk(Q, X[lo:hi], min(k, hi - lo), idx_offset=lo)
        sc_t, ix_t = search.pad_topk(torch.from_numpy(sc), torch.from_numpy(ix.astype(np.int32)), k)
        all_s, all_i = search.gather_topk(sc_t, ix_t, world)
        ms, mi = search.merge_topk_host(all_s.numpy(), all_i.numpy(), k)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), ms=ms, mi=mi)
    finally:
        dist.destroy_process_group()


def test_sharded_search_world2(tmp_path):
    n, k, world = 1501, 20, 2
    mp.spawn(_worker, args=(world, _free_port(), n, k, str(tmp_path)), nprocs=world, join=True)
    Q, X, _ = synth.retrieval_set(6, n, 32, seed=77)
    want_s, want_i = S.topk(Q, X, k)
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        np.testing.assert_array_equal(z["mi"], want_i)
        np.testing.assert_allclose(z["ms"], want_s, rtol=1e-6)


def test_short_last_shard_world2(tmp_path):
    # 25 rows over 2 ranks with k=20: the second shard holds only 12 rows and pads with (-inf, -1)
    n, k, world = 25, 20, 2
    mp.spawn(_worker, args=(world, _free_port(), n, k, str(tmp_path)), nprocs=world, join=True)
    Q, X, _ = synth.retrieval_set(6, n, 32, seed=77)
    want_s, want_i = S.topk(Q, X, k)
    z = np.load(tmp_path / "r1.npz")
    np.testing.assert_array_equal(z["mi"], want_i)


def _count_worker(rank, world, port, n, out_dir):
    """Count-based positions over row shards: local counts (oracle) -> all-reduce(SUM) over gloo == positions in the
    full ranking."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Q, X, _ = synth.retrieval_set(3, n, 32, seed=91)
        X[5] = X[700]                                           # an exact tie between two ground-truth rows
        ids = np.array([5, 700, 33, n - 1, 901], dtype=np.int64)
        lo, hi = search.shard_bounds(n, world, rank)
        sim = S.similarity(Q, X)[:, lo:hi]                      # this shard's scores (one arithmetic for rows AND ids)
        # scores of the ids: taken by the shard that owns the row, all-reduced (exactly one owner per id)
        sc = torch.zeros(3, len(ids))
        for j, p in enumerate(ids):
            if lo <= p < hi:
                sc[:, j] = sim[:, p - lo]
        dist.all_reduce(sc)
        cnt = torch.zeros(3, len(ids), dtype=torch.int64)
        for qi in range(3):
            cnt[qi] = torch.from_numpy(S.count_outranking(sim[qi].numpy(), lo, ids, sc[qi].numpy()))
        dist.all_reduce(cnt)
        np.savez(os.path.join(out_dir, f"c{rank}.npz"), pos=cnt.numpy())
    finally:
        dist.destroy_process_group()


def test_count_based_positions_world2(tmp_path):
    n, world = 1203, 2
    mp.spawn(_count_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    Q, X, _ = synth.retrieval_set(3, n, 32, seed=91)
    X[5] = X[700]
    ids = np.array([5, 700, 33, n - 1, 901], dtype=np.int64)
    ranks = S.full_rank(Q, X)                                    # [nq, n], ties -> lower index
    for r in range(world):
        pos = np.load(tmp_path / f"c{r}.npz")["pos"]
        for qi in range(3):
            where = {int(v): j for j, v in enumerate(ranks[qi])}
            assert [where[int(p)] for p in ids] == pos[qi].tolist()
    assert abs(int(pos[0][0]) - int(pos[0][1])) == 1 and pos[0][0] < pos[0][1]   # the tie: row 5 right before row 700
