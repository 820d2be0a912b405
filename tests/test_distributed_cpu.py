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
