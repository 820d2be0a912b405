"""Worker for tests/test_gpu_multi.py — run under torchrun, one rank per GPU.

Row-sharded search: NVLink peer-memory exchange (rir_sim_topk_sharded) vs the NCCL all-gather + rir_merge_topk path vs
the CPU oracle on the unsharded set.  Exits non-zero on any mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import research_image_retrieval_b200 as rir  # noqa: E402
from oracle import search_oracle as S  # noqa: E402
from oracle import synth  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cases = [
        # nq, n, d, k, dtype
        (70, 200003, 256, 100, "bf16"),   # fused tcgen05 scan per shard at 2 ranks
        (3, 50000, 128, 10, "bf16"),
        (9, 150, 64, 100, "fp32"),        # shards shorter than k: padded lists
        (130, 90000, 64, 20, "fp8"),
        (5000, 80000, 64, 10, "bf16"),    # more than one internal query group (4096) through the exchange
        (70, 150000 * world, 128, 100, "bf16"),  # every shard large enough for the fused scan at any world size
        (4, 37, 64, 100, "bf16"),         # fewer rows than k in TOTAL: k clamps to the global row count
    ]
    nq_max, k_max = 5000, 100
    first = None
    for ci, (nq, n, d, k, dtype) in enumerate(cases):
        Q, X, _ = synth.retrieval_set(nq, n, d, seed=100 + ci)
        lo, hi = rir.shard_bounds(n, world, rank)
        db = rir.Database.from_descriptors(X[lo:hi].to(dev), dtype, idx_offset=lo)
        qr, qs = db.pack_queries(Q.to(dev))
        sdb = rir.ShardedDatabase(db)
        want_sc, want_ix = sdb.search(qr, qs, min(k, n))           # NCCL all-gather + merge
        assert sdb.enable_peer_exchange(nq_max, k_max)
        for rep in range(3):                                         # epochs 1..3 (both parities, buffer re-use)
            sc, ix = sdb.search(qr, qs, min(k, n))
            torch.cuda.synchronize()
            assert torch.equal(ix, want_ix), f"case {ci} rep {rep}: peer exchange indices differ from the NCCL path"
            assert torch.equal(sc, want_sc), f"case {ci} rep {rep}: peer exchange scores differ from the NCCL path"
        # asynchronous exchange (merge on the side stream): two searches outstanding, then mixed with synchronous ones
        outs = [(torch.empty_like(want_sc), torch.empty_like(want_ix)) for _ in range(2)]
        pend = []
        for rep in range(5):
            pend.append(sdb.search_async(qr, qs, min(k, n), out=outs[rep & 1]))
            if len(pend) == 2:
                a_sc, a_ix = pend.pop(0).wait()
                torch.cuda.current_stream().synchronize()
                assert torch.equal(a_ix, want_ix) and torch.equal(a_sc, want_sc), f"case {ci} rep {rep}: async exchange differs"
        a_sc, a_ix = pend.pop(0).result()
        assert torch.equal(a_ix, want_ix) and torch.equal(a_sc, want_sc), f"case {ci}: async exchange (host sync) differs"
        sc, ix = sdb.search(qr, qs, min(k, n))                       # synchronous call right behind asynchronous ones
        assert torch.equal(ix, want_ix) and torch.equal(sc, want_sc), f"case {ci}: sync after async differs"
        # a smaller batch afterwards (stale inbox rows must not leak in)
        q2, s2 = qr[:2].contiguous(), None if qs is None else qs[:2].contiguous()
        sc2, ix2 = sdb.search(q2, s2, min(k, n))
        w_sc2, w_ix2 = sdb.search(q2, s2, min(k, n), exchange="nccl")
        assert torch.equal(ix2, w_ix2) and torch.equal(sc2, w_sc2), f"case {ci}: 2-query batch differs"
        if dtype != "fp32" or True:  # the host-buffer call over the exchange: same answer, CPU tensors
            if db.d == d:
                hs, hi_ = sdb.query_host(Q.contiguous().pin_memory(), min(k, n))
                assert torch.equal(hi_, want_ix.cpu()) and torch.equal(hs, want_sc.cpu()), f"case {ci}: query_host differs"
        if rank == 0:  # oracle on the unsharded, de-quantised set
            whole = rir.Database.from_descriptors(X.to(dev), dtype)
            Xf = whole.rows.cpu()
            if dtype == "fp8":
                Xf = Xf.view(torch.float8_e4m3fn).float() * whole.scale.cpu()[:, None]
            Qf = qr.cpu()
            if dtype == "fp8":
                Qf = Qf.view(torch.float8_e4m3fn).float() * qs.cpu()[:, None]
            Xf, Qf = Xf.float()[:, :d], Qf.float()[:, :d]
            kk = min(k, n)
            ref_sc, ref_ix = S.topk(Qf, Xf, kk)
            got = ix.cpu().numpy().astype(np.int64)
            sim_of_got = np.stack([(Xf[torch.from_numpy(got[r])] @ Qf[r]).numpy() for r in range(nq)])
            ok, msg = S.indices_match_up_to_ties(got, sim_of_got, ref_ix, ref_sc, 1e-3 if dtype != "fp32" else 1e-5)
            assert ok, f"case {ci}: {msg}"
        sdb.close()
        first = first or True
    # full-protocol mAP by counting: every shard counts the rows that outrank each ground-truth id, two
    # all-reduces join them — must equal the unsharded evaluation bit for bit (same per-row arithmetic, integer counts)
    nq, n, d = 12, 60000, 64
    Q, X, _ = synth.retrieval_set(nq, n, d, seed=555)
    gnd = synth.revisited_gnd(nq, n, seed=556, n_empty_easy=1)
    lo, hi = rir.shard_bounds(n, world, rank)
    db = rir.Database.from_descriptors(X[lo:hi].to(dev), "bf16", idx_offset=lo)
    sdb = rir.ShardedDatabase(db)
    qr, qs = db.pack_queries(Q.to(dev))
    got = rir.revisited_map_full(sdb, qr, qs, gnd)
    whole = rir.Database.from_descriptors(X.to(dev), "bf16")
    want = rir.revisited_map_full(whole, qr, qs, gnd)
    for a, b in zip(got, want):
        for x, y in zip(a, b):
            np.testing.assert_array_equal(np.asarray(x), np.asarray(y))
    dist.barrier()
    if rank == 0:
        print("multi-gpu worker OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
