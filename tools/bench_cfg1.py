#!/usr/bin/env python
"""BASELINE configs[0] (the reference's own CPU-runnable case): ROxford5k-shape synthetic — 70 queries x 4,993 database
rows x 2048-d fp32, full cosine ranking + revisited-protocol mAP (Easy / Medium / Hard).  Times the drop-in calls
(`rir.rank` + `rir.compute_map_and_print`, host tensors in, host results out) next to the reference path restated by the
oracle (torch.mm + np.argsort + compute_map x3) on the host cores, and checks that both give the same E/M/H numbers."""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import research_image_retrieval_b200 as rir  # noqa: E402
from oracle import evaluate_oracle as E  # noqa: E402
from oracle import search_oracle as S  # noqa: E402
from oracle import synth  # noqa: E402


def best_of(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts), out


def main():
    nq, n, d = 70, 4993, 2048
    Q, X, _ = synth.retrieval_set(nq, n, d, seed=1001)
    gnd = synth.revisited_gnd(nq, n, seed=1001, n_empty_easy=2)
    torch.set_num_threads(os.cpu_count() or 1)

    def ours():
        ranks = rir.rank(Q, X)                       # fp32 (the reference's arithmetic), full ranking, [N, nq] int64
        with contextlib.redirect_stdout(io.StringIO()):
            return rir.compute_map_and_print("roxford5k", "bench", "test", ranks, gnd)

    def reference():
        sim = S.similarity(Q, X).numpy()
        ranks = np.argsort(-sim, axis=1)
        return E.compute_map_and_print_values(ranks.T, gnd)

    dev = torch.device("cuda", 0)
    db = rir.Database.from_descriptors(X.to(dev), "fp32")       # gallery descriptors resident (as after extraction)
    Qp = Q.contiguous().pin_memory()

    def ours_resident():
        sc, ix = db.query(Qp.to(dev, non_blocking=True), k=n)
        ranks = ix.t().contiguous()                              # [N, nq] on the device: compute_map accepts tensors
        with contextlib.redirect_stdout(io.StringIO()):
            return rir.compute_map_and_print("roxford5k", "bench", "test", ranks, gnd)

    t_res, got_res = best_of(ours_resident)
    t_ours, got = best_of(ours)
    t_ref, want = best_of(reference, reps=3)
    t_rank, _ = best_of(lambda: rir.rank(Q, X))
    same = tuple(float(x) for x in got) == tuple(float(x) for x in want)
    print(json.dumps({"workload": "cfg-1 ROxford5k-shape: 70 x 4993 x 2048 fp32, full ranking + revisited mAP (E/M/H)",
                      "ours_ms": t_ours * 1e3, "ours_gallery_resident_ms": t_res * 1e3, "ours_rank_only_ms": t_rank * 1e3, "reference_cpu_ms": t_ref * 1e3,
                      "cores": os.cpu_count(), "speedup": t_ref / t_ours, "mapE_M_H": [float(x) for x in got],
                      "identical_2dp_maps": same}))
    assert same and tuple(float(x) for x in got_res) == tuple(float(x) for x in want), (got, got_res, want)


if __name__ == "__main__":
    main()
