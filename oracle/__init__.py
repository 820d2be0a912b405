"""oracle/ — TEST INFRASTRUCTURE ONLY.

A CPU restatement (numpy / torch-fp32) of the reference algorithms on the retrieval hot path, used as the checker by
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`.  Nothing under
`research_image_retrieval_b200/` imports this package: the product path is CUDA-only and fails loudly without
librir.so.

Pinning ("is the oracle itself right?"):
  * `oracle/ref_adapter.py` imports the UNMODIFIED reference functions from /root/reference/src/benchmark when that
    tree is present (the build container) and `tests/test_oracle_vs_reference.py` compares the restatement with them;
  * `oracle/make_golden.py` ran those same reference functions on seeded inputs and committed inputs + outputs under
    `tests/golden/` — the fixtures travel to the GPU box where /root/reference does not exist;
  * the reference has NO tests or golden vectors of its own for this path, and NO implementation of
    top-k truncation semantics, alpha query expansion or the sharded merge: for those three the oracle is a
    restatement of the published formula only ("parity unpinned");
  * `oracle/topk_logic.py` is not an oracle of the reference at all: it emulates the ranking-key / fused-threshold /
    first-phase-redo LOGIC of the tcgen05 scan in numpy, so that the exactness argument is property-tested on the CPU
    (`tests/test_topk_logic.py`).
"""
