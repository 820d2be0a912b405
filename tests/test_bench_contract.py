"""bench.py's reference arm runs on the CPU: check the JSON line it prints against the contract (keys, meanings)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600,
                       cwd=ROOT, env=e)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_json_line():
    out = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--nq", "8", "--n", "20000", "--d", "64"])
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["data"] == "synthetic"
    assert d["value"] > 0 and abs(d["ms_per_step"] * 1e-3 * d["value"] - 8) < 1e-6 * 8 + 1e-9   # nq / step time
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert "workload" in d["config"] and "model" not in d["config"]
    assert "development override" in d["metric"]          # not the BASELINE workload: the metric name says so


def test_reference_arm_other_ranks_exit_quietly():
    out = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--gpus", "2", "--nq", "4", "--n", "5000", "--d", "32"],
               env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.strip() == ""


def test_default_workload_is_baseline_cfg2():
    sys.path.insert(0, ROOT)
    import bench
    assert (bench.N_DB, bench.DIM, bench.NQ, bench.TOPK) == (1_007_323, 2048, 70, 100)
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        b = json.load(f)
    assert "1M" in b["metric"] and "2048" in b["metric"] and "top-100" in b["metric"]

    class A:
        n, d, k, dtype, nq, gpus = bench.N_DB, bench.DIM, bench.TOPK, "bf16", bench.NQ, 1
    assert bench.metric_name(A) == bench.METRIC
    assert "configs[1]" in bench.workload_config(A)["workload"]


def test_both_arms_print_the_same_config_dict():
    """The driver compares the two arms' `config`: nothing arm-specific may live there."""
    sys.path.insert(0, ROOT)
    import bench

    class A:
        n, d, k, dtype, nq, gpus = bench.N_DB, bench.DIM, bench.TOPK, "bf16", bench.NQ, 4
    c = bench.workload_config(A)
    assert set(c) == {"workload", "nq", "n_db", "dim", "k", "parallelism", "l2_note"}
    out = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--nq", "4", "--n", "3000", "--d", "32"])
    d = json.loads([l for l in out.splitlines() if l.startswith("{")][0])
    assert set(d["config"]) == set(c)
    assert "scaled" not in d["cpu_baseline"]["sample"] and "full configuration" in d["cpu_baseline"]["sample"]


def test_tie_aware_mismatch_rules():
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    ref_sc = np.array([[0.9, 0.8, 0.8000001, 0.5]], dtype=np.float32)
    ref_ix = np.array([[3, 7, 9, 1]])
    ok, n, _ = bench.tie_aware_mismatch(ref_sc.copy(), ref_ix.copy(), ref_sc, ref_ix, 1e-3)
    assert ok and n == 0
    swapped = ref_ix.copy()
    swapped[0, 1], swapped[0, 2] = 9, 7                    # a permutation inside a tie is fine
    ok, n, _ = bench.tie_aware_mismatch(ref_sc.copy(), swapped, ref_sc, ref_ix, 1e-3)
    assert ok and n == 2
    wrong = ref_ix.copy()
    wrong[0, 0] = 5                                        # a different row with no tie around: not fine
    ok, _, msg = bench.tie_aware_mismatch(ref_sc.copy(), wrong, ref_sc, ref_ix, 1e-3)
    assert not ok and "without a tie" in msg
    off = ref_sc.copy()
    off[0, 3] = 0.51                                       # score beyond tolerance
    ok, _, msg = bench.tie_aware_mismatch(off, ref_ix.copy(), ref_sc, ref_ix, 1e-3)
    assert not ok and "scores differ" in msg
