"""Property tests (CPU) of the ranking-key and fused-threshold logic the tcgen05 scan implements — the exactness
argument of DESIGN.md §4.1, on the numpy emulation in oracle/topk_logic.py."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import topk_logic as T


def _exact_rows(scores, k):
    order = np.lexsort((np.arange(scores.size), -scores.astype(np.float64)))   # score desc, row asc
    return order[:k]


@given(st.lists(st.floats(width=32, allow_nan=False, allow_infinity=False), min_size=2, max_size=60))
@settings(max_examples=200, deadline=None)
def test_ordered_bits_are_monotonic_and_round_trip(vals):
    v = np.array(vals, dtype=np.float32)
    o = T.float_to_ordered(v)
    back = T.ordered_to_float(o)
    assert np.array_equal(back, v + np.float32(0.0))
    i, j = np.argmax(v), np.argmin(v)
    assert o[i] >= o[j]
    order_by_bits = np.argsort(o, kind="stable")
    assert np.all(np.diff(v[order_by_bits].astype(np.float64)) >= 0)


def test_key_order_is_score_desc_then_row_asc():
    s = np.array([0.5, 0.5, -0.0, 0.0, 1.0, -1.0], dtype=np.float32)
    keys = T.make_key(s, np.arange(6))
    assert T.key_row(np.sort(keys)[::-1]).tolist() == [4, 0, 1, 2, 3, 5]      # -0 == +0: lower row first
    assert np.array_equal(T.key_score(keys), s + np.float32(0.0))


@given(st.integers(0, 2 ** 31), st.integers(1, 40), st.integers(40, 400))
@settings(max_examples=150, deadline=None)
def test_fused_tau_is_a_lower_bound_with_k_witnesses(seed, k, m):
    rng = np.random.default_rng(seed)
    scores = rng.normal(0, 0.05, m).astype(np.float32)
    kept = T.make_key(scores, rng.permutation(m))
    kept[rng.random(m) < 0.1] = 0                                    # unwritten slots
    tau = T.fused_tau(kept, k)
    live = np.sort(T.key_score(kept[kept != 0]))[::-1]
    if live.size < k:
        assert tau == -np.inf
        return
    assert (live >= tau).sum() >= k                                    # k witnesses: tau <= the k-th best kept score
    assert tau <= live[k - 1]
    kth_bits = int(T.float_to_ordered(np.array([live[k - 1]]))[0])     # and tight: same 24-bit bucket as the k-th best
    assert int(T.float_to_ordered(np.array([tau]))[0]) == kth_bits & 0xFFFFFF00


@pytest.mark.parametrize("layout", ["iid", "clustered_first_tile", "clustered_everywhere", "duplicates"])
@pytest.mark.parametrize("k", [1, 10, 100])
def test_fused_scan_equals_exact_topk(layout, k):
    rng = np.random.default_rng(11 * k + len(layout))
    n, grid = 40 * T.TILE + 77, 12                                     # 41 tiles (the last one partial), 12 "CTAs"
    scores = rng.normal(0, 0.03, n).astype(np.float32)
    if layout == "clustered_first_tile":
        scores[5:5 + 150] += 0.5                                       # 150 strong rows in tile 0 (a first-phase tile)
    elif layout == "clustered_everywhere":
        scores[: 3 * T.TILE] += 0.5                                    # strong head of the database
        scores[-60:] += 0.6                                            # and tail (partial tile)
    elif layout == "duplicates":
        scores[rng.integers(0, n, 500)] = np.float32(0.25)             # 500 exact ties
    ntiles = -(-n // T.TILE)
    mul = ntiles // grid
    while np.gcd(mul, ntiles) != 1:
        mul += 1
    first = [(v * mul) % ntiles for v in range(grid)]                  # the kernel's multiplicative permutation
    rows, ncand, redo = T.fused_scan_topk(scores, k, first)
    assert rows.tolist() == _exact_rows(scores, k).tolist()
    assert ncand >= k
    if layout == "clustered_first_tile" and k >= 10:
        assert redo >= 1                                               # tile 0 had to be re-scored
