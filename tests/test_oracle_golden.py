"""CPU: the oracle restatement against the committed golden fixtures (outputs of the UNMODIFIED reference)."""
import numpy as np
import torch

from conftest import csr_to_lists, load_golden
from oracle import descriptor_oracle as D
from oracle import evaluate_oracle as E
from oracle import search_oracle as S


def _gnd(g, keys):
    n = len(g[f"{keys[0]}_off"]) - 1
    lists = {k: csr_to_lists(g[f"{k}_ids"], g[f"{k}_off"]) for k in keys}
    return [{k: lists[k][i] for k in keys} for i in range(n)]


def test_map_full_bit_exact():
    g = load_golden("map_full")
    gnd = _gnd(g, ["easy", "hard", "junk"])
    res = E.compute_map_revisited(g["ranks"], gnd, [1, 5, 10])
    for name, (m, aps, pr, prs) in zip("EMH", res):
        assert m == float(g[f"map_{name}"])
        np.testing.assert_array_equal(aps, g[f"aps_{name}"])
        np.testing.assert_array_equal(pr, g[f"pr_{name}"])
        np.testing.assert_array_equal(prs, g[f"prs_{name}"])
    e, m, h = E.compute_map_and_print_values(g["ranks"], gnd)
    assert (e, m, h) == (float(g["mapE"]), float(g["mapM"]), float(g["mapH"]))
    for name, gt in zip("EMH", E.revisited_gnd(gnd)):
        m2, aps2 = E.compute_map(g["ranks"], gt)
        assert m2 == float(g[f"map_nokeep_{name}"])
        np.testing.assert_array_equal(aps2, g[f"aps_nokeep_{name}"])


def test_map_truncated_and_ragged():
    g = load_golden("map_truncated")
    gnd = _gnd(g, ["ok", "junk"])
    m, aps = E.compute_map(g["ranks"], gnd)
    assert m == float(g["map"])
    np.testing.assert_array_equal(aps, g["aps"])
    off = np.concatenate([[0], np.cumsum(g["ragged_len"])])
    ragged = [list(g["ragged_flat"][off[i]:off[i + 1]]) for i in range(len(gnd))]
    m, aps = E.compute_map(ragged, gnd, li=True)
    assert m == float(g["map_li"])
    np.testing.assert_array_equal(aps, g["aps_li"])


def test_map_known_answers():
    g = load_golden("map_kat")
    assert E.ap_from_adjusted_ranks([0, 1, 2], 3) == float(g["ap_012_3"]) == 1.0
    assert E.ap_from_adjusted_ranks([1, 3], 2) == float(g["ap_13_2"])
    m, aps, pr, prs = E.compute_map(np.arange(10).reshape(10, 1), [{"ok": [0, 3], "junk": [1]}], [1, 5])
    assert m == float(g["k1_map"])
    np.testing.assert_array_equal(pr, g["k1_pr"])
    m, aps, pr, prs = E.compute_map(np.array([[3], [9], [1], [4]]), [{"ok": [1, 2], "junk": [9]}], [1, 5])
    assert m == float(g["k2_map"]) == 0.125
    np.testing.assert_array_equal(pr, g["k2_pr"])
    m, aps, pr, prs = E.compute_map(np.array([[0, 0], [1, 1]]), [{"ok": []}, {"ok": [0]}], [1])
    assert m == float(g["k3_map"]) == 1.0
    assert np.isinf(aps[0]) and aps[1] == 1.0
    m, aps = E.compute_map(np.array([[0], [1]]), [{"ok": [1]}])
    assert m == float(g["k4_map"]) == 0.25
    m, aps = E.compute_map([[5, 6, 7]], [{"ok": [1], "junk": []}], li=True)
    assert m == float(g["k5_map"]) == 0.0


def test_pooling_heads():
    g = load_golden("pooling")
    x, x2 = torch.from_numpy(g["x"]), torch.from_numpy(g["x2"])
    rt = dict(rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(D.gem(x).numpy(), g["gem_p3"], **rt)
    np.testing.assert_allclose(D.gem(x2).numpy(), g["gem_p3_x2"], **rt)
    np.testing.assert_allclose(D.gem(x, 2.5).numpy(), g["gem_p2p5"], **rt)
    np.testing.assert_allclose(D.spoc(x).numpy(), g["spoc"], **rt)
    np.testing.assert_allclose(D.gem(x).numpy(), g["gempooling_p3"], **rt)
    np.testing.assert_allclose(D.gem(x2, 4.2).numpy(), g["gempooling_p4p2"], **rt)
    np.testing.assert_allclose(D.g2(x, 3.0, 1e-6, 1.25, -0.05).numpy(), g["g2"], **rt)
    np.testing.assert_allclose(D.mac(x).numpy().reshape(3, 24, 1), g["spp_max_l1"], **rt)
    W, b = torch.from_numpy(g["W"]), torch.from_numpy(g["b"])
    np.testing.assert_allclose(D.head(x, "gem", W=W, b=b).numpy(), g["gem_tail"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(D.head(x, "gem", W=W, b=b, l2_before_whiten=True).numpy(), g["solar_tail"], rtol=1e-4,
                               atol=1e-6)
    # PCA learn: eigenvectors are defined up to sign -> compare |W| rows and the whitened covariance
    W2, b2 = D.whitening_layer_from_pca(g["des"].astype(np.float64), 12)
    np.testing.assert_allclose(np.abs(W2.numpy()), np.abs(g["W"]), rtol=2e-3, atol=2e-4)


def test_ranking_matches_reference_call_site():
    g = load_golden("ranking")
    q = torch.nn.functional.normalize(torch.from_numpy(g["q"]), p=2, dim=1)
    x = torch.nn.functional.normalize(torch.from_numpy(g["g"]), p=2, dim=1)
    sim = S.similarity(q, x).numpy()
    np.testing.assert_allclose(sim, g["sim"], rtol=1e-6, atol=1e-7)
    ranks = S.full_rank(q, x)
    # the reference's argsort has no defined tie order: compare through the scores
    np.testing.assert_array_equal(np.take_along_axis(g["sim"], ranks, 1), np.take_along_axis(g["sim"], g["ranks"], 1))
    sc, ix = S.topk(q, x, 10)
    np.testing.assert_array_equal(ix, g["topk_idx"])
    np.testing.assert_allclose(sc, g["topk_scores"], rtol=1e-6)


def test_topk_chunked_equals_full_and_merge():
    from oracle import synth
    Q, X, _ = synth.retrieval_set(5, 3000, 32, seed=5)
    full = S.full_rank(Q, X)[:, :50]
    sc, ix = S.topk(Q, X, 50, chunk=700)
    np.testing.assert_array_equal(ix, full)
    parts = [S.topk(Q, X[lo:lo + 1000], 50, idx_offset=lo) for lo in range(0, 3000, 1000)]
    ms, mi = S.merge_shards([p[0] for p in parts], [p[1] for p in parts], 50)
    np.testing.assert_array_equal(mi, full)
    # exact ties: duplicate rows must come out in ascending index order
    X[10] = X[2000]
    X[20] = X[2000]
    sc, ix = S.topk(X[2000:2001], X, 3)
    assert list(ix[0]) == [10, 20, 2000]
