"""CPU: pin the oracle against the LIVE reference functions (only where /root/reference is mounted)."""
import numpy as np
import pytest
import torch

from oracle import descriptor_oracle as D
from oracle import evaluate_oracle as E
from oracle import ref_adapter, synth

pytestmark = pytest.mark.skipif(not ref_adapter.available(), reason="reference tree not mounted (GPU box)")


@pytest.fixture(scope="module")
def ref():
    return ref_adapter.load()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_compute_map_random_cases(ref, seed):
    nq, n = 9, 250
    rng = np.random.RandomState(seed)
    ranks = np.stack([rng.permutation(n) for _ in range(nq)], axis=1)
    gnd = synth.revisited_gnd(nq, n, seed=seed + 100)
    for g in E.revisited_gnd(gnd):
        for keeps in (None, [1, 5, 10], [3]):
            for L in (n, 40):
                r = ranks[:L]
                try:
                    want = ref.compute_map(r, g, keeps)
                except ValueError as e:
                    with pytest.raises(ValueError):
                        E.compute_map(r, g, keeps)
                    continue
                got = E.compute_map(r, g, keeps)
                assert len(got) == len(want)
                for a, b in zip(got, want):
                    np.testing.assert_array_equal(np.asarray(a), np.asarray(b))
    (want, _) = ref_adapter.quiet(ref.compute_map_and_print, "rparis6k", "x", "y", ranks, gnd)
    assert E.compute_map_and_print_values(ranks, gnd) == tuple(float(w) for w in want)


def test_reference_old_protocol_is_broken(ref):
    with pytest.raises(ValueError):
        ref_adapter.quiet(ref.compute_map_and_print, "oxford5k", "x", "y", np.arange(4).reshape(4, 1), [{"ok": [1], "junk": []}])


def test_pooling_against_live_reference(ref):
    x = synth.feature_maps(2, 32, 9, 6, seed=7) - 0.2
    with torch.no_grad():
        np.testing.assert_allclose(D.gem(x).numpy(), ref.gem()(x).numpy(), rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(D.gem(x, 4.5).numpy(), ref.GeMPooling(p=4.5)(x).numpy(), rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(D.spoc(x).numpy(), ref.spoc()(x).numpy(), rtol=1e-5, atol=1e-7)


def test_pca_whitening_learn(ref):
    gen = torch.Generator().manual_seed(3)
    X = (torch.randn(800, 20, generator=gen) * torch.linspace(0.3, 3.0, 20)).numpy().astype(np.float64)
    m_ref, P_ref = ref.pcawhitenlearn_shrinkage(X)
    m, P = D.pca_whiten_learn(X)
    np.testing.assert_allclose(m, m_ref, rtol=1e-12)
    np.testing.assert_allclose(np.abs(P), np.abs(np.real(P_ref)), rtol=1e-6, atol=1e-9)  # eigenvector sign is free
    Y = (X - m) @ P
    np.testing.assert_allclose(np.cov(Y.T, bias=True), np.eye(20), atol=1e-8)


def test_iris_copy_matches_utils_copy_and_golden():
    """a13': the duplicate evaluation functions inside iris_evaluate.py (cut out of the unmodified file) give the same
    numbers as utils/evaluate.py on the golden inputs; the committed iris_copy fixture is what they print."""
    from conftest import csr_to_lists, load_golden
    iris = ref_adapter.load_iris_copy()
    g, gi = load_golden("map_full"), load_golden("iris_copy")
    nq = len(g["easy_off"]) - 1
    lists = {k: csr_to_lists(g[f"{k}_ids"], g[f"{k}_off"]) for k in ("easy", "hard", "junk")}
    gnd = [{k: lists[k][i] for k in lists} for i in range(nq)]
    (vals, text) = ref_adapter.quiet(iris.compute_map_and_print, "roxford5k", "golden", "global", g["ranks"], gnd, [1, 5, 10], True)
    assert text == str(gi["text"]) and "mAP Easy" in text
    assert tuple(float(v) for v in vals) == (float(g["mapE"]), float(g["mapM"]), float(g["mapH"]))
    assert iris.compute_ap(np.array([1, 3]), 2) == 1 / 3 and iris.compute_ap(np.array([0, 1, 2]), 3) == 1.0   # SURVEY §8c KATs
