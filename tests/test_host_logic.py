"""CPU: host-side logic of the drop-in layer (format conversion, sharding arithmetic, exception conventions)."""
import numpy as np
import pytest
import torch

from research_image_retrieval_b200 import _lib, evaluate, search
from oracle import search_oracle as S


def test_shard_bounds_cover_everything():
    for n, w in [(1007323, 8), (10, 3), (7, 8), (4993, 1), (16, 4)]:
        seen = []
        for r in range(w):
            lo, hi = search.shard_bounds(n, w, r)
            assert 0 <= lo <= hi <= n
            seen += list(range(lo, hi)) if n < 100 else []
        assert search.shard_bounds(n, w, 0)[0] == 0 and search.shard_bounds(n, w, w - 1)[1] == n
        if n < 100:
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        search.shard_bounds(10, 2, 2)


def test_pad_dim_and_k_rules():
    assert search.pad_dim(2048, "bf16") == 2048 and search.pad_dim(2050, "bf16") == 2056
    assert search.pad_dim(100, "fp8") == 112 and search.pad_dim(5, "fp32") == 8
    assert search.clamp_k(None, 77) == 77 and search.clamp_k(100, 77) == 77 and search.clamp_k(5, 77) == 5
    with pytest.raises(ValueError):
        search.clamp_k(0, 5)
    search.check_k_supported(100, 10 ** 6)
    search.check_k_supported(16384, 16384)
    with pytest.raises(ValueError):
        search.check_k_supported(9000, 10 ** 6)


def test_ids_to_csr_keeps_duplicates_and_sorts():
    ids, off = evaluate.ids_to_csr([[5, 1, 5], [], np.array([9.0, 2.0]), np.empty(0)])
    assert ids.dtype == np.int32 and off.dtype == np.int32
    assert list(off) == [0, 3, 3, 5, 5]
    assert list(ids) == [1, 5, 5, 2, 9]


def test_ranks_to_rows_layouts():
    r = np.arange(12).reshape(4, 3)  # [L=4, nq=3]
    rows, L = evaluate.ranks_to_rows(r, False, 3)
    assert L == 4 and rows.shape == (3, 4) and rows.dtype == np.int32
    np.testing.assert_array_equal(rows, r.T)
    rows, L = evaluate.ranks_to_rows([[1, 2, 3], [], [7]], True, 3)
    assert L == 3
    np.testing.assert_array_equal(rows, [[1, 2, 3], [-1, -1, -1], [7, -1, -1]])
    rows, L = evaluate.ranks_to_rows(torch.arange(12).reshape(4, 3), False, 3)
    assert rows.dtype == torch.int32 and tuple(rows.shape) == (3, 4)


def test_finish_conventions():
    aps = np.array([0.5, np.inf])
    st_ok = np.array([_lib.RIR_MAP_OK, _lib.RIR_MAP_EMPTY_OK], dtype=np.int32)
    m, a = evaluate._finish(0.5, aps, np.zeros(0), np.zeros((2, 0)), st_ok, None)
    assert isinstance(m, np.float64) and m == 0.5 and np.isinf(a[1])
    out = evaluate._finish(0.5, aps, np.array([1.0]), np.array([[1.0], [np.inf]]), st_ok, [1])
    assert len(out) == 4 and out[3].shape == (2, 1)
    st_nopos = np.array([_lib.RIR_MAP_NO_POS_RETRIEVED, _lib.RIR_MAP_OK], dtype=np.int32)
    with pytest.raises(ValueError, match="max"):
        evaluate._finish(0.1, aps, np.array([1.0]), np.ones((2, 1)), st_nopos, [1])
    evaluate._finish(0.1, aps, np.zeros(0), np.zeros((2, 0)), st_nopos, None)  # no keeps -> fine, ap = 0
    with pytest.raises(ZeroDivisionError):
        evaluate._finish(np.nan, aps, np.zeros(0), np.zeros((2, 0)), np.array([1, 1], dtype=np.int32), None)


def test_protocol_masks():
    # bit 0/1/2 = ok from A/B/C; bit 4/5/6 = junk from A/B/C with A=easy, B=hard, C=junk (utils/evaluate.py:163-185)
    assert evaluate.PROTO_EASY == (1 | (4 | 2) << 4)
    assert evaluate.PROTO_MEDIUM == ((1 | 2) | 4 << 4)
    assert evaluate.PROTO_HARD == (2 | (4 | 1) << 4)
    assert evaluate.PROTO_OK_A_JUNK_C == (1 | 4 << 4)


def test_merge_topk_host_matches_oracle():
    rng = np.random.RandomState(0)
    G, nq, k = 4, 5, 7
    sc = np.sort(rng.rand(G, nq, k).astype(np.float32), axis=2)[:, :, ::-1].copy()
    ix = np.stack([rng.permutation(1000)[: nq * k].reshape(nq, k) + 1000 * g for g in range(G)]).astype(np.int32)
    sc[1, :, -2:] = -np.inf
    ix[1, :, -2:] = -1
    sc[2, 0, 0] = sc[0, 0, 0] = 2.0  # an exact tie across shards at the top -> lower global index first
    ms, mi = search.merge_topk_host(sc, ix, k)
    os_, oi = S.merge_shards([sc[g] for g in range(G)], [ix[g] for g in range(G)], k)
    np.testing.assert_array_equal(mi, oi)
    np.testing.assert_array_equal(ms, os_)
    assert mi[0, 0] == min(ix[0, 0, 0], ix[2, 0, 0])


def test_pad_topk():
    sc, ix = search.pad_topk(torch.ones(2, 3), torch.ones(2, 3, dtype=torch.int32), 5)
    assert tuple(sc.shape) == (2, 5) and torch.isinf(sc[:, 3:]).all() and (ix[:, 3:] == -1).all()


def test_formats_gnd_loader_and_descriptor_store(tmp_path):
    """the gnd pickle loader (dataset/configdataset.py:27-57) and the sharded descriptor store."""
    import pickle

    import torch

    from research_image_retrieval_b200 import formats
    from research_image_retrieval_b200.search import shard_bounds

    # gnd_{dataset}.pkl as revisitop ships it: imlist, qimlist, gnd (list of dicts with bbx / easy / hard / junk)
    d = tmp_path / "roxford5k"
    d.mkdir()
    gnd = [{"bbx": [0, 0, 1, 1], "easy": [3, 1], "hard": [7], "junk": [2, 9]}, {"bbx": [0, 0, 1, 1], "easy": [], "hard": [4]}]
    with open(d / "gnd_roxford5k.pkl", "wb") as f:
        pickle.dump({"imlist": [f"im{i}" for i in range(10)], "qimlist": ["q0", "q1"], "gnd": gnd}, f)
    cfg = formats.RoxfordAndRparis("ROxford5k", str(tmp_path))
    assert cfg["n"] == 10 and cfg["nq"] == 2 and cfg["dataset"] == "roxford5k"
    assert cfg["im_fname"][3].endswith("roxford5k/jpg/im3.jpg") and cfg["qim_fname"][1].endswith("q1.jpg")
    with pytest.raises(ValueError):
        formats.RoxfordAndRparis("holidays", str(tmp_path))
    csr = formats.gnd_to_csr(cfg["gnd"])
    assert csr["easy"][0].tolist() == [1, 3] and csr["easy"][1].tolist() == [0, 2, 2]
    assert csr["junk"][0].tolist() == [2, 9] and csr["junk"][1].tolist() == [0, 2, 2]   # missing 'junk' -> empty

    # descriptor store: 3 uneven shard files, read back across shard boundaries
    gen = torch.Generator().manual_seed(0)
    rows = torch.randn(1000, 16, generator=gen).to(torch.bfloat16)
    st = formats.DescriptorStore.create(str(tmp_path / "r1m"), 16, "bf16")
    for a, b in [(0, 300), (300, 301), (301, 1000)]:
        st.append(rows[a:b])
    st2 = formats.DescriptorStore(str(tmp_path / "r1m"))
    assert st2.n == 1000 and st2.dim == 16 and len(st2.shards) == 3
    for world in (1, 3, 8):
        for rank in range(world):
            lo, hi = shard_bounds(1000, world, rank)
            got, sc = st2.load_rows(lo, hi)
            assert sc is None and torch.equal(got, rows[lo:hi])
    with pytest.raises(ValueError):
        st2.load_rows(0, 1001)
    # fp8 shards carry their per-row scales
    s8 = formats.DescriptorStore.create(str(tmp_path / "r1m_fp8"), 16, "fp8")
    q8 = torch.randint(0, 255, (50, 16), dtype=torch.uint8)
    sc8 = torch.rand(50)
    s8.append(q8[:20], sc8[:20])
    s8.append(q8[20:], sc8[20:])
    got, sc = formats.DescriptorStore(str(tmp_path / "r1m_fp8")).load_rows(10, 45)
    assert torch.equal(got, q8[10:45]) and torch.equal(sc, sc8[10:45])
    with pytest.raises(ValueError):
        s8.append(q8[:5])


def test_iris_evaluate_unknown_dataset_needs_no_gpu(capsys):
    """iris_evaluate.py:263-265 — unknown dataset names print a message and return (None, None, None)."""
    from research_image_retrieval_b200 import iris_evaluate as IE
    out = IE.compute_map_and_print("holidays", "x", "global", np.zeros((3, 2), dtype=np.int64), [{}, {}])
    assert out == (None, None, None)
    assert capsys.readouterr().out == "Unknown dataset: holidays\n"
