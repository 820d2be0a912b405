"""Generate tests/golden/*.npz by running the UNMODIFIED reference functions on seeded inputs.

TEST INFRASTRUCTURE ONLY.  Run here (the build container, where /root/reference exists):

    python -m oracle.make_golden

Every fixture stores the inputs AND the reference outputs, so the GPU-box tests need neither /root/reference nor
this script.  Fixtures are deliberately tiny (a few hundred KB in total).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_adapter, synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _gnd_to_arrays(gnd, keys):
    out = {}
    for k in keys:
        out[f"{k}_off"] = np.cumsum([0] + [len(g[k]) for g in gnd]).astype(np.int64)
        out[f"{k}_ids"] = (np.concatenate([np.asarray(g[k], dtype=np.int64) for g in gnd])
                           if gnd else np.zeros(0, np.int64))
    return out


def golden_map(ref):
    """compute_map / compute_map_and_print on ROxford-like shapes (scaled down) + every probed edge case."""
    cases = {}
    # -- case A: full ranking, revisited protocol, some queries without easy positives
    nq, n, d = 12, 400, 32
    Q, X, _ = synth.retrieval_set(nq, n, d, seed=11)
    sim = torch.mm(Q, X.t()).numpy()
    ranks = np.argsort(-sim, axis=1, kind="stable").T  # [n, nq]
    gnd = synth.revisited_gnd(nq, n, seed=12)
    (mE, mM, mH), text = ref_adapter.quiet(ref.compute_map_and_print, "roxford5k", "golden", "global", ranks, gnd,
                                           [1, 5, 10], True)
    a = dict(ranks=ranks, mapE=mE, mapM=mM, mapH=mH, text=np.array(text), **_gnd_to_arrays(gnd, ["easy", "hard", "junk"]))
    from oracle.evaluate_oracle import revisited_gnd as rg
    for name, g in zip("EMH", rg(gnd)):
        m, aps, pr, prs = ref.compute_map(ranks, g, [1, 5, 10])
        a.update({f"map_{name}": m, f"aps_{name}": aps, f"pr_{name}": pr, f"prs_{name}": prs})
        m2, aps2 = ref.compute_map(ranks, g)
        a.update({f"map_nokeep_{name}": m2, f"aps_nokeep_{name}": aps2})
    cases["map_full"] = a
    # -- case B: truncated top-k lists [K, nq] and ragged li=True lists
    K = 60
    rt = ranks[:K]
    gm = rg(gnd)[1]
    m, aps = ref.compute_map(rt, gm)
    ragged = [list(ranks[: 20 + 7 * i, i]) for i in range(nq)]
    m3, aps3 = ref.compute_map(ragged, gm, li=True)
    cases["map_truncated"] = dict(ranks=rt, map=m, aps=aps, ragged_len=np.array([len(r) for r in ragged]),
                                  ragged_flat=np.concatenate([np.asarray(r) for r in ragged]), map_li=m3, aps_li=aps3,
                                  **_gnd_to_arrays(gm, ["ok", "junk"]))
    # -- case C: the hand-checked KATs of SURVEY.md §8c, re-derived from the reference itself
    kat = {}
    kat["ap_012_3"] = ref.compute_ap([0, 1, 2], 3)
    kat["ap_13_2"] = ref.compute_ap([1, 3], 2)
    m, aps, pr, prs = ref.compute_map(np.arange(10).reshape(10, 1), [{"ok": [0, 3], "junk": [1]}], [1, 5])
    kat.update(k1_map=m, k1_aps=aps, k1_pr=pr, k1_prs=prs)
    m, aps, pr, prs = ref.compute_map(np.array([[3], [9], [1], [4]]), [{"ok": [1, 2], "junk": [9]}], [1, 5])
    kat.update(k2_map=m, k2_aps=aps, k2_pr=pr, k2_prs=prs)
    m, aps, pr, prs = ref.compute_map(np.array([[0, 0], [1, 1]]), [{"ok": []}, {"ok": [0]}], [1])
    kat.update(k3_map=m, k3_aps=aps, k3_pr=pr, k3_prs=prs)
    m, aps = ref.compute_map(np.array([[0], [1]]), [{"ok": [1]}])
    kat.update(k4_map=m, k4_aps=aps)
    m, aps = ref.compute_map([[5, 6, 7]], [{"ok": [1], "junk": []}], li=True)
    kat.update(k5_map=m, k5_aps=aps)
    cases["map_kat"] = kat
    return cases


def golden_pooling(ref):
    out = {}
    x = synth.feature_maps(3, 24, 7, 5, seed=21) - 0.3  # some negative values: the clamp matters
    x2 = synth.feature_maps(2, 16, 8, 8, seed=22)
    with torch.no_grad():
        out["x"] = x.numpy()
        out["x2"] = x2.numpy()
        out["gem_p3"] = ref.gem()(x).numpy()
        out["gem_p3_x2"] = ref.gem()(x2).numpy()
        out["gem_p2p5"] = ref.gem(p=2.5, eps=1e-6)(x).numpy()
        out["spoc"] = ref.spoc()(x).numpy()
        out["gempooling_p3"] = ref.GeMPooling()(x).numpy()
        out["gempooling_p4p2"] = ref.GeMPooling(p=4.2)(x2).numpy()
        g2 = ref.G2Pooling(p=3.0)
        g2.alpha.data.fill_(1.25)
        g2.beta.data.fill_(-0.05)
        out["g2"] = g2(x).numpy()
        out["spp_max_l1"] = ref.SpatialPyramidPooling(levels=[1], pool_type="max")(x).numpy()  # [B, C, 1]
        # whitening learned by the reference's PCA code and applied as in GeM.forward_test / SOLAR.forward_test
        gen = torch.Generator().manual_seed(23)
        des = (torch.randn(500, 24, generator=gen) * torch.linspace(0.5, 2.0, 24)).numpy().astype(np.float64)
        layer = ref.ConvDimReduction(24, 12)
        layer.initialize_pca_whitening(des)
        out["W"] = layer.weight.data.numpy().reshape(12, 24)
        out["b"] = layer.bias.data.numpy()
        pooled = ref.gem()(x)
        out["gem_tail"] = F.normalize(layer(pooled).squeeze(-1).squeeze(-1), dim=-1).numpy()          # RetrievalNet.py:337-344
        out["solar_tail"] = F.normalize(layer(F.normalize(pooled, p=2.0, dim=1)).squeeze(-1).squeeze(-1), dim=-1).numpy()  # :583-590
        out["des"] = des.astype(np.float32)
    return {"pooling": out}


class _ToyNet(torch.nn.Module):
    """A stand-in with the reference's retrieval interface: outputdim + forward_test -> [1, D] L2-normalised."""

    def __init__(self, ref, D=16):
        super().__init__()
        torch.manual_seed(31)
        self.conv = torch.nn.Conv2d(3, D, 3, stride=2, padding=1)
        self.pool = ref.gem()
        self.outputdim = D

    @torch.no_grad()
    def forward_test(self, x):
        return F.normalize(self.pool(torch.relu(self.conv(x))).squeeze(-1).squeeze(-1), dim=-1)


def golden_extract(ref):
    net = _ToyNet(ref)
    gen = torch.Generator().manual_seed(32)
    images = [torch.rand(1, 3, h, w, generator=gen) for h, w in [(64, 80), (40, 48), (96, 72), (30, 90)]]
    (v1, _) = ref_adapter.quiet(ref.extract_vectors, net, images, [1], torch.device("cpu"))
    (v3, _) = ref_adapter.quiet(ref.extract_vectors, net, images, [1, 2 ** 0.5, 2 ** -0.5], torch.device("cpu"))
    out = {"v_single": v1.numpy(), "v_multi": v3.numpy(), "conv_w": net.conv.weight.data.numpy(),
           "conv_b": net.conv.bias.data.numpy()}
    for i, im in enumerate(images):
        out[f"img{i}"] = im.numpy()
    return {"extract": out}


def golden_ranking(ref):
    """iris_evaluate.py:379-386 executed verbatim (F.normalize, torch.mm, np.argsort) on a tiny set."""
    gen = torch.Generator().manual_seed(41)
    q = torch.randn(6, 64, generator=gen)
    g = torch.randn(300, 64, generator=gen)
    qn = F.normalize(q, p=2, dim=1)
    gn = F.normalize(g, p=2, dim=1)
    sim = torch.mm(qn, gn.t()).cpu().numpy()
    ranks = np.argsort(-sim, axis=1)
    sc, ix = torch.topk(torch.from_numpy(sim), k=10, dim=-1)
    return {"ranking": dict(q=q.numpy(), g=g.numpy(), sim=sim, ranks=ranks, topk_scores=sc.numpy(), topk_idx=ix.numpy())}


def golden_iris_copy(ref):
    """The duplicate copy of compute_map_and_print inside the entry script (iris_evaluate.py:189-265) run on the inputs
    of map_full.npz: its report spells "Easy", unknown names print + return (None, None, None), and the old-protocol
    branch raises ValueError like the utils copy."""
    iris = ref_adapter.load_iris_copy()
    with np.load(os.path.join(OUT, "map_full.npz")) as z:
        g = {k: z[k] for k in z.files}
    nq = len(g["easy_off"]) - 1
    gnd = [{k: g[f"{k}_ids"][g[f"{k}_off"][i]:g[f"{k}_off"][i + 1]] for k in ("easy", "hard", "junk")} for i in range(nq)]
    (vals, text) = ref_adapter.quiet(iris.compute_map_and_print, "roxford5k", "golden", "global", g["ranks"], gnd, [1, 5, 10], True)
    (unk, unk_text) = ref_adapter.quiet(iris.compute_map_and_print, "holidays", "golden", "global", g["ranks"], gnd)
    assert unk == (None, None, None)
    try:
        ref_adapter.quiet(iris.compute_map_and_print, "oxford5k", "golden", "global", g["ranks"],
                          [{"ok": x["easy"], "junk": x["junk"]} for x in gnd])
        old = "no exception"
    except Exception as e:  # noqa: BLE001
        old = type(e).__name__
    return {"iris_copy": dict(mapE=vals[0], mapM=vals[1], mapH=vals[2], text=text, unknown_text=unk_text, old_protocol=old)}


def main():
    ref = ref_adapter.load()
    os.makedirs(OUT, exist_ok=True)
    allc = {}
    only = set(sys.argv[1:])  # e.g. `python oracle/make_golden.py iris_copy` regenerates one fixture
    for fn in (golden_map, golden_pooling, golden_extract, golden_ranking, golden_iris_copy):
        if only and fn.__name__.replace("golden_", "") not in only and not (fn is golden_map and only & {"map_full", "map_truncated", "map_kat"}):
            continue
        allc.update(fn(ref))
    if only:
        allc = {k: v for k, v in allc.items() if k in only}
    for name, arrays in allc.items():
        path = os.path.join(OUT, f"{name}.npz")
        np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
        print(f"wrote {path} ({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    main()
