"""GPU parity: descriptor build kernels (through the C ABI) vs golden fixtures of the reference and the torch-fp32 oracle.
Floating point tolerance for this stage: rtol 1e-5."""
import numpy as np
import pytest
import torch

import research_image_retrieval_b200 as rir
from conftest import load_golden
from oracle import descriptor_oracle as D
from oracle import synth

pytestmark = pytest.mark.gpu
RT = dict(rtol=1e-5, atol=1e-7)


def test_pooling_golden(cuda_device):
    g = load_golden("pooling")
    x = torch.from_numpy(g["x"]).to(cuda_device)
    x2 = torch.from_numpy(g["x2"]).to(cuda_device)
    np.testing.assert_allclose(rir.gem()(x).cpu().numpy(), g["gem_p3"], **RT)
    np.testing.assert_allclose(rir.gem()(x2).cpu().numpy(), g["gem_p3_x2"], **RT)
    np.testing.assert_allclose(rir.gem(p=2.5)(x).cpu().numpy(), g["gem_p2p5"], **RT)
    np.testing.assert_allclose(rir.spoc()(x).cpu().numpy(), g["spoc"], **RT)
    # learnable-p modules are inference paths: eval() (as the reference's evaluation does) — see the refusal test below
    np.testing.assert_allclose(rir.GeMPooling().to(cuda_device).eval()(x).cpu().numpy(), g["gempooling_p3"], **RT)
    np.testing.assert_allclose(rir.GeMPooling(p=4.2).to(cuda_device).eval()(x2).cpu().numpy(), g["gempooling_p4p2"], **RT)
    g2 = rir.G2Pooling(p=3.0).eval()
    g2.alpha.data.fill_(1.25)
    g2.beta.data.fill_(-0.05)
    np.testing.assert_allclose(g2(x).cpu().numpy(), g["g2"], **RT)
    with torch.no_grad():  # training-mode module under no_grad is fine too
        np.testing.assert_allclose(rir.GeMPooling().to(cuda_device)(x).cpu().numpy(), g["gempooling_p3"], **RT)
    np.testing.assert_allclose(rir.MACPooling()(x).cpu().numpy().reshape(3, 24, 1), g["spp_max_l1"], **RT)
    assert tuple(rir.gem()(x).shape) == (3, 24, 1, 1)


def test_pooling_refuses_silent_no_grad(cuda_device):
    """The kernels record no autograd graph: plugged into a TRAINING forward they must raise, not detach silently."""
    x = torch.rand(2, 8, 4, 4, device=cuda_device)
    with pytest.raises(RuntimeError, match="inference"):
        rir.gem()(x.clone().requires_grad_(True))
    with pytest.raises(RuntimeError, match="does not train"):
        rir.GeMPooling().to(cuda_device)(x)          # training mode, learnable p, grad enabled
    with pytest.raises(RuntimeError, match="does not train"):
        rir.G2Pooling().to(cuda_device)(x)
    assert tuple(rir.gem()(x).shape) == (2, 8, 1, 1)   # plain inference input is fine


def test_heads_golden(cuda_device):
    g = load_golden("pooling")
    x = torch.from_numpy(g["x"]).to(cuda_device)
    layer = torch.nn.Conv2d(24, 12, 1, bias=True)
    layer.weight.data = torch.from_numpy(g["W"]).reshape(12, 24, 1, 1)
    layer.bias.data = torch.from_numpy(g["b"])
    layer = layer.to(cuda_device)
    gem_tail = rir.DescriptorHead("gem", whiten_layer=layer)(x)
    solar_tail = rir.DescriptorHead("gem", whiten_layer=layer, l2_before_whiten=True)(x)
    np.testing.assert_allclose(gem_tail.cpu().numpy(), g["gem_tail"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(solar_tail.cpu().numpy(), g["solar_tail"], rtol=1e-4, atol=1e-6)
    lin = torch.nn.Linear(24, 12).to(cuda_device)
    lin.weight.data = layer.weight.data.reshape(12, 24)
    lin.bias.data = layer.bias.data
    np.testing.assert_allclose(rir.DescriptorHead("gem", whiten_layer=lin)(x).cpu().numpy(), g["gem_tail"], rtol=1e-4,
                               atol=1e-6)


@pytest.mark.parametrize("shape", [(4, 64, 7, 7), (3, 40, 32, 32), (2, 33, 5, 3), (1, 8, 1, 1), (5, 16, 20, 12)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pooling_vs_oracle(cuda_device, shape, dtype):
    x = (synth.feature_maps(*shape, seed=sum(shape)) - 0.1).to(dtype)
    xd = x.to(cuda_device)
    xf = x.float()
    np.testing.assert_allclose(rir.gem_pool(xd).cpu().numpy(), D.gem(xf).numpy(), **RT)
    np.testing.assert_allclose(rir.gem_pool(xd, p=1.0).cpu().numpy(), D.gem(xf, 1.0).numpy(), **RT)
    np.testing.assert_allclose(rir.gem_pool(xd, p=2.0).cpu().numpy(), D.gem(xf, 2.0).numpy(), **RT)
    np.testing.assert_allclose(rir.gem_pool(xd, p=3.7, eps=1e-4).cpu().numpy(), D.gem(xf, 3.7, 1e-4).numpy(), **RT)
    np.testing.assert_allclose(rir.mac_pool(xd).cpu().numpy(), D.mac(xf).numpy(), rtol=0, atol=0)
    np.testing.assert_allclose(rir.spoc_pool(xd).cpu().numpy(), D.spoc(xf).numpy(), **RT)
    # p is clamped to 100 (ultron.py:198); inputs scaled < 1 so x^100 stays finite
    np.testing.assert_allclose(rir.ultron_gem_pooling(xd * 0.1, torch.tensor([250.0])).cpu().numpy(),
                               D.gem((xd * 0.1).float().cpu(), 100.0, 1e-7).numpy().reshape(shape[0], shape[1]),
                               rtol=1e-4, atol=1e-12)


def test_empty_batch_and_errors(cuda_device):
    out = rir.gem_pool(torch.zeros(0, 8, 4, 4, device=cuda_device))
    assert tuple(out.shape) == (0, 8, 1, 1)
    with pytest.raises(ValueError):
        rir.gem_pool(torch.zeros(8, 4, 4, device=cuda_device))
    with pytest.raises(TypeError):
        rir.gem_pool(torch.zeros(1, 8, 4, 4, device=cuda_device, dtype=torch.float16))
    with pytest.raises(rir.RirError):
        rir.gem_pool(torch.ones(1, 8, 4, 4, device=cuda_device), p=-1.0)


@pytest.mark.parametrize("n,d", [(1, 1), (7, 33), (64, 2048), (300, 512), (5, 4100)])
def test_l2n(cuda_device, n, d):
    gen = torch.Generator().manual_seed(n * d)
    x = torch.randn(n, d, generator=gen) * 3
    x[0] = 0  # zero row: divided by eps, stays zero
    np.testing.assert_allclose(rir.l2n(x.to(cuda_device)).cpu().numpy(), D.l2n(x).numpy(), rtol=2e-6, atol=1e-9)


@pytest.mark.parametrize("B,C,dout", [(1, 24, 12), (37, 100, 70), (256, 512, 128), (5, 2048, 2048)])
def test_whiten(cuda_device, B, C, dout):
    gen = torch.Generator().manual_seed(B + C)
    x = torch.randn(B, C, generator=gen)
    W = torch.randn(dout, C, generator=gen) / C ** 0.5
    b = torch.randn(dout, generator=gen)
    want = D.whiten(x, W, b).numpy()
    got = rir.whiten(x.to(cuda_device), W.to(cuda_device), b.to(cuda_device)).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=2e-5)
    got = rir.whiten(x.to(cuda_device).view(B, C, 1, 1), W.to(cuda_device).view(dout, C, 1, 1), None, l2_after=True)
    np.testing.assert_allclose(got.cpu().numpy(), D.l2n(D.whiten(x, W, None)).numpy(), rtol=1e-4, atol=2e-6)
    # the plain fp32 CUDA-core kernel (rir_whiten) stays available and agrees
    got = rir.whiten(x.to(cuda_device), W.to(cuda_device), b.to(cuda_device), exact_fp32=True).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("B,C,hw,dout,dtype", [(64, 2048, 8, 2048, torch.float32), (130, 2048, 4, 512, torch.float32),
                                               (9, 100, 5, 70, torch.float32), (33, 1024, 7, 300, torch.bfloat16),
                                               (256, 512, 4, 2048, torch.float32)])
@pytest.mark.parametrize("l2_before", [False, True])
def test_fused_head_tensor_core(cuda_device, B, C, hw, dout, dtype, l2_before):
    """rir_gem_l2_whiten (pool -> split-bf16 tcgen05 whitening -> bias + L2) vs the fp32 torch oracle: the split keeps
    fp32 accuracy (rtol 1e-4 like the reference-golden head tests; measured error ~1e-6)."""
    fm = synth.feature_maps(B, C, hw, hw, seed=B + C)
    if dtype == torch.bfloat16:
        fm = fm.to(torch.bfloat16).float()
    gen = torch.Generator().manual_seed(dout)
    W = torch.randn(dout, C, generator=gen) / C ** 0.5
    b = torch.randn(dout, generator=gen) / 10
    layer = torch.nn.Linear(C, dout)
    layer.weight.data, layer.bias.data = W, b
    head = rir.DescriptorHead("gem", whiten_layer=layer.to(cuda_device), l2_before_whiten=l2_before)
    got = head(fm.to(cuda_device).to(dtype)).cpu().numpy()
    want = D.head(fm, "gem", W=W, b=b, l2_before_whiten=l2_before).numpy()
    # unit-norm rows: absolute error stays below 3e-6 of the norm (measured ~1e-6) — near-zero components need the atol
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=3e-6)
    assert abs(float(np.linalg.norm(got[0])) - 1.0) < 1e-5
    # second call re-uses the prepared weights; an in-place weight update (one that autograd's version counter sees:
    # optimizer steps, load_state_dict, no_grad in-place ops) must be picked up
    with torch.no_grad():
        layer.weight.mul_(-1.0)
    got2 = head(fm.to(cuda_device).to(dtype)).cpu().numpy()
    want2 = D.head(fm, "gem", W=-W, b=b, l2_before_whiten=l2_before).numpy()
    np.testing.assert_allclose(got2, want2, rtol=1e-4, atol=3e-6)


def test_fused_head_without_whitening(cuda_device):
    fm = synth.feature_maps(5, 96, 6, 6, seed=12)
    got = rir.DescriptorHead("gem")(fm.to(cuda_device)).cpu().numpy()
    np.testing.assert_allclose(got, D.head(fm, "gem").numpy(), rtol=1e-5, atol=1e-7)
    got = rir.DescriptorHead("mac")(fm.to(cuda_device)).cpu().numpy()
    np.testing.assert_allclose(got, D.head(fm, "mac").numpy(), rtol=1e-5, atol=1e-7)


def test_pca_covariance_tensor_core_sizes(cuda_device):
    """Split-bf16 tcgen05 covariance vs fp64 numpy, shapes that are not tile multiples; exactly symmetric."""
    for N, Dm in [(3001, 300), (1000, 2048), (70, 40)]:
        X = torch.randn(N, Dm, generator=torch.Generator().manual_seed(N + Dm)).numpy()
        X = (X / np.linalg.norm(X, axis=1, keepdims=True)).astype(np.float32) + 0.01
        mean, cov = rir.pca_covariance(torch.from_numpy(X).to(cuda_device))
        Xd = X.astype(np.float64)
        cref = (Xd - Xd.mean(0)).T @ (Xd - Xd.mean(0)) / N
        scale = np.abs(cref).max()
        assert np.abs(cov.cpu().numpy() - cref).max() <= 2e-4 * scale
        assert torch.equal(cov, cov.t())


def test_scale_mean_l2(cuda_device):
    gen = torch.Generator().manual_seed(9)
    v = torch.randn(11, 3, 40, generator=gen)
    keep = torch.ones(11, 3, dtype=torch.uint8)
    keep[2, 0] = 0
    keep[5, 1:] = 0
    got = rir.scale_mean_l2(v.to(cuda_device), keep).cpu().numpy()
    np.testing.assert_allclose(got, D.scale_mean_l2(v, keep).numpy(), rtol=1e-5, atol=1e-7)
    got = rir.scale_mean_l2(v.to(cuda_device)).cpu().numpy()
    np.testing.assert_allclose(got, D.scale_mean_l2(v, None).numpy(), rtol=1e-5, atol=1e-7)


class _ToyNet(torch.nn.Module):
    """Same stand-in as oracle/make_golden.py, with the pooling running on the librir kernel."""

    def __init__(self, w, b):
        super().__init__()
        self.conv = torch.nn.Conv2d(3, w.shape[0], 3, stride=2, padding=1)
        self.conv.weight.data = torch.from_numpy(w)
        self.conv.bias.data = torch.from_numpy(b)
        self.pool = rir.gem()
        self.outputdim = w.shape[0]

    @torch.no_grad()
    def forward_test(self, x):
        return rir.l2n(self.pool(torch.relu(self.conv(x))).squeeze(-1).squeeze(-1))


def test_extract_vectors_golden(cuda_device, capsys):
    g = load_golden("extract")
    net = _ToyNet(g["conv_w"], g["conv_b"]).to(cuda_device)
    images = [torch.from_numpy(g[f"img{i}"]) for i in range(4)]
    v1 = rir.extract_vectors(net, images, [1], cuda_device)
    v3 = rir.extract_vectors(net, images, [1, 2 ** 0.5, 2 ** -0.5], cuda_device)
    assert v1.device.type == "cpu" and v1.dtype == torch.float32 and tuple(v1.shape) == (4, 16)
    # the conv runs in cuDNN here and MKL in the fixture: tolerance covers the backbone, not our kernels
    np.testing.assert_allclose(v1.numpy(), g["v_single"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(v3.numpy(), g["v_multi"], rtol=2e-4, atol=2e-5)
    assert ">>>> 4/4 done..." in capsys.readouterr().out
    # the same descriptors, kept on the device and written straight into the search layout
    db = rir.extract_database(net, images, [1, 2 ** 0.5, 2 ** -0.5], cuda_device, dtype="bf16", idx_offset=7)
    assert db.rows.is_cuda and db.rows.dtype == torch.bfloat16 and db.n == 4 and db.idx_offset == 7
    assert torch.equal(db.rows.cpu(), D.pack_bf16(v3))


def test_pack_descriptors_bit_exact(cuda_device):
    gen = torch.Generator().manual_seed(4)
    v = torch.randn(50, 72, generator=gen)
    v[3] = 0
    rows, scale = rir.pack_descriptors(v.to(cuda_device), "bf16")
    assert scale is None and rows.dtype == torch.bfloat16
    assert torch.equal(rows.cpu(), D.pack_bf16(v))
    rows, scale = rir.pack_descriptors(v.to(cuda_device), "fp8")
    q, s = D.pack_fp8(torch.nn.functional.pad(v, (0, 8)))
    assert tuple(rows.shape) == (50, 80)
    assert torch.equal(scale.cpu(), s)
    assert torch.equal(rows.cpu(), q.view(torch.uint8))


def test_pca_whitening_learn(cuda_device):
    """mean + covariance on the GPU (rir_pca_covariance), eigh on the host side; vs the oracle
    restatement of pcawhitenlearn_shrinkage (networks/backbone.py:42-58), which is pinned to the reference."""
    gen = torch.Generator().manual_seed(11)
    N, Dm = 3000, 200                                   # D not a multiple of the 128-wide tile
    scales = torch.logspace(0.5, -0.5, Dm)              # well separated eigenvalues
    X = (torch.randn(N, Dm, generator=gen) * scales + torch.linspace(-1, 1, Dm)).numpy().astype(np.float32)
    mean, cov = rir.pca_covariance(torch.from_numpy(X).to(cuda_device))
    Xd = X.astype(np.float64)
    mref = Xd.mean(0)
    cref = (Xd - mref).T @ (Xd - mref) / N
    np.testing.assert_allclose(mean.cpu().numpy(), mref, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(cov.cpu().numpy(), cref, rtol=2e-4, atol=2e-5)
    assert torch.equal(cov, cov.t())                    # exactly symmetric, like (Xcov + Xcov.T) / 2
    m, Pt = rir.pcawhitenlearn_shrinkage(X)
    m_ref, Pt_ref = D.pca_whiten_learn(Xd)
    assert m.shape == (1, Dm) and Pt.shape == (Dm, Dm) and m.dtype == np.float32
    np.testing.assert_allclose(m, m_ref, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(np.abs(Pt), np.abs(Pt_ref), rtol=5e-3, atol=5e-3)   # eigenvector sign is free
    Y = (Xd - m_ref) @ Pt.astype(np.float64)
    np.testing.assert_allclose(np.cov(Y.T, bias=True), np.eye(Dm), atol=5e-3)      # whitened
    # the layer the descriptor head consumes (networks/spca.py:215-227), dim reduction 200 -> 64
    layer = rir.ConvDimReduction(Dm, 64).to(cuda_device)
    layer.initialize_pca_whitening(X)
    W_ref, b_ref = D.whitening_layer_from_pca(Xd, 64)
    sign = torch.sign((layer.weight.data.reshape(64, Dm).cpu() * W_ref).sum(1))
    np.testing.assert_allclose((layer.weight.data.reshape(64, Dm).cpu() * sign[:, None]).numpy(), W_ref.numpy(), rtol=5e-3,
                               atol=5e-3)
    np.testing.assert_allclose((layer.bias.data.cpu() * sign).numpy(), b_ref.numpy(), rtol=5e-3, atol=5e-3)
    # and applied: whitened + L2 descriptors agree with the oracle's up to the per-component sign
    x = torch.from_numpy(X[:32]).to(cuda_device)
    got = rir.whiten(x, layer.weight, layer.bias, l2_after=True).cpu() * sign
    want = D.l2n(D.whiten(torch.from_numpy(X[:32]), W_ref, b_ref))
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=5e-3, atol=5e-4)
