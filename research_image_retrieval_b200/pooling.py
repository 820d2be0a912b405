"""Descriptor heads: pooling -> L2 -> whitening -> L2, on librir.so kernels.

Drop-in mirrors (same names, constructor arguments and output shapes) of the reference's inference-time heads:

  gem, spoc            networks/RetrievalNet.py:318-325, 359-365
  GeMPooling           models/gem_pooling.py:12-23          (learnable tensor p)
  G2Pooling            models/senet_g2.py:132-153           (alpha * gem + beta)
  mac / MACPooling     models/spoc.py:12-49 level 1 max
  l2n                  F.normalize(x, p=2, dim=-1)          (networks/RetrievalNet.py:343 ...)
  whiten               1x1 conv / Linear W x + b            (networks/RetrievalNet.py:342,588; networks/spca.py:61-64)
  DescriptorHead       GeM.forward_test / SOLAR.forward_test tail (networks/RetrievalNet.py:337-344, 583-590) —
                       one C-ABI call, rir_gem_l2_whiten (pool -> tcgen05 whitening -> bias + L2)

These are inference (`forward_test`) paths: the kernels do not record autograd graphs.  Inputs must be CUDA tensors on
a B200; there is no CPU or eager-PyTorch fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from ._lib import RIR_BF16, RIR_F32, RIR_POOL_AVG, RIR_POOL_GEM, RIR_POOL_MAX


def _pool(x: torch.Tensor, mode: int, p: float = 3.0, eps: float = 1e-6, alpha: float = 1.0, beta: float = 0.0,
          keepdim: bool = True) -> torch.Tensor:
    if x.dim() != 4:
        raise ValueError(f"expected a [B, C, H, W] feature map, got shape {tuple(x.shape)}")
    if x.dtype == torch.float32:
        dt = RIR_F32
    elif x.dtype == torch.bfloat16:
        dt = RIR_BF16
    else:
        raise TypeError(f"feature maps must be float32 or bfloat16, got {x.dtype}")
    if not x.is_cuda:
        raise TypeError("feature maps must be CUDA tensors (no CPU path)")
    if torch.is_grad_enabled() and x.requires_grad:
        # the kernels record no autograd graph: in the reference's TRAINING forward (RetrievalNet.GeM.forward) the
        # backbone would silently stop receiving gradients — refuse instead of returning a detached tensor
        raise RuntimeError("librir pooling is an inference (forward_test) path: call it under torch.no_grad() or on "
                           "detached feature maps; use the reference's torch pooling for training")
    x = x.contiguous()
    B, C, H, W = x.shape
    out = torch.empty((B, C), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().rir_pool(x.data_ptr(), dt, B, C, H * W, mode, float(p), float(eps), float(alpha),
                                        float(beta), out.data_ptr(), _lib.stream_ptr()))
    return out.view(B, C, 1, 1) if keepdim else out


def _no_param_grad(module: nn.Module) -> None:
    """Learnable pooling parameters (p, alpha, beta) get no gradient from these kernels: refuse in training mode."""
    if module.training and torch.is_grad_enabled() and any(t.requires_grad for t in module.parameters()):
        raise RuntimeError(f"{type(module).__name__} (librir) does not train its parameters: switch to .eval() / "
                           "torch.no_grad(), or use the reference module for training")


def gem_pool(x, p=3.0, eps=1e-6, keepdim=True):
    """(mean_hw clamp(x, eps)^p)^(1/p) — F.avg_pool2d(x.clamp(min=eps).pow(p), (H, W)).pow(1/p)."""
    return _pool(x, RIR_POOL_GEM, p=float(p), eps=eps, keepdim=keepdim)


def mac_pool(x, keepdim=True):
    """Global max over H x W (MAC; SpatialPyramidPooling level 1, pool_type='max')."""
    return _pool(x, RIR_POOL_MAX, keepdim=keepdim)


def spoc_pool(x, eps=1e-6, keepdim=True):
    """F.avg_pool2d(x.clamp(min=eps), (H, W))."""
    return _pool(x, RIR_POOL_AVG, eps=eps, keepdim=keepdim)


def l2n(x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """F.normalize(x, p=2, dim=-1) for a [..., D] fp32 CUDA tensor."""
    if not x.is_cuda or x.dtype != torch.float32:
        raise TypeError("l2n expects a float32 CUDA tensor")
    x = x.contiguous()
    d = x.shape[-1]
    n = x.numel() // d if d else 0
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().rir_l2_normalize(x.data_ptr(), n, d, float(eps), out.data_ptr(), _lib.stream_ptr()))
    return out


# ----------------------------------------------------------------------------------------------
# whitening on the tensor cores (split-bf16 tcgen05 contraction, fp32 accurate — csrc/dense_mma.cu)
# ----------------------------------------------------------------------------------------------
class PreparedWhitening:
    """Whitening weights in the layout the tensor-core contraction reads (exact bf16 pairs, rir_whiten_prepare).
    Prepared once per weight tensor and cached; a changed tensor (other storage or in-place update) is re-prepared."""

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor | None):
        w2 = weight.detach().reshape(weight.shape[0], -1).contiguous().float()
        if not w2.is_cuda:
            raise TypeError("whitening weights must be CUDA tensors")
        self.d_out, self.C = w2.shape
        lib = _lib.load()
        self.w12 = torch.empty(lib.rir_whiten_prepared_bytes(self.d_out, self.C), dtype=torch.uint8, device=w2.device)
        with torch.cuda.device(w2.device):
            _lib.check(lib.rir_whiten_prepare(w2.data_ptr(), self.d_out, self.C, self.w12.data_ptr(), _lib.stream_ptr()))
        self.bias = None if bias is None else bias.detach().contiguous().float()
        self._ws = {}

    def workspace(self, B: int) -> torch.Tensor:
        key = (B, torch.cuda.current_stream(self.w12.device).cuda_stream)
        ws = self._ws.get(key)
        if ws is None:
            need = _lib.load().rir_gem_l2_whiten_workspace(B, self.C, self.d_out)
            ws = self._ws[key] = torch.empty(max(need, 256), dtype=torch.uint8, device=self.w12.device)
        return ws


_prepared = {}


def clear_prepared_whitening() -> None:
    """Forget every prepared weight set (needed only after an update torch's version counter cannot see, i.e. one made
    through `.data`)."""
    _prepared.clear()


def prepare_whitening(weight: torch.Tensor, bias: torch.Tensor | None = None) -> PreparedWhitening:
    """Cached per (storage, version counter, shape): optimizer steps, load_state_dict and in-place ops under no_grad
    bump the counter and trigger a re-preparation; writes through `.data` do not — call clear_prepared_whitening()."""
    key = (weight.data_ptr(), weight._version, tuple(weight.shape),
           None if bias is None else (bias.data_ptr(), bias._version))
    pw = _prepared.get(key)
    if pw is None:
        if len(_prepared) > 16:
            _prepared.clear()
        pw = _prepared[key] = PreparedWhitening(weight, bias)
    return pw


def whiten(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None = None, l2_after: bool = False,
           l2_before: bool = False, exact_fp32: bool = False):
    """y = W x + b for pooled descriptors.

    x: [B, C] or [B, C, 1, 1];  weight: [d_out, C] (Linear) or [d_out, C, 1, 1] (1x1 Conv2d);  returns [B, d_out].
    Runs on the tensor cores with fp32 accuracy (rir_whiten_prepared: weights split once into exact bf16 pairs);
    exact_fp32=True takes the plain CUDA-core fp32 kernel (rir_whiten) instead.
    """
    x2 = x.reshape(x.shape[0], -1).contiguous().float()
    if not (x2.is_cuda and weight.is_cuda):
        raise TypeError("whiten expects CUDA tensors")
    if x2.shape[1] != weight.reshape(weight.shape[0], -1).shape[1]:
        raise ValueError(f"whiten: x has {x2.shape[1]} channels, weight expects {weight.reshape(weight.shape[0], -1).shape[1]}")
    out = torch.empty((x2.shape[0], weight.shape[0]), dtype=torch.float32, device=x2.device)
    lib = _lib.load()
    if exact_fp32:
        if l2_before:
            x2 = l2n(x2)
        w2 = weight.reshape(weight.shape[0], -1).contiguous().float()
        b = None if bias is None else bias.contiguous().float()
        with torch.cuda.device(x2.device):
            _lib.check(lib.rir_whiten(x2.data_ptr(), w2.data_ptr(), None if b is None else b.data_ptr(), x2.shape[0],
                                      x2.shape[1], w2.shape[0], 1 if l2_after else 0, out.data_ptr(), _lib.stream_ptr()))
        return out
    pw = prepare_whitening(weight, bias)
    if x2.shape[0] == 0:
        return out
    ws = pw.workspace(x2.shape[0])
    with torch.cuda.device(x2.device):
        _lib.check(lib.rir_whiten_prepared(x2.data_ptr(), pw.w12.data_ptr(), None if pw.bias is None else pw.bias.data_ptr(),
                                           x2.shape[0], pw.C, pw.d_out, 1 if l2_before else 0, 1 if l2_after else 0,
                                           out.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    return out


def gem_l2_whiten(x: torch.Tensor, mode: int = RIR_POOL_GEM, p: float = 3.0, eps: float = 1e-6, weight=None, bias=None,
                  l2_before: bool = False, l2_after: bool = True) -> torch.Tensor:
    """The fused descriptor head (rir_gem_l2_whiten): feature maps [B, C, H, W] -> descriptors [B, d_out] in one call —
    pool -> [L2] -> W v + b -> [L2]; weight=None: pool (+ L2) only."""
    if x.dim() != 4:
        raise ValueError(f"expected a [B, C, H, W] feature map, got shape {tuple(x.shape)}")
    if x.dtype == torch.float32:
        dt = RIR_F32
    elif x.dtype == torch.bfloat16:
        dt = RIR_BF16
    else:
        raise TypeError(f"feature maps must be float32 or bfloat16, got {x.dtype}")
    if not x.is_cuda:
        raise TypeError("feature maps must be CUDA tensors (no CPU path)")
    if torch.is_grad_enabled() and x.requires_grad:
        raise RuntimeError("librir descriptor head is an inference (forward_test) path: call it under torch.no_grad()")
    x = x.contiguous()
    B, C, H, W = x.shape
    lib = _lib.load()
    if weight is None:
        out = torch.empty((B, C), dtype=torch.float32, device=x.device)
        ws = torch.empty(max(lib.rir_gem_l2_whiten_workspace(max(B, 1), C, 0), 256), dtype=torch.uint8, device=x.device)
        w12 = b_ptr = None
        d_out = 0
    else:
        pw = prepare_whitening(weight, bias)
        if pw.C != C:
            raise ValueError(f"whitening layer expects {pw.C} channels, feature maps have {C}")
        out = torch.empty((B, pw.d_out), dtype=torch.float32, device=x.device)
        ws = pw.workspace(max(B, 1))
        w12, b_ptr, d_out = pw.w12.data_ptr(), (None if pw.bias is None else pw.bias.data_ptr()), pw.d_out
    with torch.cuda.device(x.device):
        _lib.check(lib.rir_gem_l2_whiten(x.data_ptr(), dt, B, C, H * W, mode, float(p), float(eps), 1.0, 0.0, w12, b_ptr,
                                         d_out, 1 if l2_before else 0, 1 if l2_after else 0, out.data_ptr(),
                                         ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    return out


class gem(nn.Module):
    """networks/RetrievalNet.py:318-325 — fixed p, returns [B, C, 1, 1]."""

    def __init__(self, p=3.0, eps=1e-6):
        super().__init__()
        self.p = p
        self.eps = eps

    def forward(self, x):
        return gem_pool(x, self.p, self.eps)


class spoc(nn.Module):
    """networks/RetrievalNet.py:359-365."""

    def __init__(self, eps=1e-6):
        super().__init__()
        self.eps = eps

    def forward(self, x):
        return spoc_pool(x, self.eps)


class GeMPooling(nn.Module):
    """models/gem_pooling.py:12-23 — p is an nn.Parameter of shape [1]."""

    def __init__(self, p=3.0, eps=1e-6):
        super().__init__()
        self.p = nn.Parameter(torch.ones(1) * p)
        self.eps = eps

    def forward(self, x):
        _no_param_grad(self)
        return gem_pool(x, float(self.p.detach().reshape(-1)[0]), self.eps)


class G2Pooling(nn.Module):
    """models/senet_g2.py:132-153 — alpha * GeM(x) + beta."""

    def __init__(self, p=3.0, eps=1e-6):
        super().__init__()
        self.p = nn.Parameter(torch.ones(1) * p)
        self.eps = eps
        self.alpha = nn.Parameter(torch.ones(1))
        self.beta = nn.Parameter(torch.zeros(1))

    def forward(self, x):
        _no_param_grad(self)
        return _pool(x, RIR_POOL_GEM, p=float(self.p.detach()[0]), eps=self.eps, alpha=float(self.alpha.detach()[0]),
                     beta=float(self.beta.detach()[0]))


class MACPooling(nn.Module):
    """Global max pooling — SpatialPyramidPooling(levels=[1], pool_type='max') (models/spoc.py:12-49)."""

    def forward(self, x):
        return mac_pool(x)


def ultron_gem_pooling(x, gamma):
    """AttentionBasedGlobalPooling.gem_pooling (models/ultron_modules/ultron.py:193-205): p clamped to [1e-7, 100],
    eps 1e-7, returns [B, C]."""
    g = float(torch.as_tensor(gamma).detach().reshape(-1)[0])
    g = min(max(g, 1e-7), 100.0)
    return _pool(x, RIR_POOL_GEM, p=g, eps=1e-7, keepdim=False)


class DescriptorHead(nn.Module):
    """pool -> [L2] -> whiten (+bias) -> L2, the tail of GeM.forward_test (networks/RetrievalNet.py:337-344,
    l2_before_whiten=False) and SOLAR.forward_test (networks/RetrievalNet.py:583-590, l2_before_whiten=True).

    `whiten_layer` may be the reference's nn.Conv2d(1x1) / nn.Linear / ConvDimReduction, or None (pool + L2 only,
    e.g. models/gem_pooling.py:86-92).
    """

    def __init__(self, pooling: str = "gem", p: float = 3.0, eps: float = 1e-6, whiten_layer: nn.Module | None = None,
                 l2_before_whiten: bool = False):
        super().__init__()
        if pooling not in ("gem", "mac", "spoc"):
            raise ValueError(f"unknown pooling {pooling!r}")
        self.pooling = pooling
        self.p = p
        self.eps = eps
        self.whiten_layer = whiten_layer
        self.l2_before_whiten = l2_before_whiten

    @torch.no_grad()
    def forward(self, feature_map: torch.Tensor) -> torch.Tensor:
        mode = {"gem": RIR_POOL_GEM, "mac": RIR_POOL_MAX, "spoc": RIR_POOL_AVG}[self.pooling]
        if self.whiten_layer is None:
            return gem_l2_whiten(feature_map, mode, self.p, self.eps)
        return gem_l2_whiten(feature_map, mode, self.p, self.eps, self.whiten_layer.weight, self.whiten_layer.bias,
                             l2_before=self.l2_before_whiten, l2_after=True)
