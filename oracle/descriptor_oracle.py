"""CPU oracle for the descriptor build (pool -> L2 -> whiten -> L2 -> multi-scale mean).  TEST INFRASTRUCTURE ONLY.

Plain torch-fp32 restatements (CPU) of
  gem / spoc                /root/reference/src/benchmark/networks/RetrievalNet.py:318-325, 359-365
  GeMPooling / G2Pooling    /root/reference/src/benchmark/models/gem_pooling.py:12-23, models/senet_g2.py:132-153
  SPP level-1 max (MAC)     /root/reference/src/benchmark/models/spoc.py:33-35
  ultron gem_pooling        /root/reference/src/benchmark/models/ultron_modules/ultron.py:193-205
  forward_test tails        /root/reference/src/benchmark/networks/RetrievalNet.py:337-344 (GeM), 583-590 (SOLAR)
  multi-scale aggregate     /root/reference/src/benchmark/utils/helpfunc.py:31-44
  PCA-whitening learn       /root/reference/src/benchmark/networks/backbone.py:42-58, networks/spca.py:215-227
The arithmetic lives in torch (pinned 2.7.1 by the reference, 2.11 here): the oracle calls the same torch ops on the
CPU.  Floating point: the CUDA kernels are compared at rtol 1e-5.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def gem(x: torch.Tensor, p: float = 3.0, eps: float = 1e-6) -> torch.Tensor:
    x = x.float()
    return x.clamp(min=eps).pow(p).mean(dim=(-2, -1), keepdim=True).pow(1.0 / p)


def spoc(x: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    return x.float().clamp(min=eps).mean(dim=(-2, -1), keepdim=True)


def mac(x: torch.Tensor) -> torch.Tensor:
    return x.float().amax(dim=(-2, -1), keepdim=True)


def g2(x, p=3.0, eps=1e-6, alpha=1.0, beta=0.0):
    return alpha * gem(x, p, eps) + beta


def l2n(x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    return x / x.norm(p=2, dim=-1, keepdim=True).clamp_min(eps)


def whiten(x: torch.Tensor, W: torch.Tensor, b: torch.Tensor | None) -> torch.Tensor:
    y = x.reshape(x.shape[0], -1).float() @ W.reshape(W.shape[0], -1).float().t()
    return y if b is None else y + b.float()


def head(x, pooling="gem", p=3.0, eps=1e-6, W=None, b=None, l2_before_whiten=False):
    v = {"gem": lambda: gem(x, p, eps), "mac": lambda: mac(x), "spoc": lambda: spoc(x, eps)}[pooling]()
    v = v.reshape(v.shape[0], -1)
    if W is None:
        return l2n(v)
    if l2_before_whiten:
        v = l2n(v)
    return l2n(whiten(v, W, b))


def scale_mean_l2(v: torch.Tensor, keep: torch.Tensor | None) -> torch.Tensor:
    """v [N,S,D]; keep [N,S] -> L2(sum_kept / #kept) — helpfunc.py:31-44 (sum in scale order, then /=, then normalize)."""
    N, S, D = v.shape
    out = torch.zeros(N, D)
    for n in range(N):
        vec = torch.zeros(D)
        cnt = 0
        for s in range(S):
            if keep is None or bool(keep[n, s]):
                vec += v[n, s]
                cnt += 1
        vec /= cnt if cnt else float("nan")
        out[n] = F.normalize(vec, p=2, dim=0)
    return out


def pack_bf16(v: torch.Tensor) -> torch.Tensor:
    return v.float().to(torch.bfloat16)


def pack_fp8(v: torch.Tensor):
    """per-row scale = amax/448, rows = round-to-nearest-even e4m3 of v/scale (matches rir_pack_descriptors)."""
    v = v.float()
    amax = v.abs().amax(dim=1)
    scale = torch.where(amax > 0, amax / 448.0, torch.ones_like(amax))
    q = (v / scale[:, None]).to(torch.float8_e4m3fn)
    return q, scale


def pca_whiten_learn(X: np.ndarray, s: float = 1.0):
    """networks/backbone.py:42-58 restated with eigh (the covariance is symmetric): returns m [1,D], P^T [D,D]."""
    N = X.shape[0]
    m = X.mean(axis=0, keepdims=True)
    Xc = X - m
    C = Xc.T @ Xc
    C = (C + C.T) / (2 * N)
    w, V = np.linalg.eigh(C)
    order = np.argsort(w)[::-1]
    w, V = w[order], V[:, order]
    P = np.diag(np.power(w, -0.5 * s)) @ V.T
    return m, P.T


def whitening_layer_from_pca(X: np.ndarray, dim: int):
    """ConvDimReduction.initialize_pca_whitening (networks/spca.py:215-227): W = P[:dim], b = -(P m)[:dim]."""
    m, Pt = pca_whiten_learn(X)
    P = Pt.T
    W = torch.tensor(P[:dim, :], dtype=torch.float32)
    b = -(torch.tensor(P, dtype=torch.float32) @ torch.tensor(m.T, dtype=torch.float32)).squeeze()[:dim]
    return W, b
