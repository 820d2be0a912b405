"""PCA-whitening learn — drop-in for `pcawhitenlearn_shrinkage` (networks/backbone.py:42-58) and
`ConvDimReduction` (networks/spca.py:205-227).

The dense part (column mean, centred covariance of the N x D descriptors) runs in `rir_pca_covariance`
(csrc/pca_whiten.cu: the covariance is a split-bf16 tcgen05 contraction, csrc/dense_mma.cu); the D x D symmetric eigen-decomposition is an offline, once-per-model step and uses
`torch.linalg.eigh` in fp64 on the GPU (the reference calls `np.linalg.eig` on the same symmetric matrix).
Eigenvectors are defined up to sign, exactly as in the reference; cosine similarities of whitened descriptors do not
depend on it.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _lib


def pca_covariance(X: torch.Tensor):
    """X [N, D] fp32 CUDA -> (mean [D], cov [D, D]) fp32 on the GPU; cov = (X-m)^T (X-m) / N."""
    if not X.is_cuda:
        raise TypeError("descriptors must be on the GPU (no CPU path)")
    X = X.float().contiguous()
    N, D = X.shape
    lib = _lib.load()
    mean = torch.empty(D, dtype=torch.float32, device=X.device)
    cov = torch.empty((D, D), dtype=torch.float32, device=X.device)
    ws = torch.empty(max(lib.rir_pca_covariance_workspace(N, D), 256), dtype=torch.uint8, device=X.device)
    with torch.cuda.device(X.device):
        _lib.check(lib.rir_pca_covariance(X.data_ptr(), N, D, mean.data_ptr(), cov.data_ptr(), ws.data_ptr(), ws.numel(),
                                          _lib.stream_ptr()))
    return mean, cov


def pcawhitenlearn_shrinkage(X, s: float = 1.0, device=None):
    """Learn PCA whitening with shrinkage from descriptors X [N, D] (numpy array or tensor).

    Returns `(m, P.T)` like the reference: m [1, D] and P^T [D, D] numpy arrays with P = diag(eigval^(-s/2)) V^T,
    eigenvalues sorted in descending order (networks/backbone.py:51-58)."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    Xt = torch.as_tensor(X).to(device=device, dtype=torch.float32)
    mean, cov = pca_covariance(Xt)
    w, V = torch.linalg.eigh(cov.double())           # ascending
    w, V = torch.flip(w, dims=[0]), torch.flip(V, dims=[1])
    P = w.pow(-0.5 * s)[:, None] * V.t()             # diag(w^-s/2) V^T as a row scaling (no D x D GEMM)
    out_dtype = X.dtype if isinstance(X, np.ndarray) else np.float32
    return mean.reshape(1, -1).cpu().numpy().astype(out_dtype), P.t().contiguous().cpu().numpy().astype(out_dtype)


class ConvDimReduction(nn.Conv2d):
    """Dimensionality reduction as a 1x1 convolution, initialised from PCA whitening (networks/spca.py:205-227).

    `forward` stays torch's Conv2d (training path); at inference use `rir.whiten(x, layer.weight, layer.bias)` or
    `rir.DescriptorHead(..., whiten_layer=layer)`."""

    def __init__(self, input_dim, dim):
        super().__init__(input_dim, dim, (1, 1), padding=0, bias=True)

    def initialize_pca_whitening(self, des):
        """des [N, input_dim] descriptors.  Fills the layer with the first `dim` whitening directions,
        weight = P[:dim] and bias = -(P m)[:dim], freezes both, and returns `(m, P.T)` exactly as
        pcawhitenlearn_shrinkage does (the reference's method returns the same pair, networks/spca.py:215-227)."""
        dev = self.weight.device if self.weight.is_cuda else None
        m, Pt = pcawhitenlearn_shrinkage(des, device=dev)
        dim = self.out_channels
        rows = torch.from_numpy(np.ascontiguousarray(Pt.T[:dim])).float()          # [dim, input_dim]
        shift = -(rows @ torch.from_numpy(np.ascontiguousarray(m.reshape(-1))).float())
        with torch.no_grad():
            self.weight.copy_(rows.reshape(dim, -1, 1, 1).to(self.weight.device))
            self.bias.copy_(shift.to(self.bias.device))
        self.weight.requires_grad_(False)
        self.bias.requires_grad_(False)
        return m, Pt
