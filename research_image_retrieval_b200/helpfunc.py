"""Descriptor extraction driver — drop-in for `extract_vectors` (utils/helpfunc.py:18-48).

Same signature and return value (a CPU fp32 tensor `vecs[N, net.outputdim]`).  Differences are mechanical only:
the per-scale descriptors stay on the GPU, the multi-scale mean + re-normalisation (utils/helpfunc.py:31-44) is one
`rir_scale_mean_l2` launch over all images, and there is ONE device->host copy at the end instead of one per image
per scale.  The backbone inside `net.forward_test` is outside this path and runs as the caller's
torch module.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib


def scale_mean_l2(v: torch.Tensor, keep: torch.Tensor | None = None) -> torch.Tensor:
    """v [N, S, D] fp32 CUDA per-scale descriptors, keep [N, S] uint8/bool -> [N, D] = L2(mean over kept scales)."""
    if not v.is_cuda or v.dtype != torch.float32 or v.dim() != 3:
        raise TypeError("scale_mean_l2 expects a float32 CUDA tensor [N, S, D]")
    v = v.contiguous()
    N, S, D = v.shape
    k = None
    if keep is not None:
        k = keep.to(device=v.device, dtype=torch.uint8).contiguous()
        if tuple(k.shape) != (N, S):
            raise ValueError("keep must be [N, S]")
    out = torch.empty((N, D), dtype=torch.float32, device=v.device)
    with torch.cuda.device(v.device):
        _lib.check(_lib.load().rir_scale_mean_l2(v.data_ptr(), None if k is None else k.data_ptr(), N, S, D,
                                                 out.data_ptr(), _lib.stream_ptr()))
    return out


@torch.no_grad()
def extract_vectors(net, loader, ms=[1], device=torch.device('cuda'), print_freq=100):
    """Per image: (multi-scale) `net.forward_test`, mean over the kept scales, re-L2; returns CPU fp32 [N, outputdim].

    Mirrors utils/helpfunc.py:18-48 including the tiny-image rules: single-scale inputs smaller than 36 px are
    upsampled so the short side is 64 (`:24-26`); in multi-scale mode a rescaled input smaller than 36 px is dropped
    from the mean (`:39-41`)."""
    return extract_vectors_device(net, loader, ms, device, print_freq).cpu()


@torch.no_grad()
def extract_database(net, loader, ms=[1], device=torch.device('cuda'), print_freq=100, dtype="bf16", idx_offset=0):
    """descriptors go from the network straight into the search layout (bf16 / fp8 + row scales) in
    HBM — no host round trip between `extract_vectors` and the similarity scan.  Returns a `search.Database`."""
    from .search import Database
    return Database.from_descriptors(extract_vectors_device(net, loader, ms, device, print_freq), dtype,
                                     idx_offset=idx_offset)


@torch.no_grad()
def extract_vectors_device(net, loader, ms=[1], device=torch.device('cuda'), print_freq=100):
    """extract_vectors without the final device->host copy: fp32 [N, outputdim] on `device`."""
    net.eval()
    n = len(loader)
    D = net.outputdim
    if len(ms) == 1:
        vecs = torch.zeros((n, D), dtype=torch.float32, device=device)
        for i, input in enumerate(loader):
            if input.shape[2] < 36 or input.shape[3] < 36:
                s = max(64 / input.shape[2], 64 / input.shape[3])
                input = F.interpolate(input, scale_factor=s, mode='bilinear', align_corners=False)
            vecs[i, :] = net.forward_test(input.to(device)).reshape(-1).float()
            if (i + 1) % print_freq == 0 or i + 1 == n:
                print('\r>>>> {}/{} done...'.format(i + 1, n), end='')
        print('')
        return vecs
    S = len(ms)
    per_scale = torch.zeros((n, S, D), dtype=torch.float32, device=device)
    keep = torch.zeros((n, S), dtype=torch.uint8)
    for i, input in enumerate(loader):
        for j, s in enumerate(ms):
            if s == 1:
                input_ = input.clone()
            else:
                input_ = F.interpolate(input, scale_factor=s, mode='bilinear', align_corners=False)
            if input_.shape[2] < 36 or input_.shape[3] < 36:
                continue
            per_scale[i, j, :] = net.forward_test(input_.to(device)).reshape(-1).float()
            keep[i, j] = 1
        if (i + 1) % print_freq == 0 or i + 1 == n:
            print('\r>>>> {}/{} done...'.format(i + 1, n), end='')
    print('')
    return scale_mean_l2(per_scale, keep)
