#!/usr/bin/env python
"""Roofline of the descriptor-build kernels on BASELINE cfg-4 (batch 256 x 2048 x 32 x 32 conv5 maps -> GeM(p=3) -> L2 ->
PCA-whitening 2048x2048 -> L2).  Prints one JSON line per input dtype.  Algorithmic bytes:
B*C*H*W*s_in + B*C*4 for the pool kernel."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import research_image_retrieval_b200 as rir  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = torch.device("cuda", 0)
    B, C, H, W = 256, 2048, 32, 32
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
        if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
    gen = torch.Generator(device=dev).manual_seed(1004)
    Wm = torch.randn(2048, C, generator=gen, device=dev) / 45.0
    b = torch.randn(2048, generator=gen, device=dev) / 10
    lin = torch.nn.Linear(C, 2048).to(dev)
    lin.weight.data, lin.bias.data = Wm, b
    head = rir.DescriptorHead("gem", whiten_layer=lin)
    for dt in (torch.float32, torch.bfloat16):
        x = (torch.randn(B, C, H, W, generator=gen, device=dev).relu_() * 2).to(dt)
        es = x.element_size()
        pool_bytes = B * C * H * W * es + B * C * 4
        ms_pool = timed(lambda: rir.gem_pool(x))
        ms_pool_gen = timed(lambda: rir.gem_pool(x, p=2.5))
        ms_max = timed(lambda: rir.mac_pool(x))
        ms_head = timed(lambda: head(x))
        head_bytes = B * C * H * W * es + B * 2048 * 4 + 2048 * C * 4      # maps + output + W (fp32)
        pooled = rir.gem_pool(x, keepdim=False)
        ms_whiten_tc = timed(lambda: rir.whiten(pooled, lin.weight, lin.bias, l2_after=True))
        ms_whiten_fp32 = timed(lambda: rir.whiten(pooled, lin.weight, lin.bias, l2_after=True, exact_fp32=True))
        # torch eager restatement of the reference ops on the same GPU, for context only
        xf = x
        ms_torch = timed(lambda: torch.nn.functional.normalize(lin(torch.nn.functional.normalize(
            torch.nn.functional.avg_pool2d(xf.float().clamp(min=1e-6).pow(3.0), (H, W)).pow(1 / 3.0).flatten(1), dim=-1)), dim=-1),
            reps=5)
        print(json.dumps({
            "workload": f"cfg-4 GeM build: {B}x{C}x{H}x{W} {str(dt).split('.')[-1]} -> gem(p=3) -> L2 -> whiten {C}->2048 -> L2",
            "pool_p3_ms": ms_pool, "pool_p3_GBps": pool_bytes / ms_pool / 1e6, "pool_p3_frac_of_measured_hbm": pool_bytes / ms_pool / 1e6 / peak,
            "pool_generic_p_ms": ms_pool_gen, "pool_max_ms": ms_max, "head_total_ms": ms_head,
            "head_algorithmic_bytes": head_bytes, "head_GBps": head_bytes / ms_head / 1e6,
            "head_frac_of_measured_hbm": head_bytes / ms_head / 1e6 / peak,
            "whiten_l2_tensor_core_ms": ms_whiten_tc, "whiten_l2_fp32_cuda_core_ms": ms_whiten_fp32,
            "algorithmic_bytes_pool": pool_bytes, "torch_eager_same_gpu_ms": ms_torch}))
        del x


def pca():
    """PCA-whitening learn, dense part: 20,000 x 2048 descriptors (cfg-4's whitening set)."""
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(7)
    X = torch.randn(20000, 2048, generator=gen, device=dev)
    X = X / X.norm(dim=1, keepdim=True)
    ms = timed(lambda: rir.pca_covariance(X), reps=5)
    mean, cov = rir.pca_covariance(X)
    Xd = X.double()
    ref = (Xd - Xd.mean(0)).t() @ (Xd - Xd.mean(0)) / X.shape[0]
    err = float((cov.double() - ref).abs().max() / ref.abs().max())
    print(json.dumps({"workload": "pca covariance 20000 x 2048 fp32 (mean + centred X^T X / N)", "ms": ms,
                      "impl": "fp32 CUDA-core SYRK" if os.environ.get("RIR_PCA_FP32") == "1" else "split-bf16 tcgen05",
                      "max_abs_err_vs_fp64_rel_to_max": err, "symmetric": bool(torch.equal(cov, cov.t()))}))


if __name__ == "__main__":
    if "--pca" in sys.argv:
        pca()
    else:
        main()
