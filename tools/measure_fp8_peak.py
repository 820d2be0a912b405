#!/usr/bin/env python
"""Dense fp8 (e4m3 x e4m3 -> bf16, fp32 accumulate) tensor-core peak of this pool's B200, measured the way the driver
measures the bf16 peak in MEASURED_PEAKS.json: a library GEMM (cuBLASLt through torch._scaled_mm) at 8192^3, best of
10 (burst) and back to back for ~3 s (sustained).  Writes profiles/fp8_peak.json; bench.py uses it as the roofline
denominator of fp8 runs instead of assuming 2 x bf16."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    dev = torch.device("cuda", 0)
    n = 8192
    a = (torch.randn(n, n, device=dev) * 0.5).to(torch.float8_e4m3fn)
    b = (torch.randn(n, n, device=dev) * 0.5).to(torch.float8_e4m3fn).t()   # column-major second operand
    one = torch.ones((), device=dev)
    f = lambda: torch._scaled_mm(a, b, scale_a=one, scale_b=one, out_dtype=torch.bfloat16)  # noqa: E731
    for _ in range(5):
        f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    flops = 2.0 * n ** 3
    burst = flops / (best * 1e-3) / 1e12
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 0
    t0 = time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < 3.0:
        for _ in range(50):
            f()
        reps += 50
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    sustained = flops * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
    # bf16 the same way, as a cross-check against MEASURED_PEAKS.json
    x, y = torch.randn(n, n, device=dev, dtype=torch.bfloat16), torch.randn(n, n, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        x @ y
    torch.cuda.synchronize()
    bb = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        x @ y
        e1.record()
        torch.cuda.synchronize()
        bb = min(bb, e0.elapsed_time(e1))
    out = {"fp8_tflops": burst, "fp8_tflops_sustained": sustained, "bf16_tflops_same_run": flops / (bb * 1e-3) / 1e12,
           "gpu_name": torch.cuda.get_device_name(0), "torch": torch.__version__,
           "how": "torch._scaled_mm e4m3 x e4m3 -> bf16 (cuBLASLt), 8192^3, 2*N^3 flops: best of 10 (burst) and back to back "
                  "for 3 s (sustained); bf16 torch.matmul best of 10 in the same process"}
    print(json.dumps(out))
    if "--write" in sys.argv:
        with open(os.path.join(ROOT, "gpurun_out", "fp8_peak.json"), "w") as f_:
            json.dump(out, f_, indent=1)


if __name__ == "__main__":
    main()
