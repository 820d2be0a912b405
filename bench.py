#!/usr/bin/env python
"""bench.py — BASELINE.json's headline metric on synthetic data of the named shape.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, librir.so)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (torch.mm + np.argsort)

Workload (BASELINE.json configs[1], "RParis6k-shape with 1M distractors"): 70 queries x 1,007,323 database rows x
2048-d bf16, exact top-100; the database is row-sharded over the N GPUs (strong scaling: total work is fixed).
One step = one 70-query batch through the whole search path:
    sample pass -> per-query threshold -> full scan (tcgen05, TMA) -> exact select [-> NCCL all-gather -> merge].
`value` = queries/s with the query batch already packed in HBM; `e2e` = the same through the host-facing call:
pinned fp32 host queries -> H2D -> bf16 pack -> search -> D2H of (scores, idx), database resident.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "queries/sec exact top-100 on 1M x 2048 bf16 db"
UNIT = "queries/s"
N_DB, DIM, NQ, TOPK = 1_007_323, 2048, 70, 100   # 6,322 RParis images + 1,001,001 distractors
SEED = 1002
CHUNK = 65536


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # development overrides (the driver never passes these; the default is the BASELINE workload)
    ap.add_argument("--nq", type=int, default=NQ)
    ap.add_argument("--n", "--n-db", dest="n", type=int, default=N_DB)  # (--n-db: unambiguous under torchrun)
    ap.add_argument("--d", type=int, default=DIM)
    ap.add_argument("--k", type=int, default=TOPK)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp8"])
    ap.add_argument("--path", default="auto", choices=["auto", "stream", "mma"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"])
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# clocks sampler (NVML, background thread)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=1.0)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------------------------
# synthetic data (identical rows whatever N is: one generator seed per 65,536-row chunk)
# ----------------------------------------------------------------------------------------------
def make_rows_fp32(lo: int, hi: int, d: int, device):
    import torch
    out = torch.empty((hi - lo, d), dtype=torch.float32, device=device)
    c0 = lo // CHUNK
    pos = lo
    while pos < hi:
        c = pos // CHUNK
        gen = torch.Generator(device=device).manual_seed(SEED * 100003 + c)
        blk = torch.randn(CHUNK, d, generator=gen, device=device)
        a, b = pos - c * CHUNK, min(hi, (c + 1) * CHUNK) - c * CHUNK
        seg = blk[a:b]
        out[pos - lo: pos - lo + (b - a)] = seg / seg.norm(dim=1, keepdim=True)
        pos += b - a
    del c0
    return out


def make_queries_fp32(nq: int, d: int):
    import torch
    gen = torch.Generator().manual_seed(SEED)
    q = torch.randn(nq, d, generator=gen)
    return q / q.norm(dim=1, keepdim=True)


# ----------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU path, restated in oracle/ (the reference is pure Python: no _ref build)
# ----------------------------------------------------------------------------------------------
def cpu_reference_step_time(nq: int, n_sample: int, d: int, reps: int, warm: int):
    """iris_evaluate.py:383-386 as written — fp32 torch.mm + np.argsort(-sim, axis=1) — on a bounded row sample."""
    import numpy as np
    import torch
    from oracle import search_oracle as S
    torch.set_num_threads(os.cpu_count() or 1)
    gen = torch.Generator().manual_seed(SEED)
    X = torch.randn(n_sample, d, generator=gen)
    X /= X.norm(dim=1, keepdim=True)
    Q = make_queries_fp32(nq, d)
    times, times_topk = [], []
    for i in range(warm + reps):
        t0 = time.perf_counter()
        sim = S.similarity(Q, X).numpy()
        t1 = time.perf_counter()
        ranks = np.argsort(-sim, axis=1)
        t2 = time.perf_counter()
        sc, ix = torch.topk(torch.from_numpy(sim), k=min(TOPK, n_sample), dim=-1)
        t3 = time.perf_counter()
        if i >= warm:
            times.append(t2 - t0)                    # mm + full argsort: what the reference does
            times_topk.append((t1 - t0) + (t3 - t2))  # mm + torch.topk: the "strong CPU" variant
        del ranks, sc, ix
    return times, times_topk


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = min(args.n, 131072)
    scale = args.n / n_sample
    times, _ = cpu_reference_step_time(args.nq, n_sample, args.d, args.steps, args.warmup)
    total = sum(times)
    per_step_full = total / len(times) * scale
    value = args.nq / per_step_full
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step_full * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, extra={"cpu_sample_rows": n_sample}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"fp32 torch.mm + np.argsort(-sim,1) (iris_evaluate.py:383-386) on {args.nq} queries x "
                                   f"{n_sample} rows per step, time scaled x{scale:.3f} (linear in rows) to {args.n} rows"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def metric_name(args):
    if (args.n, args.d, args.k, args.dtype) == (N_DB, DIM, TOPK, "bf16"):
        return METRIC
    return (f"queries/sec exact top-{args.k} on {args.n} x {args.d} {args.dtype} db "
            "(development override of the BASELINE workload)")


def workload_config(args, extra=None):
    c = {"workload": f"BASELINE configs[1]: {args.nq} queries x {args.n} db rows x {args.d}-d {args.dtype}, exact top-{args.k}, "
                     f"db row-sharded over {args.gpus} GPU(s)",
         "nq": args.nq, "n_db": args.n, "dim": args.d, "k": args.k, "parallelism": f"db-shard x{args.gpus}",
         "l2_note": "per-GPU shard (>= 516 MB) exceeds the 126 MB L2, so every step re-reads it from HBM; no explicit flush"}
    if extra:
        c.update(extra)
    return c


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import research_image_retrieval_b200 as rir
    from research_image_retrieval_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = rir.load()
    _lib.check(lib.rir_device_check())

    # ---- resident database shard ----
    lo, hi = rir.shard_bounds(args.n, world, rank)
    n_local = hi - lo
    if args.dtype == "bf16":
        rows = torch.empty((n_local, args.d), dtype=torch.bfloat16, device=dev)
        for a in range(lo, hi, 4 * CHUNK):
            b = min(hi, a + 4 * CHUNK)
            rows[a - lo: b - lo] = make_rows_fp32(a, b, args.d, dev).to(torch.bfloat16)
        db = rir.Database(rows, None, "bf16", idx_offset=lo)
    else:
        parts, scales = [], []
        for a in range(lo, hi, 4 * CHUNK):
            b = min(hi, a + 4 * CHUNK)
            r, s = rir.pack_descriptors(make_rows_fp32(a, b, args.d, dev), "fp8")
            parts.append(r)
            scales.append(s)
        db = rir.Database(torch.cat(parts), torch.cat(scales), "fp8", idx_offset=lo)
    sdb = rir.ShardedDatabase(db)
    exchange = "none (1 GPU)"
    if world > 1:
        # the one exchange step of the sharded search: NVLink peer-memory stores from the select kernel + a waiting
        # merge kernel (rir_sim_topk_sharded); --exchange nccl keeps the all-gather + merge path
        exchange = "nccl all-gather + merge kernel"
        if args.exchange == "peer" and sdb.enable_peer_exchange(args.nq, args.k):
            exchange = "nvlink peer-memory stores + waiting merge kernel (no collective call)"
    q_host = make_queries_fp32(args.nq, args.d).pin_memory()
    qr, qs = db.pack_queries(q_host.to(dev))
    k = args.k
    esz = 2 if args.dtype == "bf16" else 1

    # scan-kernel events: rir_sim_topk records (start, stop) around its full-scan launch when armed
    def step_resident():
        return sdb.search(qr, qs, k, path=args.path)

    out_host_s = torch.empty((args.nq, k), dtype=torch.float32).pin_memory()
    out_host_i = torch.empty((args.nq, k), dtype=torch.int32).pin_memory()

    host_call = args.dtype in ("bf16", "fp8") and (world == 1 or exchange.startswith("nvlink"))

    def step_e2e():
        if host_call:  # ONE C-ABI call on host buffers (rir_search_host) + a stream synchronise
            return sdb.query_host(q_host, k, out=(out_host_s, out_host_i), path=args.path)
        qd = q_host.to(dev, non_blocking=True)
        r, s = db.pack_queries(qd)
        sc, ix = sdb.search(r, s, k, path=args.path)
        out_host_s.copy_(sc, non_blocking=True)
        out_host_i.copy_(ix, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller consumes the result every step
        return out_host_s, out_host_i

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, events=None):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            if events is not None:
                lib.rir_profile_scan_events(events[i][0].cuda_event, events[i][1].cuda_event)
            fn()
        if events is not None:
            lib.rir_profile_scan_events(None, None)
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local_rank)
    events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in events:  # materialise the CUDA events before handing raw handles to the C ABI
        a.record()
        b.record()
    sampler.start()
    ms = timed(step_resident, args.steps, events)
    clocks = sampler.stop()
    scan_ms = [a.elapsed_time(b) for a, b in events]
    scan_avg_ms = sum(scan_ms) / len(scan_ms)

    for _ in range(3):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    value = args.nq * args.steps / (ms * 1e-3)
    e2e_value = args.nq * args.steps / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel (the scan): algorithmic work / measured launch duration ----
    # HBM-bound while 2*nq flop per database byte stays under the ridge (~214 flop/B, i.e. nq < ~200 for bf16),
    # tensor-bound above.
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    stream_path = args.path == "stream" or (args.path == "auto" and args.nq <= 2 and n_local < 2 * 148 * 256)
    kernel_name = "sim_stream_kernel (scan pass)" if stream_path else "sim_mma_kernel (fused sample + scan)"
    traffic = None  # dram bytes of one scan launch from the committed ncu capture of the SAME configuration, else null
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get(f"nq{args.nq}")
        if t and world == 1 and not stream_path and (args.n, args.d, args.dtype) == (N_DB, DIM, "bf16"):
            traffic = t["dram_bytes_per_launch"]
    except Exception:
        pass
    flops = 2.0 * args.nq * n_local * args.d
    alg_bytes = n_local * args.d * esz + args.nq * args.d * esz
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    if flops / alg_bytes < 214.0:
        achieved = alg_bytes / (scan_avg_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": traffic, "kernel": kernel_name, "kernel_ms": scan_avg_ms,
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "6650 GB/s (of fallback, B200_PROFILING.md)"}
    else:
        tf_peak = float(peaks.get("bf16_tflops", 1590.0)) * (2.0 if args.dtype == "fp8" else 1.0)
        achieved = flops / (scan_avg_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak,
                    "traffic": traffic, "kernel": kernel_name, "kernel_ms": scan_avg_ms,
                    "algorithmic_flops_per_launch": flops,
                    "peak_source": ("MEASURED_PEAKS.json bf16_tflops, burst: kernel timed alone (of measured)" if peaks
                                    else "1590 TFLOP/s (of fallback, B200_PROFILING.md)") +
                                   (" x2 for fp8" if args.dtype == "fp8" else "")}

    # our kernels per step (memsets are not kernels): fused tcgen05 path = scan + select (which also redoes overflowed
    # queries itself); the three-launch route adds sample + threshold; tiny shards: scan-all + select
    if n_local <= 16384:
        kernels_per_step = 2
    elif stream_path or n_local < 2 * 148 * 256 or args.k > 592:
        kernels_per_step = 4
    else:
        kernels_per_step = 2
    kernels_per_step += 1 if world > 1 else 0  # merge

    if rank == 0:
        line = {
            "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic", "config": workload_config(args, {"path": args.path, "exchange": exchange}),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": args.nq * args.d * 4,
                    "d2h_bytes_per_step": args.nq * k * 8, "ms_per_step": ms_e2e / args.steps,
                    "note": "rir_search_host: pinned fp32 host queries -> H2D -> pack -> search -> D2H (scores, idx) -> stream sync, "
                            "every step; database resident"},
            "gpu_launches": kernels_per_step * args.steps,
            "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            n_sample = min(args.n, 131072)
            times, times_topk = cpu_reference_step_time(args.nq, n_sample, args.d, reps=5, warm=1)
            scale = args.n / n_sample
            t_full = min(times) * scale
            line["cpu_baseline"] = {
                "value": args.nq / t_full, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                "sample": f"fp32 torch.mm + np.argsort(-sim,1) (iris_evaluate.py:383-386) on {args.nq} queries x {n_sample} rows, "
                          f"best of 5, time scaled x{scale:.3f} (linear in rows) to {args.n} rows",
                "topk_variant_value": args.nq / (min(times_topk) * scale) if times_topk else None,  # mm + torch.topk
            }
        print(json.dumps(line))
    if world > 1:
        sdb.close()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
