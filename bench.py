#!/usr/bin/env python
"""bench.py — BASELINE.json's headline metric on synthetic data of the named shape.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, librir.so)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (torch.mm + np.argsort)

Workload (BASELINE.json configs[1], "RParis6k-shape with 1M distractors"): 70 queries x 1,007,323 database rows x
2048-d bf16, exact top-100; the database is row-sharded over the N GPUs (strong scaling: total work is fixed).
One step = one 70-query batch through the whole search path:
    fused tcgen05 scan (first round = sample, in-kernel threshold) -> exact select [-> NVLink peer exchange -> merge].
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
    ap.add_argument("--sync-exchange", action="store_true")  # N > 1: keep the merge on the search's own stream
    ap.add_argument("--no-parity", action="store_true")   # skip the in-process parity block
    ap.add_argument("--no-extras", action="store_true")   # skip the 1-query / 1024-query points
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
# reference arm: the reference's own CPU path (iris_evaluate.py:383-386), on the FULL configuration
# ----------------------------------------------------------------------------------------------
def make_rows_cpu_fp32(n: int, d: int):
    """n unit-norm fp32 rows in host memory, generated chunk-wise on all cores (numpy releases the GIL)."""
    import numpy as np
    import torch
    from concurrent.futures import ThreadPoolExecutor
    out = np.empty((n, d), dtype=np.float32)

    def fill(c):
        a, b = c * CHUNK, min(n, (c + 1) * CHUNK)
        blk = np.random.default_rng(SEED * 100003 + c).standard_normal((b - a, d), dtype=np.float32)
        blk /= np.linalg.norm(blk, axis=1, keepdims=True)
        out[a:b] = blk

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
        list(ex.map(fill, range(-(-n // CHUNK))))
    return torch.from_numpy(out)


def cpu_reference_steps(nq: int, n: int, d: int, k: int, reps: int, warm: int, topk_variant: bool = True):
    """The reference's similarity + ranking exactly as written — fp32 `torch.mm(q, g.t())` then
    `np.argsort(-similarity, axis=1)` (iris_evaluate.py:383-386, restated in oracle/search_oracle.py) — on ALL n rows,
    with every host core torch / numpy will use.  Returns (per-step seconds, seconds of the mm + torch.topk variant)."""
    import numpy as np
    import torch
    from oracle import search_oracle as S
    torch.set_num_threads(os.cpu_count() or 1)
    X = make_rows_cpu_fp32(n, d)
    Q = make_queries_fp32(nq, d)
    times = []
    for i in range(warm + reps):
        t0 = time.perf_counter()
        sim = S.similarity(Q, X).numpy()
        ranks = np.argsort(-sim, axis=1)
        t1 = time.perf_counter()
        if i >= warm:
            times.append(t1 - t0)
        del ranks
    t_topk = None
    if topk_variant:  # the "strong CPU" variant: mm + torch.topk(k) instead of the full argsort
        t0 = time.perf_counter()
        sim = S.similarity(Q, X)
        sc, ix = torch.topk(sim, k=min(k, n), dim=-1)
        t_topk = time.perf_counter() - t0
        del sc, ix
    return times, t_topk


def cpu_sample_text(args):
    return (f"fp32 torch.mm + np.argsort(-sim, axis=1) (iris_evaluate.py:383-386) on the full configuration: "
            f"{args.nq} queries x {args.n} rows x {args.d}-d per step, {os.cpu_count() or 1} host threads")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times, t_topk = cpu_reference_steps(args.nq, args.n, args.d, args.k, max(args.steps, 1), args.warmup)
    per_step = sum(times) / len(times)
    value = args.nq / per_step
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args),
        "step_ms": step_stats([t * 1e3 for t in times]),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": cpu_sample_text(args),
                         "topk_variant_value": (args.nq / t_topk) if t_topk else None},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def metric_name(args):
    if (args.n, args.d, args.k, args.dtype) == (N_DB, DIM, TOPK, "bf16"):
        return METRIC
    return (f"queries/sec exact top-{args.k} on {args.n} x {args.d} {args.dtype} db "
            "(development override of the BASELINE workload)")


def workload_config(args):
    """Identical in both arms (the driver compares the dicts): only what defines the workload."""
    default = (args.nq, args.n, args.d, args.k, args.dtype) == (NQ, N_DB, DIM, TOPK, "bf16")
    name = "BASELINE configs[1] (RParis6k-shape + 1M distractors)" if default else "development override"
    return {"workload": f"{name}: {args.nq} queries x {args.n} db rows x {args.d}-d {args.dtype}, exact top-{args.k}, "
                        f"db row-sharded over {args.gpus} GPU(s)",
            "nq": args.nq, "n_db": args.n, "dim": args.d, "k": args.k, "parallelism": f"db-shard x{args.gpus}",
            "l2_note": "per-GPU shard (>= 516 MB) exceeds the 126 MB L2, so every step re-reads it from HBM; no explicit flush"}


def step_stats(ms_list):
    s = sorted(ms_list)
    if not s:
        return None
    return {"min": s[0], "median": s[len(s) // 2], "max": s[-1], "n": len(s)}


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def tie_aware_mismatch(got_sc, got_ix, ref_sc, ref_ix, eps):
    """Two top-k lists of the same queries, both sorted by (score desc, index asc), scored by different arithmetic
    (tensor-core vs CUDA-core accumulation order).  They must agree position by position in score (relative eps) and
    in index wherever the neighbouring scores are further apart than eps.  Returns (ok, #index differences inside
    ties, message)."""
    import numpy as np
    tol = eps * np.maximum(np.abs(ref_sc), 1e-6)
    if got_sc.shape != ref_sc.shape or not np.all(np.abs(got_sc - ref_sc) <= tol):
        bad = np.argwhere(~(np.abs(got_sc - ref_sc) <= tol))[:3].tolist() if got_sc.shape == ref_sc.shape else "shape"
        return False, 0, f"scores differ beyond {eps} relative at {bad}"
    diff = got_ix != ref_ix
    n_tie = int(diff.sum())
    for q, j in np.argwhere(diff):
        # a differing index is only acceptable when `ref` holds another row within eps of this score (a tie)
        near = np.abs(ref_sc[q] - got_sc[q, j]) <= 2 * tol[q, j]
        if near.sum() < 2 and not (j == got_sc.shape[1] - 1):
            return False, n_tie, f"query {q} position {j}: index {got_ix[q, j]} vs {ref_ix[q, j]} without a tie"
    return True, n_tie, "ok"


def run_ours(args):
    import ctypes

    import numpy as np
    import torch
    import torch.distributed as dist

    import research_image_retrieval_b200 as rir
    from research_image_retrieval_b200 import _lib, search as rsearch

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
    steps, warmup = max(args.steps, 1), max(args.warmup, 3)

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
    extras_on = not args.no_extras and args.nq == NQ
    exchange = "none (1 GPU)"
    if world > 1:
        # the one exchange step of the sharded search: NVLink peer-memory stores from the select kernel + a waiting
        # merge kernel (rir_sim_topk_sharded); --exchange nccl keeps the all-gather + merge path
        exchange = "nccl all-gather + merge kernel"
        if args.exchange == "peer" and sdb.enable_peer_exchange(max(args.nq, 1024 if extras_on else 1), args.k):
            exchange = "nvlink peer-memory stores + waiting merge kernel (no collective call)"
    k = args.k
    esz = 2 if args.dtype == "bf16" else 1
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    bf16_burst = float(peaks.get("bf16_tflops", 1590.0))
    bf16_sustained = float(peaks.get("bf16_tflops_sustained", 1400.0))
    fp8_peak = None
    try:  # measured on this pool by tools/measure_fp8_peak.py (an own tcgen05 kind::f8f6f4 loop + torch._scaled_mm)
        with open(os.path.join(ROOT, "profiles", "fp8_peak.json")) as f:
            fp8_peak = float(json.load(f)["fp8_tflops"])
    except Exception:
        pass

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(nq, n_steps, n_warm, with_e2e):
        """n_steps timed steps of an nq-query batch: resident (packed queries in HBM) and, optionally, end to end
        through the host-buffer call.  Device-timed with CUDA events, max over ranks."""
        q_host = make_queries_fp32(nq, args.d).pin_memory()
        qr, qs = db.pack_queries(q_host.to(dev))
        out_s = torch.empty((nq, k), dtype=torch.float32).pin_memory()
        out_i = torch.empty((nq, k), dtype=torch.int32).pin_memory()
        host_call = world == 1 or exchange.startswith("nvlink")

        res_out = (torch.empty((nq, k), dtype=torch.float32, device=dev), torch.empty((nq, k), dtype=torch.int32, device=dev))
        peer_or_single = world == 1 or exchange.startswith("nvlink")

        res_out2 = (torch.empty_like(res_out[0]), torch.empty_like(res_out[1]))
        state = {"i": 0, "prev": None}
        async_exchange = world > 1 and exchange.startswith("nvlink") and not args.sync_exchange

        def step_resident():
            if async_exchange:
                # the merge of step i runs on the exchange's side stream next to the scan of step i+1; this stream
                # waits for the merge of step i-1 (its results are consumed now) — every step's result is joined
                h = sdb.search_async(qr, qs, k, path=args.path, out=res_out if state["i"] & 1 == 0 else res_out2)
                state["i"] += 1
                if state["prev"] is not None:
                    state["prev"].wait()
                state["prev"] = h
                return h
            return sdb.search(qr, qs, k, path=args.path, out=res_out if peer_or_single else None)

        def drain_resident():
            if state["prev"] is not None:
                state["prev"].wait()
                state["prev"] = None

        def step_e2e():
            if host_call:  # ONE C-ABI call on host buffers (rir_search_host) + a stream synchronise
                return sdb.query_host(q_host, k, out=(out_s, out_i), path=args.path)
            qd = q_host.to(dev, non_blocking=True)
            r, s = db.pack_queries(qd)
            sc, ix = sdb.search(r, s, k, path=args.path)
            out_s.copy_(sc, non_blocking=True)
            out_i.copy_(ix, non_blocking=True)
            torch.cuda.current_stream().synchronize()  # the caller consumes the result every step
            return out_s, out_i

        SAMPLE = 4   # every 4th step carries the scan-kernel events and a step mark (see rir_profile_scan_pause)

        def timed(fn, profile_scan):
            marks = [(0, torch.cuda.Event(enable_timing=True))]
            e_end = torch.cuda.Event(enable_timing=True)
            sync_all()
            if profile_scan:
                lib.rir_profile_scan_begin()
            marks[0][1].record()
            sampled = 0
            for i in range(n_steps):
                if profile_scan:
                    lib.rir_profile_scan_pause(0 if i % SAMPLE == 0 else 1)
                    sampled += 1 if i % SAMPLE == 0 else 0
                fn()
                if i == n_steps - 1 and fn is step_resident:
                    drain_resident()          # the last step's merge is inside the timed region too
                if (i + 1) % SAMPLE == 0 and i + 1 < n_steps:
                    ev = torch.cuda.Event(enable_timing=True)
                    ev.record()
                    marks.append((i + 1, ev))
            e_end.record()
            marks.append((n_steps, e_end))
            sync_all()
            scan = None
            if profile_scan:
                cap = 64 * n_steps + 64
                buf = (ctypes.c_float * cap)()
                cnt = ctypes.c_int(0)
                _lib.check(lib.rir_profile_scan_end(buf, cap, ctypes.byref(cnt)))
                scan = [float(buf[i]) for i in range(min(cnt.value, cap))]
            ms = marks[0][1].elapsed_time(e_end)
            per = [a[1].elapsed_time(b[1]) / (b[0] - a[0]) for a, b in zip(marks[:-1], marks[1:])]  # ms per step, per window
            if world > 1:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            return ms, per, scan, sampled

        for _ in range(n_warm):
            step_resident()
        drain_resident()
        ms, per, scan, sampled = timed(step_resident, True)
        res = {"ms": ms, "per_step": per, "scan_ms_per_step": sum(scan) / sampled if scan else None,
               "scan_launches_per_step": (len(scan) / sampled) if scan else 0, "scan_steps_sampled": sampled,
               "qr": qr, "qs": qs}
        if with_e2e:
            for _ in range(3):
                step_e2e()
            res["ms_e2e"], res["per_step_e2e"], _, _ = timed(step_e2e, False)
            if host_call:
                # the serving loop (HostQueryPipeline): the same copies every step, but two batches in flight and the
                # H2D + pack of step i+1 on a copy stream under the scan of step i — step i+1 is enqueued before the
                # host waits for and reads step i, so the device does not idle during the host's turnaround
                pipe = rir.HostQueryPipeline(sdb, nq, k, depth=2, path=args.path)
                qh = [q_host, q_host.clone().pin_memory()]

                def run_pipelined(n):
                    prev, acc = None, 0.0
                    for i in range(n):
                        h = pipe.submit(qh[i & 1])
                        if prev is not None:
                            sc, _ = prev.result()
                            acc += float(sc[0, 0])           # the host consumes every step's result
                        prev = h
                    sc, _ = prev.result()
                    return acc + float(sc[0, 0])

                run_pipelined(4)
                sync_all()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                run_pipelined(n_steps)
                e1.record()
                sync_all()
                ms_p = e0.elapsed_time(e1)
                if world > 1:
                    t = torch.tensor([ms_p], device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms_p = float(t.item())
                res["ms_e2e_pipelined"] = ms_p
        return res

    def kernels_per_step(nq):
        """Our kernels per resident step (no memsets any more: the select kernel leaves the workspace header clean)."""
        groups = -(-nq // 4096)
        if n_local <= 16384:
            per_group = 2                                   # scan-all + select
        elif args.path == "stream" or n_local < 2 * 148 * 256 or args.k > 592:
            per_group = 4                                   # sample, threshold, scan, select
        else:
            per_group = 2                                   # fused scan + select
        # + the exchange merge: its own kernel on the side stream (asynchronous exchange) or for batches beyond one CTA
        # per SM; otherwise folded into the select kernel
        merge = 0 if world == 1 else (1 if (not args.sync_exchange or nq > 148 or args.exchange == "nccl") else 0)
        return groups * per_group + merge

    def roofline_of(nq, scan_ms):
        flops = 2.0 * nq * n_local * args.d
        alg_bytes = n_local * args.d * esz + nq * args.d * esz
        if flops / alg_bytes < 214.0:
            achieved = alg_bytes / (scan_ms * 1e-3) / 1e9
            return {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "kernel_ms": scan_ms, "algorithmic_bytes_per_launch": alg_bytes,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "6650 GB/s (of fallback, B200_PROFILING.md)"}
        if args.dtype == "fp8":
            tf_peak = fp8_peak if fp8_peak else 2.0 * bf16_burst
            src = ("profiles/fp8_peak.json (of measured, tools/measure_fp8_peak.py)" if fp8_peak
                   else "2 x MEASURED_PEAKS.json bf16_tflops (assumed: no fp8 measurement committed)")
        else:
            tf_peak = bf16_burst
            src = ("MEASURED_PEAKS.json bf16_tflops, burst: kernel timed alone (of measured)" if peaks
                   else "1590 TFLOP/s (of fallback, B200_PROFILING.md)")
        achieved = flops / (scan_ms * 1e-3) / 1e12
        return {"bound": "tensor", "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak,
                "kernel_ms": scan_ms, "algorithmic_flops_per_launch": flops, "peak_source": src}

    # ---- headline: the BASELINE batch ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    head = measure(args.nq, steps, warmup, with_e2e=True)
    clocks = sampler.stop()
    value = args.nq * steps / (head["ms"] * 1e-3)
    # End to end from HOST buffers, every step: H2D of that step's pinned fp32 queries, pack, search, the top-k written
    # into pinned host buffers and read by the host.  Headline = the serving loop (HostQueryPipeline: two batches in
    # flight, copies on a copy stream); the one-batch-at-a-time flavour (a stream synchronise after every step, the
    # device idle during the host's turnaround) is reported beside it.
    sync_e2e = {"value": args.nq * steps / (head["ms_e2e"] * 1e-3), "unit": UNIT, "ms_per_step": head["ms_e2e"] / steps,
                "step_ms": step_stats(head["per_step_e2e"]),
                "note": "rir_search_host (ONE C-ABI call on host buffers) + a stream synchronise after every step: one batch "
                        "in flight, the device idles while the host turns around"}
    e2e = {"unit": UNIT, "h2d_bytes_per_step": args.nq * args.d * 4, "d2h_bytes_per_step": args.nq * k * 8,
           "gpu_launches": (kernels_per_step(args.nq) + 1) * steps}
    if "ms_e2e_pipelined" in head:
        e2e.update({"value": args.nq * steps / (head["ms_e2e_pipelined"] * 1e-3), "ms_per_step": head["ms_e2e_pipelined"] / steps,
                    "note": "HostQueryPipeline.submit / .result(): per step pinned fp32 host queries -> H2D -> pack (copy "
                            "stream) -> search -> top-k stored into pinned host buffers -> read by the host; TWO batches in "
                            "flight, database resident", "sync_each_step": sync_e2e})
    else:
        e2e.update({k_: v_ for k_, v_ in sync_e2e.items()})
    stream_path = args.path == "stream" or (args.path == "auto" and args.nq <= 2 and n_local < 2 * 148 * 256)
    roofline = roofline_of(args.nq, head["scan_ms_per_step"])
    roofline["kernel"] = "sim_stream_kernel (scan pass)" if stream_path else "sim_mma_kernel (fused sample + scan)"
    roofline["kernel_ms_source"] = (f"CUDA events around the scan launch on its stream, every 4th step of the timed region "
                                    f"({head['scan_steps_sampled']} of {steps} steps; an event pair per step would sit between "
                                    "PDL-chained kernels and cost several percent of a sharded step)")
    roofline["traffic"] = None  # dram bytes of one scan launch from the committed ncu capture of the SAME configuration
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        t = tj.get(f"nq{args.nq}")
        if t and world == 1 and not stream_path and (args.n, args.d, args.dtype) == (N_DB, DIM, "bf16"):
            roofline["traffic"] = t["dram_bytes_per_launch"]
            roofline["traffic_source"] = f"profiles/traffic.json ({tj.get('captured_at', 'ncu --set full capture')})"
    except Exception:
        pass

    # ---- parity, visible to the driver: the timed path against independent implementations, same process ----
    parity = None
    if not args.no_parity:
        nqc = min(8, args.nq)
        qr, qs = head["qr"], head["qs"]
        if world > 1 and exchange.startswith("nvlink") and not args.sync_exchange:
            sc_p, ix_p = sdb.search_async(qr, qs, k, path=args.path).wait()   # the path that was timed
            sc_s, ix_s = sdb.search(qr, qs, k, path=args.path)                # and the in-stream-order flavour of it
            assert torch.equal(ix_p, ix_s) and torch.equal(sc_p, sc_s), "asynchronous exchange != synchronous exchange"
        else:
            sc_p, ix_p = sdb.search(qr, qs, k, path=args.path)            # the path that was timed
        peer_eq_nccl = None
        if world > 1 and exchange.startswith("nvlink"):
            sc_n, ix_n = sdb.search(qr, qs, k, path=args.path, exchange="nccl")
            flag = torch.tensor([int(torch.equal(ix_p, ix_n) and torch.equal(sc_p, sc_n))], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            peer_eq_nccl = bool(flag.item())
        # independent: RIR_PATH_EXACT (one CTA per query, CUDA-core fp32 FMAs, running top-k — shares no scan / select
        # code with the timed path) on every shard, all-gather, host merge with the order rule (score desc, index asc)
        k_loc = min(k, n_local)
        q8 = qr[:nqc].contiguous()
        s8 = None if qs is None else qs[:nqc].contiguous()
        sc_x, ix_x = rsearch.sim_topk(q8, db.rows, k_loc, dtype=args.dtype, q_scale=s8, x_scale=db.scale,
                                      idx_offset=lo, path="exact")
        sc_x, ix_x = rsearch.pad_topk(sc_x, ix_x, k)
        if world > 1:
            all_s, all_i = rsearch.gather_topk(sc_x, ix_x, world)
        else:
            all_s, all_i = sc_x[None], ix_x[None]
        ref_sc, ref_ix = rsearch.merge_topk_host(all_s.cpu().numpy(), all_i.cpu().numpy(), k)
        ok, n_tie, msg = tie_aware_mismatch(sc_p[:nqc].cpu().numpy(), ix_p[:nqc].cpu().numpy(), ref_sc, ref_ix, 1e-3)
        flag = torch.tensor([int(ok)], device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        parity = {"peer_eq_nccl": peer_eq_nccl, "vs_exact": "ok" if bool(flag.item()) else f"MISMATCH: {msg}",
                  "queries": nqc, "index_differences_inside_ties": n_tie, "tolerance": "1e-3 relative (bf16 bar)",
                  "note": "timed path vs RIR_PATH_EXACT per shard + host merge" +
                          ("; peer-memory exchange vs NCCL all-gather path bit-identical on all queries" if peer_eq_nccl else "")}

    # ---- the two north-star regimes, same process: 1 query (HBM-bound) and 1024 queries (tensor-bound) ----
    extras = None
    if extras_on:
        extras = {}
        r1 = measure(1, 40, 5, with_e2e=False)
        rf = roofline_of(1, r1["scan_ms_per_step"])
        extras["q1"] = {"value": 1 * 40 / (r1["ms"] * 1e-3), "unit": UNIT, "ms_per_step": r1["ms"] / 40,
                        "step_ms": step_stats(r1["per_step"]),
                        "scan_ms": r1["scan_ms_per_step"], "hbm_gbs": rf["achieved"], "hbm_frac": rf["frac"],
                        "target": "north_star: >= 0.80 of HBM peak at 1 query"}
        rk = measure(1024, 12, 3, with_e2e=False)
        rf = roofline_of(1024, rk["scan_ms_per_step"])
        extras["q1024"] = {"value": 1024 * 12 / (rk["ms"] * 1e-3), "unit": UNIT, "ms_per_step": rk["ms"] / 12,
                           "step_ms": step_stats(rk["per_step"]),
                           "scan_ms": rk["scan_ms_per_step"], "tflops": rf["achieved"],
                           "frac_of_burst": rf["achieved"] / bf16_burst if args.dtype == "bf16" else rf["frac"],
                           "frac_of_sustained": rf["achieved"] / bf16_sustained if args.dtype == "bf16" else None,
                           "target": "north_star: >= 0.60 tensor pipe at 1k queries"}

    if rank == 0:
        line = {
            "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": head["ms"] / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic", "config": workload_config(args),
            "impl_detail": {"path": args.path, "exchange": exchange,
                            "exchange_mode": None if world == 1 else (
                                "asynchronous: merge of step i on a side stream under the scan of step i+1, every result "
                                "joined before the next one is issued (ShardedDatabase.search_async)"
                                if (exchange.startswith("nvlink") and not args.sync_exchange) else "in stream order")},
            "step_ms": dict(step_stats(head["per_step"]), note="ms per step over 4-step windows of the timed region"),
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": kernels_per_step(args.nq) * steps,
            "roofline": roofline,
        }
        if parity is not None:
            line["parity"] = parity
        if extras is not None:
            line["extras"] = extras
        if world == 1 and not args.no_cpu_baseline:
            times, t_topk = cpu_reference_steps(args.nq, args.n, args.d, args.k, reps=3, warm=1)
            line["cpu_baseline"] = {
                "value": args.nq / min(times), "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                "sample": cpu_sample_text(args) + ", best of 3 after 1 warm-up",
                "topk_variant_value": (args.nq / t_topk) if t_topk else None,  # mm + torch.topk
            }
        print(json.dumps(line))
    if world > 1:
        sdb.close()
        dist.destroy_process_group()
    if parity is not None and (parity["vs_exact"] != "ok" or parity["peer_eq_nccl"] is False):
        raise SystemExit(3)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
