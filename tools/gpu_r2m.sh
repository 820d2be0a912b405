#!/bin/bash
# (2 GPUs; run on the deferred-merge variant, commit 87d3f32..35a0a04) multi-rank tests, single-GPU suites, benches at full and 1/8-size shards
mkdir -p gpurun_out
timeout 900 python -m pytest -x -q -m gpu tests/test_gpu_multi.py > gpurun_out/test_multi.log 2>&1; echo "test_multi rc=$?"; tail -4 gpurun_out/test_multi.log; grep -E "Error|assert" gpurun_out/multi_worker_w2.log | head -5
bash tools/gpu_check.sh; echo "gpu_check rc=$?"
fmt='import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]); r=d["roofline"]; e=d["e2e"]; p=d.get("parity") or {}
print("gpus=%d nq=%d n=%d q/s=%.0f ms/step=%.4f (min %.4f) e2e=%.0f sync_e2e=%s scan_ms=%.4f parity=%s/%s"%(d["n_gpus"],d["config"]["nq"],d["config"]["n_db"],d["value"],d["ms_per_step"],d["step_ms"]["min"],e["value"],(e.get("sync_each_step") or {}).get("value"),r["kernel_ms"],p.get("peer_eq_nccl"),p.get("vs_exact")))'
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 "$@" 2> gpurun_out/bench_multi_2.err; }
echo "--- 2 GPUs, 1/8-size shards, async"; run --steps 200 --warmup 10 --no-extras --n-db 251831 | python -c "$fmt" || grep -v "^\[W\|Warning\|^$" gpurun_out/bench_multi_2.err | tail -8
echo "--- 2 GPUs, 1/8-size shards, sync"; run --steps 200 --warmup 10 --no-extras --n-db 251831 --sync-exchange | python -c "$fmt"
echo "--- 2 GPUs full"; run --steps 20 --warmup 5 | python -c "$fmt" || grep -v "^\[W\|Warning\|^$" gpurun_out/bench_multi_2.err | tail -8
echo "--- 1 GPU 1/8 shard"; timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extras --n-db 125916 2>/dev/null | python -c "$fmt"
echo "--- 1 GPU full"; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "$fmt"
