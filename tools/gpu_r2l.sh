#!/bin/bash
# (2 GPUs) does the cooperative scan wait for the side-stream merge?  + verification of the pending small changes
mkdir -p gpurun_out
fmt='import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]); r=d["roofline"]; e=d["e2e"]; p=d.get("parity") or {}
print("gpus=%d nq=%d q/s=%.0f ms/step=%.4f (min %.4f) e2e=%.0f sync_e2e=%s scan_ms=%.4f parity=%s/%s"%(d["n_gpus"],d["config"]["nq"],d["value"],d["ms_per_step"],d["step_ms"]["min"],e["value"],(e.get("sync_each_step") or {}).get("value"),r["kernel_ms"],p.get("peer_eq_nccl"),p.get("vs_exact")))'
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --no-extras --n-db 251831 "$@" 2> gpurun_out/bench_multi_2.err; }
echo "--- async (default)"; run --steps 200 --warmup 10 | python -c "$fmt"
echo "--- async, no cooperative attribute"; RIR_FUSED_LAUNCH_MODE=4 run --steps 200 --warmup 10 | python -c "$fmt"
echo "--- sync exchange (folded merge)"; run --steps 200 --warmup 10 --sync-exchange | python -c "$fmt"
timeout 600 python -m pytest -x -q -m gpu tests/test_gpu_descriptor.py tests/test_gpu_search.py -k "pca or whiten or head or host or pipeline" > gpurun_out/test_misc.log 2>&1; echo "misc tests rc=$?"; tail -3 gpurun_out/test_misc.log
timeout 300 python tools/bench_descriptor.py --pca 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "$fmt"
