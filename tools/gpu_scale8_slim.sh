#!/bin/bash
# slim 8-GPU call: the committed bench at 4 and 8 GPUs (peer exchange, async), 1-query batches at 8 GPUs
mkdir -p gpurun_out
fmt='import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]); r=d["roofline"]; e=d["e2e"]; p=d.get("parity") or {}; x=d.get("extras") or {}
print("gpus=%d nq=%d q/s=%.0f ms/step=%.4f (min %.4f) e2e=%.0f sync_e2e=%s scan_ms=%.4f frac=%.3f parity=%s/%s q1=%s q1024=%s"%(d["n_gpus"],d["config"]["nq"],d["value"],d["ms_per_step"],d["step_ms"]["min"],e["value"],(e.get("sync_each_step") or {}).get("value"),r["kernel_ms"],r["frac"],p.get("peer_eq_nccl"),p.get("vs_exact"),(x.get("q1") or {}).get("value"),(x.get("q1024") or {}).get("value")))'
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n "$@" 2> gpurun_out/bench_multi_$n.err; }
run 8 --steps 200 --warmup 10 | tee gpurun_out/final_n8_peer.json | python -c "$fmt" || grep -v "^\[W\|Warning\|^$" gpurun_out/bench_multi_8.err | tail -5
run 8 --steps 20 --warmup 5 | tee gpurun_out/final_n8_peer_k20.json | python -c "$fmt"
run 4 --steps 200 --warmup 10 | tee gpurun_out/final_n4_peer.json | python -c "$fmt"
run 8 --steps 200 --warmup 10 --nq 1 --no-extras | tee gpurun_out/final_n8_q1.json | python -c "$fmt"
