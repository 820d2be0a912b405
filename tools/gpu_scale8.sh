#!/bin/bash
mkdir -p gpurun_out
N=${N:-8}
fmt='import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]); r=d["roofline"]; print("gpus=%d nq=%d q/s=%.0f ms/step=%.3f e2e=%.0f e2e_ms=%.3f scan_ms=%.3f frac=%.3f  %s"%(d["n_gpus"],d["config"]["nq"],d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"]["ms_per_step"],r["kernel_ms"],r["frac"],d["config"]["exchange"][:12]))'
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline "$@" 2> gpurun_out/bench_multi.err | tee -a gpurun_out/scale_n$N.jsonl | python -c "$fmt" || tail -5 gpurun_out/bench_multi.err; }
run --nq 70 --exchange peer
run --nq 70 --exchange nccl
run --nq 1 --exchange peer
run --nq 1024 --exchange peer
