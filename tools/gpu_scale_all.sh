#!/bin/bash
# One 8-GPU box: strong-scaling points of the bench (N = 2, 4, 8) + the other operating points at N = 8.
mkdir -p gpurun_out; rm -f gpurun_out/scale_all.jsonl
fmt='import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]); r=d["roofline"]; c=d["config"]; print("gpus=%d %s nq=%d n=%d k=%d q/s=%.0f ms/step=%.4f e2e=%.0f e2e_ms=%.4f scan_ms=%.4f %s frac=%.3f  %s"%(d["n_gpus"],d["dtype"],c["nq"],c["n_db"],c["k"],d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"]["ms_per_step"],r["kernel_ms"],r["bound"],r["frac"],c["exchange"][:12]))'
run() { N=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $N --warmup 5 --no-cpu-baseline "$@" 2> gpurun_out/bench_multi.err | tee -a gpurun_out/scale_all.jsonl | python -c "$fmt" || tail -5 gpurun_out/bench_multi.err; }
python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | tee -a gpurun_out/scale_all.jsonl | python -c "$fmt"
run 2 --steps 100 --nq 70
run 4 --steps 100 --nq 70
run 8 --steps 100 --nq 70
run 8 --steps 100 --nq 70 --exchange nccl
run 8 --steps 100 --nq 1
run 8 --steps 50 --nq 1024
run 8 --steps 3 --nq 100000 --n-db 1580470 --k 10 --dtype fp8
