#!/bin/bash
# Short 8-GPU check: the bench at 4 and 8 GPUs (70 queries) and 1 query at 8 GPUs.
mkdir -p gpurun_out; rm -f gpurun_out/scale_short.jsonl
fmt='import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]); r=d["roofline"]; c=d["config"]; print("gpus=%d nq=%d q/s=%.0f ms/step=%.4f e2e=%.0f e2e_ms=%.4f scan_ms=%.4f frac=%.3f"%(d["n_gpus"],c["nq"],d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"]["ms_per_step"],r["kernel_ms"],r["frac"]))'
run() { N=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $N --warmup 5 --no-cpu-baseline "$@" 2> gpurun_out/bench_multi.err | tee -a gpurun_out/scale_short.jsonl | python -c "$fmt" || tail -5 gpurun_out/bench_multi.err; }
run 8 --steps 200 --nq 70
run 4 --steps 200 --nq 70
run 2 --steps 200 --nq 70
run 8 --steps 200 --nq 1
