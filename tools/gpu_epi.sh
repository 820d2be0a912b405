#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest -x -q -m gpu tests/test_gpu_search.py -k "fused_scan or boundaries or overflow or clustered" > gpurun_out/test_epi.log 2>&1; tail -3 gpurun_out/test_epi.log
fmt='import sys,json
d=json.loads(sys.stdin.read()); r=d["roofline"]; print("nq=%d q/s=%.0f ms/step=%.3f scan_ms=%.3f frac=%.3f tflops=%.0f"%(d["config"]["nq"],d["value"],d["ms_per_step"],r["kernel_ms"],r["frac"], 2*d["config"]["nq"]*d["config"]["n_db"]*d["config"]["dim"]/r["kernel_ms"]/1e9))'
run() { echo -n "$1 nq=$2 :: "; env $1 timeout 300 python bench.py --nq $2 --steps 30 --no-cpu-baseline 2>gpurun_out/sweep.err | tail -1 | python -c "$fmt" || tail -3 gpurun_out/sweep.err; }
run "X=0" 70
run "X=0" 1024
run "RIR_MMA_MB=2" 1024
run "RIR_MMA_MB=2 RIR_MMA_DEBUG=2" 1024
run "X=0" 4096
run "RIR_MMA_MB=2" 4096
