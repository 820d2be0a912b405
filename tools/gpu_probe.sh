#!/bin/bash
fmt='import sys,json
d=json.loads(sys.stdin.read()); r=d["roofline"]; print("nq=%d q/s=%.0f ms/step=%.3f scan_ms=%.3f tflops=%.0f"%(d["config"]["nq"],d["value"],d["ms_per_step"],r["kernel_ms"], 2*d["config"]["nq"]*d["config"]["n_db"]*d["config"]["dim"]/r["kernel_ms"]/1e9))'
run() { echo -n "$1 nq=$2 :: "; env $1 timeout 300 python bench.py --nq $2 --steps 20 --no-cpu-baseline 2>gpurun_out/sweep.err | tail -1 | python -c "$fmt" || tail -3 gpurun_out/sweep.err; }
run "RIR_MMA_TWO=1" 1024
run "RIR_MMA_TWO=1 RIR_MMA_DEBUG=2" 1024
run "RIR_MMA_TWO=1 RIR_MMA_DEBUG=8" 1024
run "RIR_MMA_TWO=1 RIR_MMA_DEBUG=16" 1024
run "RIR_MMA_TWO=1 RIR_MMA_MB=1" 1024
run "RIR_MMA_TWO=1 RIR_MMA_MB=1 RIR_MMA_DEBUG=8" 1024
run "RIR_MMA_TWO=1 RIR_MMA_MB=1 RIR_MMA_DEBUG=16" 1024
run "RIR_MMA_TWO=1 RIR_MMA_MB=1 RIR_MMA_DEBUG=2" 1024
