#!/bin/bash
fmt='import sys,json
d=json.loads(sys.stdin.read()); r=d["roofline"]; print("nq=%d q/s=%.0f ms/step=%.3f scan_ms=%.3f hbm_frac=%.3f tflops=%.0f"%(d["config"]["nq"],d["value"],d["ms_per_step"],r["kernel_ms"],r["frac"], 2*d["config"]["nq"]*d["config"]["n_db"]*d["config"]["dim"]/r["kernel_ms"]/1e9))'
run() { echo -n "$1 :: "; env $1 python bench.py --nq $2 --steps 40 --no-cpu-baseline 2>gpurun_out/sweep.err | tail -1 | python -c "$fmt" || tail -3 gpurun_out/sweep.err; }
run "X=0" 70
run "RIR_MMA_PERM=0" 70
run "RIR_MMA_FUSED=0" 70
run "RIR_MMA_FUSED=0 RIR_MMA_TRIM=0" 70
run "RIR_MMA_TRIM=0" 70
run "RIR_MMA_CLUSTER=1" 70
run "RIR_MMA_CLUSTER=1 RIR_MMA_PERM=0" 70
run "RIR_MMA_RINGS=2,5" 70
run "X=0" 16
run "X=0" 1024
run "RIR_MMA_PERM=0" 1024
run "RIR_MMA_FUSED=0" 1024
run "RIR_MMA_MB=1" 1024
run "RIR_MMA_CLUSTER=1" 1024
run "RIR_MMA_RINGS=2,4" 1024
run "X=0" 256
run "RIR_MMA_MB=1" 256
