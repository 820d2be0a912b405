#!/bin/bash
fmt='import sys,json
d=json.loads(sys.stdin.read()); r=d["roofline"]; print("nq=%d q/s=%.0f ms/step=%.3f scan_ms=%.3f hbm_frac=%.3f tflops=%.0f"%(d["config"]["nq"],d["value"],d["ms_per_step"],r["kernel_ms"],r["frac"], 2*d["config"]["nq"]*d["config"]["n_db"]*d["config"]["dim"]/r["kernel_ms"]/1e9))'
run() { echo -n "$1 nq=$2 :: "; env $1 python bench.py --nq $2 --steps 30 --no-cpu-baseline 2>gpurun_out/sweep.err | tail -1 | python -c "$fmt" || tail -3 gpurun_out/sweep.err; }
for q in 16 70 1024; do
run "X=0" $q
run "RIR_MMA_DEBUG=1" $q
run "RIR_MMA_DEBUG=2" $q
run "RIR_MMA_DEBUG=3" $q
done
run "RIR_MMA_MB=1 RIR_MMA_DEBUG=3" 1024
run "RIR_MMA_MB=1 RIR_MMA_DEBUG=1" 1024
run "RIR_MMA_MB=1 RIR_MMA_DEBUG=2" 1024
run "RIR_MMA_CLUSTER=1 RIR_MMA_DEBUG=3" 1024
run "RIR_MMA_CLUSTER=1 RIR_MMA_DEBUG=3" 70
run "RIR_MMA_TRIM=0 RIR_MMA_DEBUG=3" 70
run "RIR_MMA_TRIM=0" 16
