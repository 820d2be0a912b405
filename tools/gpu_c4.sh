#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest -x -q -m gpu tests/test_gpu_search.py -k "fused_scan or boundaries or paths" > gpurun_out/test_c4.log 2>&1; tail -4 gpurun_out/test_c4.log
fmt='import sys,json
d=json.loads(sys.stdin.read()); r=d["roofline"]; print("nq=%d q/s=%.0f ms/step=%.3f scan_ms=%.3f frac=%.3f tflops=%.0f"%(d["config"]["nq"],d["value"],d["ms_per_step"],r["kernel_ms"],r["frac"], 2*d["config"]["nq"]*d["config"]["n_db"]*d["config"]["dim"]/r["kernel_ms"]/1e9))'
run() { echo -n "$1 nq=$2 :: "; env $1 timeout 300 python bench.py --nq $2 --steps 30 --no-cpu-baseline 2>gpurun_out/sweep.err | tail -1 | python -c "$fmt" || tail -3 gpurun_out/sweep.err; }
for q in 256 1024 4096; do
run "RIR_MMA_CLUSTER4=1" $q
run "RIR_MMA_CLUSTER4=0" $q
done
