#!/bin/bash
mkdir -p gpurun_out
RIR_MMA_TWO=1 timeout 600 python -m pytest -x -q -m gpu tests/test_gpu_search.py -k "fused_scan" > gpurun_out/test_two.log 2>&1; tail -5 gpurun_out/test_two.log
RIR_MMA_TWO=1 RIR_MMA_MB=1 timeout 600 python -m pytest -x -q -m gpu tests/test_gpu_search.py -k "fused_scan" > gpurun_out/test_two_mb1.log 2>&1; tail -5 gpurun_out/test_two_mb1.log
fmt='import sys,json
d=json.loads(sys.stdin.read()); r=d["roofline"]; print("nq=%d q/s=%.0f ms/step=%.3f e2e=%.0f scan_ms=%.3f hbm_frac=%.3f tflops=%.0f"%(d["config"]["nq"],d["value"],d["ms_per_step"],d["e2e"]["value"],r["kernel_ms"],r["frac"], 2*d["config"]["nq"]*d["config"]["n_db"]*d["config"]["dim"]/r["kernel_ms"]/1e9))'
run() { echo -n "$1 nq=$2 :: "; env $1 timeout 300 python bench.py --nq $2 --steps 30 --no-cpu-baseline 2>gpurun_out/sweep.err | tail -1 | python -c "$fmt" || tail -3 gpurun_out/sweep.err; }
run "X=0" 1024
run "RIR_MMA_TWO=1" 1024
run "RIR_MMA_TWO=1 RIR_MMA_MB=1" 1024
run "RIR_MMA_MB=1" 1024
run "RIR_MMA_TWO=1 RIR_MMA_DEBUG=2" 1024
run "RIR_MMA_TWO=1 RIR_MMA_MB=1 RIR_MMA_DEBUG=2" 1024
run "RIR_MMA_TWO=1" 512
run "X=0" 512
run "RIR_MMA_TWO=1 RIR_MMA_MB=1" 256
run "X=0" 256
run "RIR_MMA_TWO=1" 4096
