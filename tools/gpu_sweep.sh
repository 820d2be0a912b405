#!/bin/bash
# development sweep: cluster size x query count
fmt='import sys,json
d=json.loads(sys.stdin.read()); r=d["roofline"]; print("q/s=%.0f ms/step=%.3f scan_ms=%.3f hbm_frac=%.3f tflops=%.0f"%(d["value"],d["ms_per_step"],r["kernel_ms"],r["frac"], 2*d["config"]["nq"]*d["config"]["n_db"]*d["config"]["dim"]/r["kernel_ms"]/1e9))'
for c in 1 2 4; do for q in 70 128 256 1024; do echo -n "cluster $c nq $q: "; RIR_MMA_CLUSTER=$c python bench.py --nq $q --steps 30 --no-cpu-baseline 2>&1 | tail -1 | python -c "$fmt"; done; done
