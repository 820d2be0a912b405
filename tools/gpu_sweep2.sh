#!/bin/bash
fmt='import sys,json
d=json.loads(sys.stdin.read()); r=d["roofline"]; print("q/s=%.0f ms/step=%.3f scan_ms=%.3f hbm_frac=%.3f"%(d["value"],d["ms_per_step"],r["kernel_ms"],r["frac"]))'
for c in 1 2; do for r in "2,6" "3,5" "4,4"; do echo -n "cluster $c rings $r nq 70: "; RIR_MMA_RINGS=$r RIR_MMA_CLUSTER=$c python bench.py --nq 70 --steps 50 --no-cpu-baseline 2>&1 | tail -1 | python -c "$fmt"; done; done
for q in 5 8 16 24 32 48 64; do echo -n "nq $q: "; python bench.py --nq $q --steps 50 --no-cpu-baseline 2>&1 | tail -1 | python -c "$fmt"; done
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_q70c.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sim_mma_kernel -s 7 -c 1 -o gpurun_out/prof_mma_q70_v2 $B > gpurun_out/ncu_full_q70_v2.log 2>&1
