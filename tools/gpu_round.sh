#!/bin/bash
# One call: GPU parity suites, smoke, headline bench lines, launch list and full ncu captures (each ncu command runs
# only after the identical plain command exited 0).  Outputs under gpurun_out/.
mkdir -p gpurun_out
bash tools/gpu_check.sh
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
STEPS=100 bash tools/gpu_bench.sh > gpurun_out/bench_all.log 2>&1; grep -c '"metric"' gpurun_out/bench_all.log
fmt='import sys,json
d=json.loads(sys.stdin.read()); r=d["roofline"]; print("nq=%d path=%s q/s=%.0f ms/step=%.3f e2e=%.0f scan_ms=%.3f hbm_frac=%.3f tflops=%.0f"%(d["config"]["nq"],d["config"]["path"],d["value"],d["ms_per_step"],d["e2e"]["value"],r["kernel_ms"],r["frac"], 2*d["config"]["nq"]*d["config"]["n_db"]*d["config"]["dim"]/r["kernel_ms"]/1e9))'
for q in 1 2 4 8; do python bench.py --nq $q --path mma --steps 50 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "$fmt"; done
for q in 2 3 4; do python bench.py --nq $q --path stream --steps 50 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "$fmt"; done
bash tools/gpu_ncu.sh > gpurun_out/ncu_all.log 2>&1; tail -5 gpurun_out/ncu_all.log
