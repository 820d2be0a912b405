#!/bin/bash
# all single-GPU suites + smoke + the multi-GPU suites and benches on N GPUs
mkdir -p gpurun_out
bash tools/gpu_check.sh; echo "gpu_check rc=$?"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
N=${N:-2} bash tools/gpu_multi.sh
fmt='import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]); r=d["roofline"]; e=d["e2e"]; print("nq=%d n=%d q/s=%.0f ms/step=%.4f (min %.4f) e2e=%.0f (%.4f) pipelined=%s scan_ms=%.4f frac=%.3f parity=%s"%(d["config"]["nq"],d["config"]["n_db"],d["value"],d["ms_per_step"],d["step_ms"]["min"],e["value"],e["ms_per_step"],(e.get("pipelined") or {}).get("ms_per_step"),r["kernel_ms"],r["frac"],d.get("parity",{}).get("vs_exact")))'
for n in 125916 251831 1007323; do
  echo "--- 1 GPU n=$n"; timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extras --n-db $n 2> gpurun_out/s8.err | python -c "$fmt" || tail -5 gpurun_out/s8.err
done
