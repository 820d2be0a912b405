#!/bin/bash
# tensor regime: blocks per CTA (MB) and epilogue headroom at 1024 / 2048 / 4096 queries
mkdir -p gpurun_out
fmt='import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]); r=d["roofline"]; print("nq=%d q/s=%.0f ms/step=%.4f scan_ms=%.4f %s=%.1f frac=%.3f parity=%s"%(d["config"]["nq"],d["value"],d["ms_per_step"],r["kernel_ms"],r["unit"],r["achieved"],r["frac"],d.get("parity",{}).get("vs_exact")))'
for q in 1024 2048 4096; do for e in "RIR_MMA_MB=1" "RIR_MMA_MB=2" "RIR_MMA_MB=1 RIR_MMA_DEBUG=2" "RIR_MMA_MB=2 RIR_MMA_DEBUG=2"; do
  echo "--- nq=$q $e"; env $e timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-parity --nq $q 2> gpurun_out/s8.err | python -c "$fmt" || tail -5 gpurun_out/s8.err
done; done
