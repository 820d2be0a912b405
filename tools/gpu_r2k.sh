#!/bin/bash
# split epilogue (EH=2) for the two-block pair shape: parity + timing
mkdir -p gpurun_out
timeout 900 python -m pytest -x -q -m gpu tests/test_gpu_search.py -k "not (stream or exact)" > gpurun_out/test_search_mma.log 2>&1; echo "search_mma rc=$?"; tail -3 gpurun_out/test_search_mma.log
RIR_MMA_MB=2 timeout 900 python -m pytest -x -q -m gpu tests/test_gpu_search.py -k "not (stream or exact)" > gpurun_out/test_search_mma_mb2.log 2>&1; echo "search_mma MB=2 rc=$?"; tail -3 gpurun_out/test_search_mma_mb2.log
fmt='import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]); r=d["roofline"]; print("nq=%d q/s=%.0f ms/step=%.4f scan_ms=%.4f %s=%.1f frac=%.3f parity=%s"%(d["config"]["nq"],d["value"],d["ms_per_step"],r["kernel_ms"],r["unit"],r["achieved"],r["frac"],d.get("parity",{}).get("vs_exact")))'
for q in 1024 2048 4096; do for e in "RIR_MMA_MB=2 RIR_MMA_EH=1" "RIR_MMA_MB=2 RIR_MMA_EH=2" "RIR_MMA_MB=1"; do
  echo "--- nq=$q $e"; env $e timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --nq $q 2> gpurun_out/s8.err | python -c "$fmt" || tail -5 gpurun_out/s8.err
done; done
for e in "RIR_MMA_EH=1" "RIR_MMA_EH=2"; do echo "--- fp8 cfg-5 shard nq=4096 $e"; env $e timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --nq 4096 --n-db 197559 --k 10 --dtype fp8 2> gpurun_out/s8.err | python -c "$fmt" || tail -5 gpurun_out/s8.err; done
