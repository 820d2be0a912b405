#!/bin/bash
# even tile heights in range mode: parity + timing on shard emulations
mkdir -p gpurun_out
RIR_MMA_EVEN_TILES=1 timeout 900 python -m pytest -x -q -m gpu tests/test_gpu_search.py -k "not (stream or exact)" > gpurun_out/test_search_mma.log 2>&1; echo "search_mma rc=$?"; tail -3 gpurun_out/test_search_mma.log
fmt='import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]); r=d["roofline"]; e=d["e2e"]; print("nq=%d n=%d q/s=%.0f ms/step=%.4f (min %.4f) e2e=%.0f (%.4f) scan_ms=%.4f frac=%.3f parity=%s"%(d["config"]["nq"],d["config"]["n_db"],d["value"],d["ms_per_step"],d["step_ms"]["min"],e["value"],e["ms_per_step"],r["kernel_ms"],r["frac"],d.get("parity",{}).get("vs_exact")))'
for n in 125916 251831; do for e in "RIR_MMA_EVEN_TILES=1" "RIR_MMA_EVEN_TILES=0"; do for q in 70 1; do
  echo "--- n=$n $e nq=$q"; env $e timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extras --n-db $n --nq $q 2> gpurun_out/s8.err | python -c "$fmt" || tail -5 gpurun_out/s8.err
done; done; done
RIR_MMA_EVEN_TILES=1 timeout 120 python tools/timeline.py > gpurun_out/timeline_q70_shard8_even.txt 2>&1; tail -27 gpurun_out/timeline_q70_shard8_even.txt
