#!/bin/bash
timeout 900 python -m pytest -x -q -m gpu tests/test_gpu_search.py > gpurun_out/test_search_all.log 2>&1; tail -3 gpurun_out/test_search_all.log
fmt='import sys,json
d=json.loads(sys.stdin.read()); r=d["roofline"]; print("nq=%d n=%d q/s=%.0f ms/step=%.3f e2e=%.0f scan_ms=%.3f frac=%.3f"%(d["config"]["nq"],d["config"]["n_db"],d["value"],d["ms_per_step"],d["e2e"]["value"],r["kernel_ms"],r["frac"]))'
run() { echo -n "$1 :: "; env $1 timeout 300 python bench.py --steps 100 --no-cpu-baseline "${@:2}" 2>gpurun_out/sweep.err | tail -1 | python -c "$fmt" || tail -3 gpurun_out/sweep.err; }
for n in 125916 251831 503662; do for q in 70 1; do
run "RIR_MMA_TILE128=0" --n $n --nq $q
run "RIR_MMA_TILE128=1" --n $n --nq $q
done; done
