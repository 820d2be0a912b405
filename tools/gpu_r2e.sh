#!/bin/bash
# range-mode scan: parity (search suites) + shard sizes 1/8, 1/4, 1/2, 1/1 with range mode on/off and sample-tile choice
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 900 python -m pytest -x -q -m gpu "$@" > gpurun_out/test_$name.log 2>&1; r=$?; tail -4 gpurun_out/test_$name.log; }
run search_mma tests/test_gpu_search.py -k "not (stream or exact)"
run descriptor tests/test_gpu_descriptor.py
fmt='import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]); r=d["roofline"]; print("nq=%d n=%d q/s=%.0f ms/step=%.4f (min %.4f med %.4f) scan_ms=%.4f frac=%.3f parity=%s"%(d["config"]["nq"],d["config"]["n_db"],d["value"],d["ms_per_step"],d["step_ms"]["min"],d["step_ms"]["median"],r["kernel_ms"],r["frac"],d.get("parity",{}).get("vs_exact")))'
for n in 125916 251831 503662 1007323; do
  S="python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extras --n-db $n"
  for e in "RIR_MMA_RANGE=0" "RIR_MMA_RANGE=1" "RIR_MMA_SAMPLE_TAIL=0" "RIR_MMA_SAMPLE_TAIL=1"; do
    echo "--- n=$n $e"; env $e timeout 300 $S 2> gpurun_out/s8.err | python -c "$fmt" || tail -5 gpurun_out/s8.err
  done
done
echo "--- nq=1 shard8 range"; timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extras --n-db 125916 --nq 1 2> gpurun_out/s8.err | python -c "$fmt" || tail -5 gpurun_out/s8.err
echo "--- nq=1 shard8 classic"; RIR_MMA_RANGE=0 timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extras --n-db 125916 --nq 1 2> gpurun_out/s8.err | python -c "$fmt" || tail -5 gpurun_out/s8.err
timeout 120 python tools/timeline.py > gpurun_out/timeline_q70_shard8_range.txt 2>&1; tail -32 gpurun_out/timeline_q70_shard8_range.txt
timeout 300 python tools/bench_descriptor.py 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['workload'][:60], 'pool', round(d['pool_p3_ms'],4), 'head', round(d['head_total_ms'],4), 'frac', round(d['head_frac_of_measured_hbm'],3), 'whiten_tc', round(d['whiten_l2_tensor_core_ms'],4))"
timeout 300 python tools/bench_descriptor.py --pca 2>&1 | tail -1
