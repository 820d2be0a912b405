#!/bin/bash
# parity suites (incl. the count-based full-ranking mAP), fp8 peak, ncu launch lists (head, default bench, shard8)
mkdir -p gpurun_out
bash tools/gpu_check.sh; echo "gpu_check rc=$?"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 300 python tools/measure_fp8_peak.py --write 2>&1 | tail -2
timeout 300 python tools/bench_descriptor.py > gpurun_out/bench_descriptor.jsonl 2> gpurun_out/bench_descriptor.err; cat gpurun_out/bench_descriptor.jsonl | cut -c1-900
timeout 300 python tools/bench_descriptor.py --pca 2>&1 | tail -1
P="python tools/bench_descriptor.py"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_head.csv $P > gpurun_out/ncu_launch_head.log 2>&1; echo "ncu head rc=$?"
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-parity"
$B > gpurun_out/plain_q70.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_q70.csv $B > gpurun_out/ncu_launch_q70.log 2>&1; echo "ncu q70 rc=$?"
$B --n-db 125916 > gpurun_out/plain_q70_shard8.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_q70_shard8.csv $B --n-db 125916 > gpurun_out/ncu_launch_q70_shard8.log 2>&1; echo "ncu shard8 rc=$?"
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_q70.json 2> gpurun_out/bench_q70.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_q70.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['pipelined'], d['parity']['vs_exact'])"
