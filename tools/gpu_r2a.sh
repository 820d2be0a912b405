#!/bin/bash
# Round-2 first 1-GPU call: parity suites + smoke, default bench (with parity block, extras, full-config CPU baseline),
# the reference arm, then the per-step cost of an 8-way shard (125,916 rows) with the launch features toggled.
mkdir -p gpurun_out
bash tools/gpu_check.sh; echo "gpu_check rc=$?"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_default.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm rc=$?"
fmt='import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]); r=d["roofline"]; print("nq=%d n=%d q/s=%.0f ms/step=%.4f (min %.4f med %.4f) e2e=%.0f (%.4f ms) scan_ms=%.4f frac=%.3f parity=%s"%(d["config"]["nq"],d["config"]["n_db"],d["value"],d["ms_per_step"],d["step_ms"]["min"],d["step_ms"]["median"],d["e2e"]["value"],d["e2e"]["ms_per_step"],r["kernel_ms"],r["frac"],d.get("parity",{}).get("vs_exact")))'
S8="python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extras --n-db 125916"
echo "--- shard8 default"; timeout 300 $S8 2> gpurun_out/s8.err | python -c "$fmt" || tail -5 gpurun_out/s8.err
echo "--- shard8 RIR_PDL=0"; RIR_PDL=0 timeout 300 $S8 2> gpurun_out/s8.err | python -c "$fmt" || tail -5 gpurun_out/s8.err
echo "--- shard8 no cooperative (mode 4)"; RIR_FUSED_LAUNCH_MODE=4 timeout 300 $S8 2> gpurun_out/s8.err | python -c "$fmt" || tail -5 gpurun_out/s8.err
echo "--- shard8 coop only (mode 2)"; RIR_FUSED_LAUNCH_MODE=2 timeout 300 $S8 2> gpurun_out/s8.err | python -c "$fmt" || tail -5 gpurun_out/s8.err
echo "--- shard8 select threads 256"; RIR_SELECT_THREADS=256 timeout 300 $S8 2> gpurun_out/s8.err | python -c "$fmt" || tail -5 gpurun_out/s8.err
echo "--- shard8 zero-copy queries"; RIR_HOST_ZERO_COPY=3 timeout 300 $S8 2> gpurun_out/s8.err | python -c "$fmt" || tail -5 gpurun_out/s8.err
echo "--- shard8 nq=1"; timeout 300 $S8 --nq 1 2> gpurun_out/s8.err | python -c "$fmt" || tail -5 gpurun_out/s8.err
echo "--- full default again, 200 steps"; timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extras 2> gpurun_out/s8.err | python -c "$fmt" || tail -5 gpurun_out/s8.err
timeout 120 python tools/timeline.py > gpurun_out/timeline_q70_shard8.txt 2>&1; tail -30 gpurun_out/timeline_q70_shard8.txt
B8="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-parity --n-db 125916"
$B8 > gpurun_out/plain_q70_shard8.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_q70_shard8.csv $B8 > gpurun_out/ncu_launch_q70_shard8.log 2>&1
echo "ncu rc=$?"
