#!/bin/bash
# last check of a round: parity suites, smoke, default bench, refreshed launch lists + the 70-query capture
mkdir -p gpurun_out
bash tools/gpu_check.sh 2>&1 | tail -12
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_q70.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_q70.csv $B > gpurun_out/ncu_launch_q70.log 2>&1
B8="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --n 125916"
$B8 > gpurun_out/plain_q70_shard8.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_q70_shard8.csv $B8 > gpurun_out/ncu_launch_q70_shard8.log 2>&1
$B > gpurun_out/plain_mma_q70.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sim_mma_kernel -s 4 -c 1 -f -o gpurun_out/prof_mma_q70 $B > gpurun_out/ncu_full_mma_q70.log 2>&1
ls -la gpurun_out/prof_mma_q70.ncu-rep
