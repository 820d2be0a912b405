#!/bin/bash
# Evidence for profiles/: launch lists + full ncu captures of the scan kernel (each ncu run follows the identical
# plain command having exited 0).
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_q70.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_q70.csv $B > gpurun_out/ncu_launch_q70.log 2>&1
B8="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --n 125916"
$B8 > gpurun_out/plain_q70_shard8.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_q70_shard8.csv $B8 > gpurun_out/ncu_launch_q70_shard8.log 2>&1
$B > gpurun_out/plain_q70b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sim_mma_kernel -s 4 -c 1 -f -o gpurun_out/prof_mma_q70 $B > gpurun_out/ncu_full_q70.log 2>&1
B2="python bench.py --nq 1024 --steps 3 --warmup 3 --no-cpu-baseline"
$B2 > gpurun_out/plain_q1024.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sim_mma_kernel -s 4 -c 1 -f -o gpurun_out/prof_mma_q1024 $B2 > gpurun_out/ncu_full_q1024.log 2>&1
B1="python bench.py --nq 1 --steps 3 --warmup 3 --no-cpu-baseline"
$B1 > gpurun_out/plain_q1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sim_mma_kernel -s 4 -c 1 -f -o gpurun_out/prof_mma_q1 $B1 > gpurun_out/ncu_full_q1.log 2>&1
ls -la gpurun_out/*.ncu-rep
