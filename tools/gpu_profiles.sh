#!/bin/bash
# Evidence for profiles/: launch lists + full ncu captures (each ncu run follows the identical plain command having
# exited 0).  Outputs under gpurun_out/; summarise with tools/ncu_summary.py.
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_q70.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_q70.csv $B > gpurun_out/ncu_launch_q70.log 2>&1
B8="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --n 125916"
$B8 > gpurun_out/plain_q70_shard8.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_q70_shard8.csv $B8 > gpurun_out/ncu_launch_q70_shard8.log 2>&1
cap() { name=$1; shift; kern=$1; shift; C="python bench.py --steps 3 --warmup 3 --no-cpu-baseline $*"
  $C > gpurun_out/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$kern -s 4 -c 1 -f -o gpurun_out/prof_$name $C > gpurun_out/ncu_full_$name.log 2>&1; }
cap mma_q70 sim_mma_kernel
cap mma_q1024 sim_mma_kernel --nq 1024
cap mma_q1 sim_mma_kernel --nq 1
cap mma_q4096 sim_mma_kernel --nq 4096
cap select_q70 final_select_kernel
P="python tools/bench_descriptor.py"
$P > gpurun_out/plain_pool.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pool_kernel -s 3 -c 1 -f -o gpurun_out/prof_pool_cfg4 $P > gpurun_out/ncu_full_pool.log 2>&1
ls -la gpurun_out/*.ncu-rep
