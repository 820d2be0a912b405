#!/bin/bash
# Evidence for profiles/: launch lists + full ncu captures (each ncu run follows the identical plain command having
# exited 0).  Outputs under gpurun_out/; summarise with tools/ncu_summary.py and tools/make_traffic.py.
# (Under ncu the fused scan drops its cooperative attribute — Nsight Compute cannot replay it)
mkdir -p gpurun_out
F="--steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-parity"
B="python bench.py $F"
$B > gpurun_out/plain_q70.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_q70.csv $B > gpurun_out/ncu_launch_q70.log 2>&1
B8="python bench.py $F --n-db 125916"
$B8 > gpurun_out/plain_q70_shard8.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_q70_shard8.csv $B8 > gpurun_out/ncu_launch_q70_shard8.log 2>&1
cap() { name=$1; shift; kern=$1; shift; skip=$1; shift; C="$*"
  $C > gpurun_out/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c 1 -f -o gpurun_out/prof_$name $C > gpurun_out/ncu_full_$name.log 2>&1; echo "cap $name rc=$?"; }
cap mma_q70 sim_mma_kernel 4 $B
cap mma_q70_shard8 sim_mma_kernel 4 $B8
cap mma_q1024 sim_mma_kernel 4 $B --nq 1024
cap mma_q1 sim_mma_kernel 4 $B --nq 1
cap select_q70 final_select_kernel 4 $B
cap select_q70_shard8 final_select_kernel 4 $B8
P="python tools/bench_descriptor.py"
$P > gpurun_out/plain_pool.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_head.csv $P > gpurun_out/ncu_launch_head.log 2>&1
cap pool_cfg4 pool_kernel 3 $P
cap dense_cfg4 dense_nt_kernel 3 $P
cap finish_cfg4 whiten_finish_kernel 3 $P
cap dense_pca dense_nt_kernel 2 $P --pca
ls -la gpurun_out/*.ncu-rep
