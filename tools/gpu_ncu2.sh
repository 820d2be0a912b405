#!/bin/bash
# full ncu capture of the scan kernel for one configuration: gpu_ncu2.sh <name> <env...> -- <bench args>
name=$1; shift
envs=()
while [ "$1" != "--" ]; do envs+=("$1"); shift; done
shift
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline $*"
env "${envs[@]}" $B > gpurun_out/plain_$name.log 2>&1 &&
env "${envs[@]}" ncu --set full --clock-control none --import-source on -k regex:sim_mma_kernel -s 4 -c 1 -f -o gpurun_out/prof_$name $B > gpurun_out/ncu_$name.log 2>&1
ls -la gpurun_out/prof_$name.ncu-rep
