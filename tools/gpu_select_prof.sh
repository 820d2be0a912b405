#!/bin/bash
timeout 300 python -m pytest -x -q -m gpu tests/test_gpu_search.py -k "clustered" 2>&1 | tail -3
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_sel.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:final_select_kernel -s 4 -c 1 -f -o gpurun_out/prof_select_q70 $B > gpurun_out/ncu_select.log 2>&1
ls -la gpurun_out/prof_select_q70.ncu-rep
