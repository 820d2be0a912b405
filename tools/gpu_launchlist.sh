#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_q70.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_q70.csv $B > gpurun_out/ncu_launch_q70.log 2>&1
for r in "2,6" "3,5" "4,4" "2,5" "1,6"; do echo "rings $r"; RIR_MMA_RINGS=$r python bench.py --steps 50 --no-cpu-baseline | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('q/s=%.0f ms/step=%.3f scan_ms=%.3f frac=%.3f'%(d['value'],d['ms_per_step'],r['kernel_ms'],r['frac']))"; done
for q in 8 16 32 128; do echo "nq $q"; python bench.py --nq $q --steps 50 --no-cpu-baseline | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('q/s=%.0f ms/step=%.3f scan_ms=%.3f frac=%.3f'%(d['value'],d['ms_per_step'],r['kernel_ms'],r['frac']))"; done
