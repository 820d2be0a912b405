#!/bin/bash
# Headline bench lines (1 GPU): BASELINE cfg-2 at 70 queries, plus the 1-query and 1k-query operating points.
mkdir -p gpurun_out
python bench.py --steps ${STEPS:-100} --warmup 5 > gpurun_out/bench_q70.json 2> gpurun_out/bench_q70.err; echo "q70 rc=$?"; cat gpurun_out/bench_q70.json
python bench.py --nq 1 --steps ${STEPS:-100} --warmup 5 --no-cpu-baseline > gpurun_out/bench_q1.json 2> gpurun_out/bench_q1.err; echo "q1 rc=$?"; cat gpurun_out/bench_q1.json
python bench.py --nq 1 --path mma --steps ${STEPS:-100} --warmup 5 --no-cpu-baseline > gpurun_out/bench_q1_mma.json 2> gpurun_out/bench_q1_mma.err; echo "q1mma rc=$?"; cat gpurun_out/bench_q1_mma.json
python bench.py --nq 1024 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_q1024.json 2> gpurun_out/bench_q1024.err; echo "q1024 rc=$?"; cat gpurun_out/bench_q1024.json
tail -3 gpurun_out/bench_*.err
