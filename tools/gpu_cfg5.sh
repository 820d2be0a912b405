#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 8 --steps 3 --warmup 2 --no-cpu-baseline --nq 100000 --n-db 1580470 --k 10 --dtype fp8 2> gpurun_out/cfg5.err | tee gpurun_out/cfg5_8gpu.json | cut -c1-400; tail -3 gpurun_out/cfg5.err
