#!/bin/bash
# 2-GPU call: peer-exchange parity test, then the bench with both exchange flavours
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m pytest -x -q -m gpu tests/test_gpu_multi.py > gpurun_out/test_multi.log 2>&1; tail -15 gpurun_out/test_multi.log
N=${N:-2}
fmt='import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]); r=d["roofline"]; print("gpus=%d nq=%d q/s=%.0f ms/step=%.3f e2e=%.0f scan_ms=%.3f frac=%.3f  %s"%(d["n_gpus"],d["config"]["nq"],d["value"],d["ms_per_step"],d["e2e"]["value"],r["kernel_ms"],r["frac"],d["config"]["exchange"][:20]))'
for ex in peer nccl; do for q in 70 1 1024; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $N --steps 50 --warmup 5 --nq $q --exchange $ex --no-cpu-baseline 2> gpurun_out/bench_multi.err | python -c "$fmt" || tail -5 gpurun_out/bench_multi.err
done; done
