#!/bin/bash
# N-GPU call (N=2 by default; `N=8 bash tools/gpu_multi.sh` on an 8-GPU box): the peer-exchange parity tests for every
# world size the box can host, then the bench (with its parity block) at each N, both exchange flavours.
mkdir -p gpurun_out
N=${N:-2}
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 1200 python -m pytest -x -q -m gpu tests/test_gpu_multi.py > gpurun_out/test_multi.log 2>&1; echo "test_multi rc=$?"; tail -6 gpurun_out/test_multi.log
fmt='import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]); r=d["roofline"]; e=d["e2e"]; p=d.get("parity") or {}; x=d.get("extras") or {}
print("gpus=%d nq=%d q/s=%.0f ms/step=%.4f (min %.4f) e2e=%.0f pipelined=%s scan_ms=%.4f frac=%.3f parity=%s/%s q1=%s q1024=%s %s"%(d["n_gpus"],d["config"]["nq"],d["value"],d["ms_per_step"],d["step_ms"]["min"],e["value"],(e.get("pipelined") or {}).get("value"),r["kernel_ms"],r["frac"],p.get("peer_eq_nccl"),p.get("vs_exact"),(x.get("q1") or {}).get("value"),(x.get("q1024") or {}).get("value"),d["impl_detail"]["exchange"][:12]))'
for n in 2 4 8; do
  [ $n -gt $N ] && break
  for ex in peer nccl; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --steps 100 --warmup 5 --exchange $ex 2> gpurun_out/bench_multi_$n.err | tee gpurun_out/bench_n${n}_$ex.json | python -c "$fmt" || tail -5 gpurun_out/bench_multi_$n.err
  done
done
