#!/bin/bash
# final single-GPU pass: suites, smoke, default bench (+ reference arm), shard emulations, then the ncu evidence
mkdir -p gpurun_out
bash tools/gpu_check.sh; echo "gpu_check rc=$?"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_default.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm rc=$?"
fmt='import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]); r=d["roofline"]; e=d["e2e"]; print("nq=%d n=%d q/s=%.0f ms/step=%.4f (min %.4f) e2e=%.0f (%.4f) sync=%s scan_ms=%.4f frac=%.3f parity=%s"%(d["config"]["nq"],d["config"]["n_db"],d["value"],d["ms_per_step"],d["step_ms"]["min"],e["value"],e["ms_per_step"],(e.get("sync_each_step") or {}).get("ms_per_step"),r["kernel_ms"],r["frac"],d.get("parity",{}).get("vs_exact")))'
python -c "$fmt" < gpurun_out/bench_default.json
for n in 125916 251831 503662 1007323; do
  echo "--- 1 GPU n=$n"; timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extras --n-db $n 2> gpurun_out/s8.err | python -c "$fmt" || tail -5 gpurun_out/s8.err
done
timeout 120 python tools/timeline.py > gpurun_out/timeline_q70_shard8.txt 2>&1
timeout 120 python tools/timeline.py --nq 1 > gpurun_out/timeline_q1_shard8.txt 2>&1
timeout 300 python tools/bench_descriptor.py > gpurun_out/bench_descriptor.jsonl 2>/dev/null; timeout 300 python tools/bench_descriptor.py --pca >> gpurun_out/bench_descriptor.jsonl 2>/dev/null; RIR_PCA_FP32=1 timeout 300 python tools/bench_descriptor.py --pca >> gpurun_out/bench_descriptor.jsonl 2>/dev/null
timeout 300 python tools/bench_cfg1.py > gpurun_out/bench_cfg1.json 2>/dev/null; echo "cfg1 rc=$?"
bash tools/gpu_profiles.sh 2>&1 | tail -16
