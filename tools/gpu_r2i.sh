#!/bin/bash
# HEAD validation: all suites, smoke, the driver's bench commands
mkdir -p gpurun_out
bash tools/gpu_check.sh; echo "gpu_check rc=$?"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_default.err
fmt='import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]); r=d["roofline"]; e=d["e2e"]; x=d.get("extras") or {}; print("nq=%d n=%d q/s=%.0f ms/step=%.4f (min %.4f) e2e=%.0f (%.4f) sync=%s scan_ms=%.4f frac=%.3f parity=%s q1=%s q1024=%s"%(d["config"]["nq"],d["config"]["n_db"],d["value"],d["ms_per_step"],d["step_ms"]["min"],e["value"],e["ms_per_step"],(e.get("sync_each_step") or {}).get("ms_per_step"),r["kernel_ms"],r["frac"],d.get("parity",{}).get("vs_exact"),x.get("q1"),x.get("q1024")))'
python -c "$fmt" < gpurun_out/bench_default.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_smoke.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ncu_smoke.log 2>&1; echo "ncu smoke rc=$?"; grep -c "rir::" gpurun_out/launches_smoke.csv
