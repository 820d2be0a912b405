#!/bin/bash
# One 1-GPU call: GPU parity suites, smoke, the default bench + the reference arm, then the ncu evidence for profiles/
# (tools/gpu_profiles.sh: every ncu command runs only after the identical plain command exited 0).
mkdir -p gpurun_out
bash tools/gpu_check.sh
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; echo "reference arm rc=$?"
bash tools/gpu_profiles.sh 2>&1 | tail -8
