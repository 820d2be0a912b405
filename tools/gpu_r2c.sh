#!/bin/bash
# descriptor head + covariance on the tensor cores; ncu environment probe; e2e at full size with launch modes
mkdir -p gpurun_out
timeout 900 python -m pytest -x -q -m gpu tests/test_gpu_descriptor.py > gpurun_out/test_descriptor.log 2>&1; echo "descriptor rc=$?"; tail -15 gpurun_out/test_descriptor.log
timeout 300 python tools/bench_descriptor.py > gpurun_out/bench_descriptor.jsonl 2> gpurun_out/bench_descriptor.err; echo "bench_descriptor rc=$?"; cat gpurun_out/bench_descriptor.jsonl; tail -3 gpurun_out/bench_descriptor.err
timeout 300 python tools/bench_descriptor.py --pca 2>&1 | tail -2
RIR_PCA_FP32=1 timeout 300 python tools/bench_descriptor.py --pca 2>&1 | tail -2
echo "=== ncu env probe"
ncu --metrics gpu__time_duration.sum -c 1 python -c "
import os
print({k: v for k, v in os.environ.items() if any(t in k.upper() for t in ('NV', 'CUDA', 'INJECT', 'NSIGHT', 'PROF', 'LD_PRELOAD'))})
print([l.split()[-1] for l in open('/proc/self/maps') if any(t in l.lower() for t in ('nsight', 'inject', 'nvperf', 'ncu', 'cupti'))][:10])
" 2>&1 | tail -8
for e in "RIR_X=1" "RIR_PDL=0" "RIR_FUSED_LAUNCH_MODE=4" "RIR_FUSED_LAUNCH_MODE=2"; do echo "--- e2e probe full $e"; env $e python tools/e2e_probe.py --n 1007323 --steps 100; done
echo "--- e2e probe full PDL=0 mode 4"; RIR_PDL=0 RIR_FUSED_LAUNCH_MODE=4 python tools/e2e_probe.py --n 1007323 --steps 100
