#!/bin/bash
# Run the GPU parity suites as separate processes (a trapped kernel poisons its CUDA context) and keep the logs.
# usage (under gpurun): bash tools/gpu_check.sh [extra pytest args]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
rc=0
run() { name=$1; shift; echo "=== $name" ; timeout 900 python -m pytest -x -q -m gpu "$@" > gpurun_out/test_$name.log 2>&1; r=$?; tail -5 gpurun_out/test_$name.log; [ $r -ne 0 ] && rc=$r; }
run descriptor tests/test_gpu_descriptor.py
run evaluate tests/test_gpu_evaluate.py
run search_stream tests/test_gpu_search.py -k "stream or exact"
run search_mma tests/test_gpu_search.py -k "not (stream or exact)"
exit $rc
