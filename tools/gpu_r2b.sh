#!/bin/bash
# Which launch attribute does ncu refuse?  + where the e2e step goes
mkdir -p gpurun_out
probe() { name=$1; shift; echo "=== ncu probe: $name"; env "$@" python tools/mini_search.py > gpurun_out/mini_$name.log 2>&1 && env "$@" ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/mini_$name.csv python tools/mini_search.py > gpurun_out/mini_ncu_$name.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/mini_ncu_$name.log; grep -c "rir::" gpurun_out/mini_$name.csv; }
probe nocoop_nopdl RIR_FUSED_LAUNCH_MODE=4 RIR_PDL=0
probe coop_nopdl RIR_PDL=0
probe nocoop_pdl RIR_FUSED_LAUNCH_MODE=4
probe default RIR_X=1
for e in "RIR_X=1" "RIR_PDL=0" "RIR_FUSED_LAUNCH_MODE=4" "RIR_HOST_ZERO_COPY=0" "RIR_HOST_ZERO_COPY=3"; do echo "--- e2e probe shard8 $e"; env $e python tools/e2e_probe.py; done
echo "--- e2e probe full"; python tools/e2e_probe.py --n 1007323 --steps 100
