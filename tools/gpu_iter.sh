#!/bin/bash
# quick iteration: search parity + headline bench lines + launch list
mkdir -p gpurun_out
timeout 900 python -m pytest -x -q -m gpu tests/test_gpu_search.py > gpurun_out/test_search_all.log 2>&1; tail -3 gpurun_out/test_search_all.log
fmt='import sys,json
d=json.loads(sys.stdin.read()); r=d["roofline"]; print("nq=%d q/s=%.0f ms/step=%.3f e2e=%.0f scan_ms=%.3f hbm_frac=%.3f tflops=%.0f"%(d["config"]["nq"],d["value"],d["ms_per_step"],d["e2e"]["value"],r["kernel_ms"],r["frac"], 2*d["config"]["nq"]*d["config"]["n_db"]*d["config"]["dim"]/r["kernel_ms"]/1e9))'
for q in 1 4 70 1024; do python bench.py --nq $q --steps 50 --no-cpu-baseline 2>gpurun_out/bench_iter_$q.err | tail -1 | tee gpurun_out/bench_iter_$q.json | python -c "$fmt"; done
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_q70.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_q70.csv $B > gpurun_out/ncu_launch_q70.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_q70.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r[0]=='ID'][0]
h=rows[hdr]; data=rows[hdr+1:]
ki=h.index('Kernel Name'); vi=h.index('Metric Value')
for r in data[-8:]: print(f"{r[ki][:50]:52s} {float(r[vi].replace(',',''))/1e3:9.1f} us")
PY
