#!/bin/bash
# quick check: GPU parity tests + bound microbench (no CPU baseline, no run())
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu $BENCH_ARGS > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_q.err
python - <<PY
import json
b=json.load(open('gpurun_out/bench_q.json')); print({k:b[k] for k in ('value','ms_per_step')}, b['e2e']['value'], b['roofline']['frac'], b['roofline']['kernel_ms'])
if 'bnb' in b:
    bn=b['bnb']; print({k:bn[k] for k in bn if k!='levels'})
    for l in bn['levels']: print({k:l[k] for k in ('span','cubes','icps','evals','best_sse','survivors','ms_ub','ms_icp','ms_lb')})
PY
