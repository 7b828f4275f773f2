#!/bin/bash
for k in 0 16 64 256; do
echo "=== wave1 $k"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --wave1 $k > gpurun_out/bench_w$k.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
python - <<PY
import json
b=json.load(open('gpurun_out/bench_w$k.json'))
bn=b['bnb']; print({k:bn[k] for k in bn if k!='levels'})
for l in bn['levels']: print({k:l[k] for k in ('span','cubes','icps','evals','best_sse','survivors','ms_ub','ms_icp','ms_lb')})
PY
done
