#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
for cs in auto 1; do
echo "=== cluster $cs"
if [ "$cs" = "auto" ]; then unset FGOICP_BNB_CLUSTER; else export FGOICP_BNB_CLUSTER=$cs; fi
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c$cs.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<PY
import json
b=json.load(open('gpurun_out/bench_c$cs.json'))
print({k:b[k] for k in ('value','ms_per_step')})
bn=b['bnb']; print({k:bn[k] for k in bn if k!='levels'})
for l in bn['levels']: print(l)
PY
done
