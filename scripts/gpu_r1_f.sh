#!/bin/bash
# state check of HEAD: GPU parity tests, bench (both arms), run() wave sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_f.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_f.json 2> gpurun_out/bench_ref_f.err; echo "ref rc=$?"
for k in 32 128 256 512; do
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --wave1 $k > gpurun_out/bench_fw$k.json 2>> gpurun_out/bench_f.err
python - <<PY
import json
b=json.load(open('gpurun_out/bench_fw$k.json'))
bn=b['bnb']; print('wave1', $k, {k:bn[k] for k in bn if k!='levels'})
for l in bn['levels']: print({k:l[k] for k in ('span','cubes','icps','evals','best_sse','survivors','ms_ub','ms_icp','ms_lb')})
PY
done
python - <<PY
import json
b=json.load(open('gpurun_out/bench_f.json')); print({k:b[k] for k in ('value','ms_per_step','e2e','roofline','cpu_baseline','clocks')}); print(b['bnb']['bnb_ms'])
print(open('gpurun_out/bench_ref_f.json').read()[:600])
PY
