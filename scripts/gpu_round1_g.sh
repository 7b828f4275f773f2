#!/bin/bash
# round-1 evidence: GPU tests, smoke, both bench arms, ncu full capture of the roofline kernel, launch lists
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/bench.json'))
print({k:b[k] for k in ('value','ms_per_step','e2e','roofline','cpu_baseline','ctor_ms','gpu_launches','clocks')})
bn=b['bnb']; print({k:bn[k] for k in bn if k!='levels'})
r=json.load(open('gpurun_out/bench_ref.json')); print('reference arm', r['value'], r['cpu_baseline']['kind'])
PY
timeout 300 python scripts/profile_phased.py 4096 > gpurun_out/plain_phased.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bounds_phased -s 1 -c 1 -f -o gpurun_out/prof_phased_r01 python scripts/profile_phased.py 4096 > gpurun_out/ncu_phased.log 2>&1
tail -2 gpurun_out/ncu_phased.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-bnb > gpurun_out/bench_short.json 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_short.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-bnb > gpurun_out/ncu_launches.log 2>&1
wc -l gpurun_out/launches_bench_short.csv
bash scripts/gpu_launches_run.sh | tail -20
