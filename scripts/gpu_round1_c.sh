#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python scripts/run_cli_demo.py gpurun_out/cli_demo > gpurun_out/cli_demo.log 2>&1; echo "cli rc=$?"; tail -25 gpurun_out/cli_demo.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/bench.json'))
print({k:b[k] for k in ('value','ms_per_step','e2e','roofline','cpu_baseline','ctor_ms')})
bn=b['bnb']; print({k:bn[k] for k in bn if k!='levels'})
for l in bn['levels']: print(l)
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -3 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
