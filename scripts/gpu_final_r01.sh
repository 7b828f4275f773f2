#!/bin/bash
# end-of-round evidence with the final code: bench line, reference arm, smoke, GPU test suite, one NN variant
mkdir -p gpurun_out
timeout 120 python bench.py > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err; echo "bench rc=$?"
timeout 60 python bench.py --impl reference > gpurun_out/bench_reference_arm_final.json 2> gpurun_out/bench_ref_final.err; echo "ref arm rc=$?"
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_final.log
timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log
tail -3 gpurun_out/pytest_gpu_final.log
timeout 45 python scripts/bench_repo_clouds.py --only "W1 bunny res 0.005" --skip "mse 1e-5" --cap 15 --reps 3 --out repo_clouds_bunny_baselines.json 2>&1 | tail -1
FGOICP_LIB=build/variants/lib_incr.so timeout 40 python scripts/bench_repo_clouds.py --no-baselines --reps 1 --only "W3 dragon mse" --out nn_incr_rows.json 2>&1 | tail -1
python - <<PY
import json
b=json.load(open('gpurun_out/bench_n1_final.json')); print({k:b[k] for k in ('value','ms_per_step')}, b['e2e']['value'], b['roofline']['frac'], b['bnb']['bnb_ms'], b['bnb']['bnb_ms_all_runs'])
PY
