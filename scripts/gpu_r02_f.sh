#!/bin/bash
# round 2, call F: SVD convergence at unit roundoff, S4+S5 / S8+S9 merged: tests, loop vs chain
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_f.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_f.log
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_trimming.py tests/test_golden.py tests/test_golden_clouds.py -x -q -m gpu > gpurun_out/pytest_f.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_f.log
run() { L=$1; shift
  env "$@" timeout 200 python scripts/bench_repo_clouds.py --no-baselines --reps 3 --only "W1 bunny res 0.005,W3 dragon mse,W4,W5" --skip "mse 1e-5" --out f_$L.json 2> gpurun_out/f_$L.err | sed "s/^/[$L] /" | cut -c1-175
}
run loop FGOICP_ICP_LOG=1
run chain FGOICP_ICP_MODE=1
grep "icp loop" gpurun_out/f_loop.err | grep "jobs 1504\|jobs 8 slots 8 grid 296\|jobs 28 " | tail -4
timeout 600 python -m pytest tests/test_fullsize_parity.py -x -q -m gpu > gpurun_out/pytest_full_f.log 2>&1; echo "pytest fullsize rc=$?"; tail -5 gpurun_out/pytest_full_f.log
