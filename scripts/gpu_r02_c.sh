#!/bin/bash
# round 2, call C: far-query coarse boxes (tests, sweep), loop vs chain, parity traces, ncu captures
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -x -q -k "nn or far or icp or memo" > gpurun_out/pytest_nn_c.log 2>&1; RC=$?; echo "pytest nn rc=$RC"; tail -4 gpurun_out/pytest_nn_c.log
if [ $RC -ne 0 ]; then export FGOICP_NN_COARSE=0; echo "coarse path FAILED its tests: continuing with it off"; fi
run() { # label, env...
  L=$1; shift
  env "$@" timeout 200 python scripts/bench_repo_clouds.py --no-baselines --reps 2 --only "W1 bunny res 0.005,W3 dragon mse,W4,W5" --skip "mse 1e-5" --out c_$L.json 2> gpurun_out/c_$L.err | sed "s/^/[$L] /" | cut -c1-200
}
run default FGOICP_ICP_LOG=1
run chain FGOICP_ICP_MODE=1
run nocoarse FGOICP_NN_COARSE=0
run rows64 FGOICP_NN_COARSE_MIN_ROWS=64
run rows160 FGOICP_NN_COARSE_MIN_ROWS=160
run rows640 FGOICP_NN_COARSE_MIN_ROWS=640
grep "icp loop" gpurun_out/c_default.err | grep "jobs 1504\|jobs 8 slots 8 grid 296" | tail -4
timeout 300 python scripts/run_parity_r02.py --only "W4,W3" --mse 1e-2 --trace --cap 120 --out run_parity_traces_r02.json > gpurun_out/run_parity_c.log 2>&1; echo "parity traces rc=$?"; tail -4 gpurun_out/run_parity_c.log | cut -c1-400
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_c.log
# ---- ncu: the kernels run() spends its time in (each command already exited 0 above without ncu)
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 400 $NCU -k regex:k_bnb_r3m -s 11 -c 1 -o gpurun_out/ncu_bnb_r02 python scripts/profile_run.py > gpurun_out/ncu_bnb_c.log 2>&1; echo "ncu bnb rc=$?"
timeout 400 $NCU -k regex:k_icp_loop -s 3 -c 1 -o gpurun_out/ncu_icp_loop_w5_r02 python scripts/profile_run.py > gpurun_out/ncu_icp_c.log 2>&1; echo "ncu icp loop rc=$?"
FGOICP_ICP_MODE=1 timeout 600 $NCU -k regex:k_nn_grid -s 1500 -c 2 -o gpurun_out/ncu_nn_grid_w3_r02 python scripts/profile_nn_w3.py > gpurun_out/ncu_nn_c.log 2>&1; echo "ncu nn rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_run_r02.csv python scripts/profile_run.py > gpurun_out/ncu_launches_c.log 2>&1; echo "ncu launches rc=$?"
ls -la gpurun_out/*.ncu-rep
