#!/bin/bash
# round 2, call E: group-of-four miss dealing + cooperative long rows, cell-size sweep, loop vs chain; NVTX-renamed launch list
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -x -q -k "nn or far or icp or memo" > gpurun_out/pytest_nn_e.log 2>&1; RC=$?; echo "pytest nn rc=$RC"; tail -4 gpurun_out/pytest_nn_e.log
run() { # label, env...
  L=$1; shift
  env "$@" timeout 200 python scripts/bench_repo_clouds.py --no-baselines --reps 2 --only "W1 bunny res 0.005,W3 dragon mse,W4,W5" --skip "mse 1e-5" --out e_$L.json 2> gpurun_out/e_$L.err | sed "s/^/[$L] /" | cut -c1-175
}
run loop FGOICP_ICP_LOG=1
run chain FGOICP_ICP_MODE=1
run loop_c07 FGOICP_NN_CELL_SCALE=0.7
run chain_c07 FGOICP_ICP_MODE=1 FGOICP_NN_CELL_SCALE=0.7
run loop_c05 FGOICP_NN_CELL_SCALE=0.5
run chain_c05 FGOICP_ICP_MODE=1 FGOICP_NN_CELL_SCALE=0.5
run loop_c035 FGOICP_NN_CELL_SCALE=0.35
run chain_c035 FGOICP_ICP_MODE=1 FGOICP_NN_CELL_SCALE=0.35
run loop_c05_r128 FGOICP_NN_CELL_SCALE=0.5 FGOICP_NN_COARSE_MIN_ROWS=128
grep "icp loop" gpurun_out/e_loop.err | grep "jobs 1504\|jobs 8 slots 8 grid 296" | tail -3
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_e.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_e.log
timeout 300 ncu --nvtx --print-nvtx-rename kernel --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_nvtx_r02.csv python scripts/profile_run.py > gpurun_out/ncu_nvtx_e.log 2>&1; echo "ncu nvtx rc=$?"; head -c 600 gpurun_out/launches_nvtx_r02.csv | tail -c 300
