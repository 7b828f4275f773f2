#!/bin/bash
# round 2, call G: cooperative heavy scans, loop shapes, C++ class pool (1 GPU part)
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -x -q -k "nn or far or icp or memo or scan_sched" > gpurun_out/pytest_nn_g.log 2>&1; RC=$?; echo "pytest nn rc=$RC"; tail -4 gpurun_out/pytest_nn_g.log
run() { L=$1; shift
  env "$@" timeout 200 python scripts/bench_repo_clouds.py --no-baselines --reps 3 --only "W1 bunny res 0.005,W3 dragon mse,W4,W5" --skip "mse 1e-5" --out g_$L.json 2> gpurun_out/g_$L.err | sed "s/^/[$L] /" | cut -c1-175
}
run loop FGOICP_ICP_LOG=1
run nocoop FGOICP_NN_HEAVY_ROWS=0
run heavy1000 FGOICP_NN_HEAVY_ROWS=1000 FGOICP_ICP_LOG=1
run heavy6000 FGOICP_NN_HEAVY_ROWS=6000
run s384x3 FGOICP_ICP_SHAPE=384x3
run s640x2 FGOICP_ICP_SHAPE=640x2
run chain FGOICP_ICP_MODE=1
grep "icp loop" gpurun_out/g_loop.err | grep "jobs 1504\|jobs 8 slots 8 grid 296\|jobs 28 " | tail -4
grep "icp loop" gpurun_out/g_heavy1000.err | grep "jobs 1504\|jobs 8 slots 8 grid 296\|jobs 28 " | tail -3
timeout 600 python -m pytest tests/test_fullsize_parity.py tests/test_cpp_api_gpu.py -x -q -m gpu > gpurun_out/pytest_full_g.log 2>&1; echo "pytest fullsize+cpp rc=$?"; tail -5 gpurun_out/pytest_full_g.log
