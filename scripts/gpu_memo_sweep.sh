#!/bin/bash
# winner-memo margin sweep on the ICP-heavy repo clouds (one process per setting: the margin is read once)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
for m in 0 2.5e-4 1e-3 4e-3; do
  echo "== FGOICP_NN_MARGIN=$m"
  FGOICP_NN_MARGIN=$m timeout 120 python scripts/bench_repo_clouds.py --no-baselines --reps 1 --only "W3 dragon mse,W4,W5,W1 bunny res 0.005 mse" --out memo_sweep_$m.json 2>&1 | tail -4
done
