#!/bin/bash
# rows-per-lane sweep of the cell-grid NN search (variants built by scripts/build_variant.py rplN nn_icp.cu "-DNN_RPL=N")
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_golden_clouds.py tests/test_gpu_properties.py -m gpu -x -q -k "icp_batch or icp_vs or nn or cuda_vs or full_run" > gpurun_out/pytest_nn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_nn.log
tail -4 gpurun_out/pytest_nn.log
for v in default rpl1 rpl8; do
  echo "== $v"
  lib=""; [ $v != default ] && lib=build/variants/lib_$v.so
  FGOICP_LIB=$lib timeout 100 python scripts/bench_repo_clouds.py --no-baselines --reps 1 --only "W3 dragon mse,W5" --out rpl_sweep_$v.json 2>&1 | tail -2
done
