#!/bin/bash
# round 2, call J: ncu evidence of the tree as it stands -- launch lists of bench.py and run(), full captures of the in-search
# kernel (k_bnb_r3m: the 1504-cube leaf wave and a 128-cube wave), of the ICP loop kernel and of the flat roofline kernel
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 1 --no-bnb --no-cpu > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r02.csv \
    python bench.py --steps 2 --warmup 1 --no-bnb --no-cpu > gpurun_out/j_ncu_bench.log 2>&1
echo "bench launches rc=$?"
timeout 200 python scripts/profile_run.py > gpurun_out/j_run.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_run_r02b.csv \
    python scripts/profile_run.py > gpurun_out/j_ncu_run.log 2>&1
echo "run launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_bnb_r3m -s 11 -c 1 -f -o gpurun_out/ncu_bnb_leaf_r02 python scripts/profile_run.py > gpurun_out/j_ncu_bnb1.log 2>&1; echo "bnb leaf rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_bnb_r3m -s 5 -c 1 -f -o gpurun_out/ncu_bnb_w128_r02 python scripts/profile_run.py > gpurun_out/j_ncu_bnb2.log 2>&1; echo "bnb w128 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_icp_loop -s 2 -c 1 -f -o gpurun_out/ncu_icp_loop_r02 python scripts/profile_run.py > gpurun_out/j_ncu_icp.log 2>&1; echo "icp loop rc=$?"
timeout 200 python scripts/profile_phased.py > gpurun_out/j_phased.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_bounds_phased -s 2 -c 1 -f -o gpurun_out/ncu_phased_r02 python scripts/profile_phased.py > gpurun_out/j_ncu_phased.log 2>&1; echo "phased rc=$?"
ls -la gpurun_out/*.ncu-rep
