#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/profile_run.py > gpurun_out/plain_run.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_bnb_r3 -s 6 -c 1 -f -o gpurun_out/prof_bnb python scripts/profile_run.py > gpurun_out/ncu_bnb.log 2>&1
tail -3 gpurun_out/ncu_bnb.log
