#!/bin/bash
# ncu --set full capture of the phased bound kernel at bench geometry (after a plain run exits 0)
mkdir -p gpurun_out
timeout 200 python scripts/profile_phased.py 4096 > gpurun_out/plain_phased.log 2>&1 || { echo "plain run failed"; tail gpurun_out/plain_phased.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bounds_phased -s 1 -c 1 -f -o gpurun_out/prof_phased_r01b python scripts/profile_phased.py 4096 > gpurun_out/ncu_phased.log 2>&1
tail -3 gpurun_out/ncu_phased.log
