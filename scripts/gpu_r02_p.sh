#!/bin/bash
# round 2, call P: hunt the sporadic host-side stall of the round-synchronous (trimmed) searches: two processes, one GPU each
mkdir -p gpurun_out
for g in 0 1; do
  CUDA_VISIBLE_DEVICES=$g FGOICP_BNBR_LOG=1 python scripts/run_repo_case.py skull 0.005 1e-3 14 0.1 > gpurun_out/p_$g.log 2> gpurun_out/p_$g.err &
done
wait
for g in 0 1; do
  cut -c1-90 gpurun_out/p_$g.log | awk '{print $9}' | tr '\n' ' '; echo
  python - <<PY
import re
L=open("gpurun_out/p_$g.err").read().splitlines()
for i,l in enumerate(L):
    m=re.search(r"\+([\d.]+) us", l)
    if m and float(m[1])>20000:
        print("rank $g line", i); print("\n".join(L[max(0,i-2):i+2]))
PY
done
