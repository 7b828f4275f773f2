#!/bin/bash
# round 2, call K: how many full scans of the refinements are memo failures on near-ties -- clearance slack variants
for v in "" slack2e6 slack5e7; do
  L=""; [ -n "$v" ] && L="FGOICP_LIB=/root/repo/build/variants/lib_$v.so"
  for pair in dragon overlap; do
    env $L FGOICP_ICP_MODE=2 FGOICP_ICP_LOG=1 python scripts/run_repo_case.py $pair 0.005 1e-4 1 2> gpurun_out/k_$pair$v.err | tail -1 | cut -c1-150
    python - <<PY
import re
a=b=t=0
for l in open("gpurun_out/k_$pair$v.err"):
    m=re.search(r"trips (\d+) full scans: rooted (\d+) squared (\d+)", l)
    if m: t+=int(m[1]); a+=int(m[2]); b+=int(m[3])
print("   [$pair ${v:-default}] loop-kernel trips %d, full scans rooted %.3e squared %.3e" % (t,a,b))
PY
    env $L python scripts/run_repo_case.py $pair 0.005 1e-4 1 | tail -1 | cut -c1-110
  done
done
