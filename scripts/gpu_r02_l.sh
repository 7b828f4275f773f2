#!/bin/bash
# round 2, call L: one-thread scans of small balls inside the ICP loop kernel (fg_nn_near): parity, scan counts, times
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_golden_clouds.py tests/test_golden.py -x -q -m gpu 2>&1 | tail -3
FGOICP_ICP_MODE=2 timeout 900 python -m pytest tests/test_fullsize_parity.py -x -q -m gpu 2>&1 | tail -3
for near in 1 0; do
  for c in "w5 0.005 1e-4" "dragon 0.005 1e-4" "overlap 0.005 1e-4" "bunny 0.005 1e-3"; do
    set -- $c
    FGOICP_NN_NEAR=$near FGOICP_ICP_MODE=2 FGOICP_ICP_LOG=1 python scripts/run_repo_case.py $1 $2 $3 2 2> gpurun_out/l_$1_$near.err | tail -1 | cut -c1-120
    python - <<PY
import re
a=b=t=na=nb=0
st=[0.0]*9
for l in open("gpurun_out/l_$1_$near.err"):
    m=re.search(r"trips (\d+) full scans: rooted (\d+) squared (\d+) \(heavy \d+\) one-thread scans: rooted (\d+) squared (\d+) \| stage us: (.*)", l)
    if m:
        t+=int(m[1]); a+=int(m[2]); b+=int(m[3]); na+=int(m[4]); nb+=int(m[5])
        for k,v in enumerate(m[6].split()[:9]): st[k]+=float(v)
print("   [$1 near=$near, both runs] trips %d | warp scans rooted %.3e squared %.3e | one-thread rooted %.3e squared %.3e | stage ms %s" % (t,a,b,na,nb," ".join("%.1f"%(x/1e3) for x in st)))
PY
  done
done
python scripts/run_repo_case.py dragon 0.005 1e-4 1 | tail -1 | cut -c1-120
