#!/bin/bash
mkdir -p gpurun_out
for g in 32 64 128; do
echo "=== FGOICP_L2_FETCH=$g"
FGOICP_L2_FETCH=$g python - <<'PY'
import json, sys
sys.path.insert(0,'.')
import numpy as np
from fast_go_icp_b200 import capi
pts=np.random.default_rng(0).uniform(-1,1,(64,3)).astype(np.float32)
ctx=capi.Context(pts,pts,pts.min(0),pts.max(0),0.2,flags=0)
for nbytes in (200<<20, 1200<<20):
    print(nbytes>>20, "MB", {w: round(ctx.gather_probe(nbytes,w,8)) for w in (16,32,64,128)})
PY
FGOICP_L2_FETCH=$g python scripts/sampler_sweep.py gpurun_out/sampler_sweep_l2f$g.json 2>&1 | grep -v "true" 
done
