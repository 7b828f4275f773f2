#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python scripts/tex_conformance.py gpurun_out/tex_conformance.json > gpurun_out/tex_conformance.log 2>&1
python - <<'PY'
import json
r=json.load(open('gpurun_out/tex_conformance.json'))
for k,v in r['candidates'].items(): print(k, v)
PY
python - <<'PY' > gpurun_out/gather_probe.json
import json, sys
sys.path.insert(0,'.')
import numpy as np
from fast_go_icp_b200 import capi
pts=np.random.default_rng(0).uniform(-1,1,(64,3)).astype(np.float32)
ctx=capi.Context(pts,pts,pts.min(0),pts.max(0),0.2,flags=0)
out={}
for nbytes in (64<<20, 1200<<20, 8<<30):
    for w in (16,32,64,128):
        for bps in (4,8):
            out["%dMB_w%d_bps%d"%(nbytes>>20,w,bps)] = ctx.gather_probe(nbytes,w,bps)
print(json.dumps(out,indent=1))
PY
cat gpurun_out/gather_probe.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 300 python scripts/profile_bounds.py packed 1024 > gpurun_out/plain_profile.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_bounds_multi -s 3 -c 2 -f -o gpurun_out/prof_bounds_packed python scripts/profile_bounds.py packed 1024 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-bnb > gpurun_out/bench_short.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_short.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-bnb > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log; wc -l gpurun_out/launches_bench_short.csv
