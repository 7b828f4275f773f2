#!/bin/bash
# round 2, call A: persistent ICP loop sanity + timing, sin constants from the reference build, run() parity vs the reference
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu_a.txt
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_a.log 2>&1; SM=$?; echo "smoke rc=$SM"; tail -2 gpurun_out/smoke_a.log
if [ $SM -ne 0 ]; then export FGOICP_ICP_MODE=1; echo "persistent loop FAILED smoke: continuing with the launch chain"; fi
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "icp or memo or nn or reference" > gpurun_out/pytest_icp_a.log 2>&1; echo "pytest icp rc=$?"; tail -5 gpurun_out/pytest_icp_a.log
# sin constants: reference build vs the library under test
timeout 60 python - > gpurun_out/reference_sin.json 2> gpurun_out/reference_sin.err <<'PY'
import json, numpy as np
from oracle import ref as REF
from fast_go_icp_b200 import capi, workloads
spans = np.array([1.0, 0.5, 0.25, 0.125, 0.0625, 0.03125], np.float32)
r = REF.rot_sin(spans)
w = workloads.synthetic_pair(nt=500, ns=50, seed=1)
from fast_go_icp_b200 import driver
pp = driver.preprocess(w["model"], w["data"])
ctx = capi.Context(pp["model"], pp["data"], pp["bbox_min"], pp["bbox_max"], 0.1, flags=0)
d = ctx.rot_sin(spans)
import math
h = np.array([np.sin(np.float32(np.float32(np.float32(s * np.float32(1.732050807568877)) * np.float32(3.141592653589793)) * np.float32(0.5)), dtype=np.float32) for s in spans], np.float32)
print(json.dumps(dict(spans=spans.tolist(), reference_build_bits=[int(x) for x in r.view(np.uint32)], library_bits=[int(x) for x in d.view(np.uint32)],
                      numpy_sinf_bits=[int(x) for x in h.view(np.uint32)], reference_build=r.tolist())))
PY
echo "sin rc=$?"; cat gpurun_out/reference_sin.json
# timing: loop kernel vs launch chain on W1 / W5 / W3
for MODE in 0 1; do
  if [ $SM -ne 0 ] && [ $MODE -eq 0 ]; then continue; fi
  FGOICP_ICP_MODE=$MODE timeout 240 python scripts/bench_repo_clouds.py --no-baselines --reps 2 --only "W1 bunny res 0.005,W3 dragon mse,W4,W5" --skip "mse 1e-5" --out repo_clouds_icpmode$MODE.json 2>&1 | tail -6
done
timeout 900 python scripts/run_parity_r02.py --cap 150 > gpurun_out/run_parity_a.log 2>&1; echo "parity rc=$?"; tail -12 gpurun_out/run_parity_a.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_a.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_a.log
