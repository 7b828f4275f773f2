#!/bin/bash
# round 2, call B: ICP loop kernel shapes + stage timers, sorted-run inner search vs bitonic, new bench.py, full GPU suite
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_b.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_b.log
for SHAPE in 512x2 512x2o 256x5o 256x4 512x1; do
  echo "== ICP shape $SHAPE"
  FGOICP_ICP_SHAPE=$SHAPE FGOICP_ICP_LOG=1 timeout 200 python scripts/bench_repo_clouds.py --no-baselines --reps 2 --only "W3 dragon mse,W5" --out icp_shape_$SHAPE.json 2> gpurun_out/icp_shape_$SHAPE.err | tail -2
  grep "icp loop" gpurun_out/icp_shape_$SHAPE.err | awk '{n++; if (n<=3 || n%40==0) print}' | head -12
done
echo "== inner search kernels on W5"
for K in bitonic merge; do
  FGOICP_BNB_KERNEL=$K timeout 120 python - <<PY
import json, numpy as np
from fast_go_icp_b200 import capi, driver, workloads
w = workloads.synthetic_pair()
out = []
for rep in range(3):
    g = driver.FastGoICP(w["model"], w["data"], 0.005, 1e-4, flags=capi.BUILD_PACKED)
    g.run(); st = g.stats
    out.append((st["run_ms"], st["ms_bnb_ub"], st["ms_bnb_lb"], st["ms_icp"], float(g.best_sse), st["bound_evals"]))
    lv = [(l["cubes"], round(l["ms_ub"], 2), round(l["ms_icp"], 2), round(l["ms_lb"], 2)) for l in st["level_log"]]
    g.close()
print("$K", out, lv)
PY
done
timeout 600 python bench.py > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_b.err
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref_b.json 2> gpurun_out/bench_ref_b.err; echo "bench ref rc=$?"; tail -3 gpurun_out/bench_ref_b.err
python - <<'PY'
import json
try:
    b = json.load(open('gpurun_out/bench_b.json'))
    print({k: b[k] for k in ('value', 'ms_per_step')}, 'e2e', b['e2e']['value'], 'frac', b['roofline']['frac'], 'in_search', b['roofline'].get('in_search'))
    print('ref-shape', b.get('e2e_reference_call_shape'))
    bn = b['bnb']; print({k: bn[k] for k in bn if k != 'levels'})
    for l in bn['levels']: print({k: l[k] for k in ('span', 'cubes', 'icps', 'evals', 'ms_ub', 'ms_icp', 'ms_lb')})
    for r in b.get('bnb_repo_clouds', []): print({k: r.get(k) for k in ('case', 'bnb_ms', 'ms_icp', 'ms_bnb_ub', 'sse', 'bound_evals_local', 'in_search_evals_per_s_local', 'error')})
    print('cpu', b['cpu_baseline'])
except Exception as e:
    print('bench parse failed', e)
try:
    r = json.load(open('gpurun_out/bench_ref_b.json')); print('REF', r['value'], r.get('bnb'), r['cpu_baseline']['sample'][:80])
except Exception as e:
    print('ref parse failed', e)
PY
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_b.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_b.log
