#!/bin/bash
# round 2, call I: strong scaling of run() on W5 over N GPUs (Python driver over NCCL + the C++ class in one process)
N=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 5 --warmup 3 --no-repo-clouds --no-cpu > gpurun_out/bench_n${N}_i.json 2> gpurun_out/bench_n${N}_i.err
echo "bench N=$N rc=$?"
python - <<PY
import json
b=json.load(open("gpurun_out/bench_n${N}_i.json"))
bn=b["bnb"]
print({k:bn[k] for k in ("bnb_ms","bnb_ms_all_runs","ms_bnb_ub","ms_icp","ms_bnb_lb","ms_first_icp","ms_final_icp","ms_search_wall","ms_in_abi_calls","ms_in_exchange","exchanges","sse","in_search_evals_per_s")})
for l in bn["levels"]: print({k:(round(v,2) if isinstance(v,float) else v) for k,v in l.items()})
print(bn.get("cpp_class"))
print(bn.get("per_rank_ms"))
print("value", b["value"], "e2e", b["e2e"]["value"])
PY
if [ "$N" = "2" ]; then timeout 300 python -m pytest tests/test_cpp_api_gpu.py -x -q -m gpu -k "two_gpus or sharded" 2>&1 | tail -3; fi
# C++ class alone on 1 device (fresh process, no warm-up run: kernels are preloaded at context creation)
python - <<PY
import bench
w, _, _ = bench.build_workload(0)
print("cpp 1 device:", bench.cpp_class_run(w, 1))
PY
