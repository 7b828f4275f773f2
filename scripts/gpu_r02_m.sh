#!/bin/bash
# round 2, call M: the default bench line at N GPUs exactly as the driver launches it (repo clouds included)
N=${1:-8}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_n1_m.json 2> gpurun_out/bench_n1_m.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n${N}_m.json 2> gpurun_out/bench_n${N}_m.err
fi
echo "bench N=$N rc=$?"; tail -3 gpurun_out/bench_n${N}_m.err
python - <<PY
import json
b=json.load(open("gpurun_out/bench_n${N}_m.json"))
print("value %.4g e2e %.4g frac %.3f in_search %.3f" % (b["value"], b["e2e"]["value"], b["roofline"]["frac"], b["roofline"]["in_search"]["frac"]))
bn=b["bnb"]; print("W5 bnb_ms", bn["bnb_ms"], bn["bnb_ms_all_runs"], "cpp", bn.get("cpp_class",{}).get("bnb_ms"), bn.get("cpp_class",{}).get("bnb_ms_all_runs"))
for r in b.get("bnb_repo_clouds",[]): print("  ", r.get("case"), r.get("bnb_ms"), "ub", r.get("ms_bnb_ub"), "icp", r.get("ms_icp"), "sse", r.get("sse"), r.get("error"))
PY
