#!/bin/bash
# round 2, call N: end-of-round evidence on one GPU -- smoke, the whole GPU suite, both bench arms exactly as the driver runs them
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_n.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_n.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_n.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_n.log
S=$(date +%s); timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_ref_n.json 2> gpurun_out/bench_ref_n.err; echo "ref rc=$? $(( $(date +%s) - S )) s"; cut -c1-300 gpurun_out/bench_ref_n.json
S=$(date +%s); timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_n1_n.json 2> gpurun_out/bench_n1_n.err; echo "bench rc=$? $(( $(date +%s) - S )) s"
python - <<PY
import json
b=json.load(open("gpurun_out/bench_n1_n.json"))
print("value %.4g e2e %.4g frac %.3f in_search %.3f clocks %s" % (b["value"], b["e2e"]["value"], b["roofline"]["frac"], b["roofline"]["in_search"]["frac"], b["clocks"]))
bn=b["bnb"]; print("W5 bnb_ms", bn["bnb_ms"], bn["bnb_ms_all_runs"], "cpp", bn.get("cpp_class",{}).get("bnb_ms"))
for r in b.get("bnb_repo_clouds",[]): print("  ", r.get("case"), "| ms", r.get("bnb_ms"), "ub", r.get("ms_bnb_ub"), "icp", r.get("ms_icp"), "sse", r.get("sse"), "rot", r.get("rot_err_deg"), "t", r.get("t_err_rel"), r.get("error"))
print(b["cpu_baseline"]["value"], b.get("e2e_reference_call_shape",{}).get("value"))
PY
