#!/bin/bash
# round 2, call H (re-entry): state of the tree -- smoke, the whole GPU suite, both bench arms, round sizes of the inner searches on W5
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_h.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_h.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_h.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_h.log
timeout 600 python bench.py > gpurun_out/bench_n1_h.json 2> gpurun_out/bench_n1_h.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_n1_h.json
timeout 400 python bench.py --impl reference --steps 5 > gpurun_out/bench_ref_h.json 2> gpurun_out/bench_ref_h.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref_h.json
FGOICP_BNBR_MIN_CUBES=1 FGOICP_BNBR_LOG=1 timeout 200 python scripts/profile_run.py > gpurun_out/rounds_h.log 2> gpurun_out/rounds_h.err; echo "rounds rc=$?"; cat gpurun_out/rounds_h.log; grep -c bnbr gpurun_out/rounds_h.err
