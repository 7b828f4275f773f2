#!/bin/bash
# first GPU visit: parity tests, smoke, texture conformance, sampler sweep, first bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -5 gpurun_out/smoke.log
timeout 300 python scripts/tex_conformance.py gpurun_out/tex_conformance.json > gpurun_out/tex_conformance.log 2>&1; tail -40 gpurun_out/tex_conformance.log
timeout 600 python scripts/sampler_sweep.py gpurun_out/sampler_sweep.json > gpurun_out/sampler_sweep.log 2>&1; tail -12 gpurun_out/sampler_sweep.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
