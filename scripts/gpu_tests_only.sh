#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 300 python scripts/tex_conformance.py gpurun_out/tex_conformance.json > gpurun_out/tex_conformance.log 2>&1; tail -3 gpurun_out/tex_conformance.log
