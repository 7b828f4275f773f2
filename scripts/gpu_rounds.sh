#!/bin/bash
timeout 300 python -m pytest tests -m gpu -x -q -k "bnb or phase" 2>&1 | tail -3
echo "== persistent"; FGOICP_BNBR_MIN_CUBES=1000000 REPS=2 timeout 100 python scripts/run_bench.py
echo "== rounds (auto)"; LEVELS_LOG=1 timeout 100 python scripts/run_bench.py
for mp in 1000 6000 12000; do echo "== rounds min_pairs $mp"; FGOICP_BNBR_MIN_PAIRS=$mp REPS=2 timeout 100 python scripts/run_bench.py; done
