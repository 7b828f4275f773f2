#!/bin/bash
for lag in 0 1 2 4 8; do echo "== lag $lag"; FGOICP_PHASED_LAG=$lag timeout 120 bash scripts/gpu_phased.sh 2>&1 | grep -E "phased 1|equal"; done
