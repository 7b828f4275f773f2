#!/bin/bash
cp fast_go_icp_b200/libfgoicp_b200.so /tmp/lib_orig.so
for v in g4w16 g2w32 g2w16; do for lag in 1 2; do
cp lib_$v.so fast_go_icp_b200/libfgoicp_b200.so
echo "== variant $v lag $lag"; FGOICP_PHASED_LAG=$lag timeout 120 bash scripts/gpu_phased.sh 2>&1 | grep -E "phased 1 n_rot 4096 fix_rot False|fix_rot False lb"
done; done
cp /tmp/lib_orig.so fast_go_icp_b200/libfgoicp_b200.so
