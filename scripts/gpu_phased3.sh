#!/bin/bash
cp fast_go_icp_b200/libfgoicp_b200.so /tmp/lib_orig.so
for iw in 32 64 128; do for lag in 0 2 3; do
cp lib_iw$iw.so fast_go_icp_b200/libfgoicp_b200.so
echo "== inv_w $iw lag $lag"; FGOICP_PHASED_LAG=$lag timeout 120 bash scripts/gpu_phased.sh 2>&1 | grep -E "phased 1 n_rot 4096 fix_rot False|fix_rot False lb"
done; done
cp /tmp/lib_orig.so fast_go_icp_b200/libfgoicp_b200.so
