#!/bin/bash
cp fast_go_icp_b200/libfgoicp_b200.so /tmp/lib_orig.so
for v in mb2 mb3; do
cp lib_$v.so fast_go_icp_b200/libfgoicp_b200.so
echo "== variant $v"; timeout 120 bash scripts/gpu_phased.sh 2>&1 | grep -E "phased|fix_rot False lb"
done
cp /tmp/lib_orig.so fast_go_icp_b200/libfgoicp_b200.so
