#!/bin/bash
# round 2, GPU call 8: WIDE kernel vs two-sub-segment kernel on the bench scene, element by element
set -u
O=gpurun_out/r2h
mkdir -p $O
python benchmarks/experiments/compare_libs.py run $O/shipped.npz > $O/log.txt 2>&1
for v in same nowide w4gb2 w4wpb8; do
  DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_$v.so python benchmarks/experiments/compare_libs.py run $O/$v.npz >> $O/log.txt 2>&1
done
for v in same nowide w4gb2 w4wpb8; do echo "== shipped vs $v" >> $O/log.txt; python benchmarks/experiments/compare_libs.py diff $O/shipped.npz $O/$v.npz >> $O/log.txt 2>&1; done
cat $O/log.txt
