#!/bin/bash
# round 2, GPU call 56: new zero-impedance-block test; trilinear scatter loop with two samples per trip (variant) vs one
set -u
O=gpurun_out/r2bd
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "zero_impedance_block or one_pass_rays_pose or zero_over_zero" > $O/pytest.log 2>&1; tail -6 $O/pytest.log
for lib in shipped scatter_u2 shipped scatter_u2; do
  if [ $lib != shipped ]; then export DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_$lib.so; else unset DIFFUS_B200_LIB; fi
  DIFFUS_CONFIG4_GATHER=texture timeout 300 python benchmarks/experiments/config4_step.py --steps 10 > $O/tmp.json 2>> $O/config4.err
  echo "$lib $(cat $O/tmp.json)" | tee -a $O/config4_ab.txt
done
unset DIFFUS_B200_LIB
