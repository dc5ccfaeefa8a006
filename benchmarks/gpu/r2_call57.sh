#!/bin/bash
# round 2, GPU call 57: config 5, the sample before a pass taken from the previous warp's buffer (one more barrier) vs the extra single-lane gather
set -u
O=gpurun_out/r2be
mkdir -p $O
for lib in shipped coop_nb shipped coop_nb; do
  if [ $lib != shipped ]; then export DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_$lib.so; else unset DIFFUS_B200_LIB; fi
  timeout 600 python benchmarks/experiments/config5_step.py --poses 1024 > $O/tmp.json 2>> $O/config5.err
  echo "$lib $(python -c "import json; d=json.load(open('$O/tmp.json')); print(d['ms_per_step'], d['gsamples_per_s'])")" | tee -a $O/config5_ab.txt
done
export DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_coop_nb.so
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "multi_pass or config5 or stress or zero_impedance" > $O/pytest_nb.log 2>&1; tail -2 $O/pytest_nb.log
