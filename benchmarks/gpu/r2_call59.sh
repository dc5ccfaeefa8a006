#!/bin/bash
# round 2, GPU call 59: pose loads before the attenuation fill (shipped) vs after (late_pose) vs the previous commit (prev: seven butterfly sums too)
set -u
O=gpurun_out/r2bg
mkdir -p $O
for lib in shipped late_pose prev shipped late_pose prev; do
  if [ $lib != shipped ]; then export DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_$lib.so; else unset DIFFUS_B200_LIB; fi
  timeout 600 python bench.py --steps 300 --no-extras --no-cpu-baseline > $O/tmp.json 2>> $O/bench.err
  echo "$lib $(python -c "import json; d=json.load(open('$O/tmp.json')); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])")" | tee -a $O/ab.txt
done
unset DIFFUS_B200_LIB
timeout 600 python benchmarks/experiments/config5_step.py --poses 1024 > $O/config5.json 2> $O/config5.err
python -c "import json; d=json.load(open('$O/config5.json')); print('config5', d['ms_per_step'], d['gsamples_per_s'])"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x > $O/pytest.log 2>&1; tail -2 $O/pytest.log
