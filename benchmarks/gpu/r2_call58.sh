#!/bin/bash
# round 2, GPU call 58: the ray's seven final sums in one transposing warp reduction (shipped) vs seven butterfly reductions (variant)
set -u
O=gpurun_out/r2bf
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x > $O/pytest.log 2>&1; tail -2 $O/pytest.log
for lib in shipped sum_old shipped sum_old; do
  if [ $lib != shipped ]; then export DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_$lib.so; else unset DIFFUS_B200_LIB; fi
  timeout 600 python bench.py --steps 300 --no-extras --no-cpu-baseline > $O/tmp.json 2>> $O/bench.err
  echo "$lib $(python -c "import json; d=json.load(open('$O/tmp.json')); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])")" | tee -a $O/ab.txt
done
