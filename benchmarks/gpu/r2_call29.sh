#!/bin/bash
# round 2, GPU call 29: deeper gather batches with the 196 KB carveout
set -u
O=gpurun_out/r2ac
mkdir -p $O
for v in gb4 gb4pipe gb8 gb16; do
  DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_$v.so timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline > $O/bench_$v.json 2> $O/bench_$v.err
done
python -c "
import json
for f in ['gb4','gb4pipe','gb8','gb16']:
    d=json.load(open('$O/bench_%s.json'%f)); print(f, round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['roofline']['frac'], d['loss'])
"
