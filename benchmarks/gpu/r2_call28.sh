#!/bin/bash
# round 2, GPU call 28: gather batch depth again, now with the 196 KB carveout (60 KB of L1 / texture cache)
set -u
O=gpurun_out/r2ab
mkdir -p $O
timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline > $O/bench_shipped.json 2> $O/bench_shipped.err
for v in gb2pipe gb2 gb1 gb4; do
  DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_$v.so timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline > $O/bench_$v.json 2> $O/bench_$v.err
done
python -c "
import json
for f in ['shipped','gb2pipe','gb2','gb1','gb4']:
    d=json.load(open('$O/bench_%s.json'%f)); print(f, round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['roofline']['frac'])
"
timeout 300 python benchmarks/run_configs.py --configs 3f --layout texture 2>&1 | cut -c1-140
