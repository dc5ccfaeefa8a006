#!/bin/bash
# round 2, GPU call 39: config 4 gather layout per sampler again, now with the right-sized carveout
set -u
O=gpurun_out/r2am
mkdir -p $O
for g in brick texture; do
  DIFFUS_CONFIG4_GATHER=$g timeout 600 python bench.py --steps 20 --no-cpu-baseline --config5-poses 0 > $O/bench_$g.json 2> $O/bench_$g.err
  python -c "import json; d=json.load(open('$O/bench_$g.json')); print('$g', {k:(round(v['ms_per_step'],3)) for k,v in d['config4'].items() if isinstance(v,dict)})"
done
