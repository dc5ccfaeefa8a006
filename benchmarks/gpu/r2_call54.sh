#!/bin/bash
# round 2, GPU call 54: the end-to-end loop through the three public APIs (fan parameters / explicit directions graphed / eager autograd)
set -u
O=gpurun_out/r2bb
mkdir -p $O
for m in fan graph eager; do
  timeout 600 python bench.py --steps 200 --e2e $m --no-extras --no-cpu-baseline > $O/bench_$m.json 2> $O/bench_$m.err
  python -c "
import json; d=json.load(open('$O/bench_$m.json')); e=d['e2e']
print('$m', d['ms_per_step'], e['ms_per_step'], e['value'], e['h2d_bytes_per_step'], e['d2h_bytes_per_step'])"
done
