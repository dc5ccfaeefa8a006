#!/bin/bash
# round 2, GPU call 55: a single pose-recovery step with two to four warps per ray (SPLIT): parity, config-2 timings on / off; e2e modes
set -u
O=gpurun_out/r2bc
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "split_over_warps or multi_pass or graphed or config2 or fused_mse or edge_shapes or randomised" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
for c in 1 0; do
  DIFFUS_SPLIT=$c timeout 300 python benchmarks/run_configs.py --configs 2 > $O/config2_split$c.jsonl 2> $O/config2_split$c.err
  echo "split=$c"; python -c "
import json
for l in open('$O/config2_split$c.jsonl'):
    d=json.loads(l); print('  ', round(d['ms']*1000,2), 'us', d['config'][:90])"
done
for m in graph eager; do
  timeout 600 python bench.py --steps 200 --e2e $m --no-extras --no-cpu-baseline > $O/bench_$m.json 2> $O/bench_$m.err
  python -c "
import json; d=json.load(open('$O/bench_$m.json')); e=d['e2e']
print('$m', d['ms_per_step'], e['ms_per_step'], e['value'], e['h2d_bytes_per_step'], e['d2h_bytes_per_step'])"
done
