#!/bin/bash
# round 2, GPU call 47: both reductions of the fused pose step in one launch -- parity, headline, and the 128-pose shard of a strong-scaled sweep
set -u
O=gpurun_out/r2au
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x > $O/pytest.log 2>&1; tail -3 $O/pytest.log
for P in 1024 128; do
  timeout 600 python bench.py --steps 300 --poses $P --no-extras --no-cpu-baseline > $O/bench_p$P.json 2> $O/bench_p$P.err; tail -c 300 $O/bench_p$P.err
  python -c "
import json; d=json.load(open('$O/bench_p$P.json'))
print('poses $P: value', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], 'launches', d['gpu_launches'])"
done
timeout 300 python benchmarks/run_configs.py --configs 2 > $O/config2.jsonl 2> $O/config2.err; cut -c1-200 $O/config2.jsonl
