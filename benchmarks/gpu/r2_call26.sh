#!/bin/bash
# round 2, GPU call 26: right-sized shared-memory carveout for every render kernel; WIDE at 3 CTAs with the larger L1; tests
set -u
O=gpurun_out/r2z
mkdir -p $O
timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline > $O/bench_carve.json 2> $O/bench_carve.err
DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_w3carve.so timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline > $O/bench_w3carve.json 2> $O/bench_w3carve.err
timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline --layout brick > $O/bench_brick.json 2> $O/bench_brick.err
python -c "
import json
for f in ['carve','w3carve','brick']:
    d=json.load(open('$O/bench_%s.json'%f)); print(f, d['ms_per_step'], d['e2e']['ms_per_step'], d['loss'])
"
timeout 600 python benchmarks/run_configs.py --configs 1,2,3f > $O/configs_brick.jsonl 2>&1; cut -c1-150 $O/configs_brick.jsonl
timeout 300 python benchmarks/run_configs.py --configs 3f --layout texture > $O/configs_tex.jsonl 2>&1; cut -c1-150 $O/configs_tex.jsonl
timeout 600 python benchmarks/experiments/config5_step.py --poses 1024 --layout texture | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -q -rf > $O/pytest.log 2>&1; tail -3 $O/pytest.log
