#!/bin/bash
# round 2, GPU call 33: texture axes without explicit clamps, 32-bit ray -> pose division: tests + bench
set -u
O=gpurun_out/r2ag
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rf > $O/pytest.log 2>&1; tail -4 $O/pytest.log
timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline > $O/bench.json 2> $O/bench.err
python -c "import json; d=json.load(open('$O/bench.json')); print(round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['loss'], d['roofline']['frac'])"
timeout 300 python benchmarks/run_configs.py --configs 3f --layout texture 2>&1 | cut -c1-140
