#!/bin/bash
# round 2, GPU call 19: smoke() with the fused step and the MLP, host-overhead numbers after the host-path tweaks, GPU suite
set -u
O=gpurun_out/r2s
mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -5 $O/smoke.log
timeout 600 python benchmarks/experiments/host_overhead.py > $O/host.txt 2> $O/host.err; grep host_us $O/host.txt
timeout 600 python benchmarks/run_configs.py --configs 1,2 > $O/configs12.jsonl 2>&1; cut -c1-160 $O/configs12.jsonl
timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline --e2e eager > $O/bench_eager.json 2> $O/bench_eager.err
python -c "import json; d=json.load(open('$O/bench_eager.json')); print('eager', d['ms_per_step'], d['e2e']['ms_per_step'])"
timeout 1500 python -m pytest tests -m gpu -q -rf > $O/pytest.log 2>&1; tail -3 $O/pytest.log
