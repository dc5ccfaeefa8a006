#!/bin/bash
# round 2, GPU call 25: shared-memory carveout of the WIDE kernels sized to their resident CTAs (more L1 / texture cache)
set -u
O=gpurun_out/r2y
mkdir -p $O
timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline > $O/bench_carve.json 2> $O/bench_carve.err
DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_nocarve.so timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline > $O/bench_nocarve.json 2> $O/bench_nocarve.err
timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline > $O/bench_carve2.json 2> $O/bench_carve2.err
python -c "
import json
for f in ['carve','nocarve','carve2']:
    d=json.load(open('$O/bench_%s.json'%f)); print(f, d['ms_per_step'], d['e2e']['ms_per_step'], d['loss'])
"
for s in trilinear nearest; do timeout 300 python benchmarks/experiments/config4_step.py --sampler $s; DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_nocarve.so timeout 300 python benchmarks/experiments/config4_step.py --sampler $s; done
