#!/bin/bash
# round 2, GPU call 2: tests (new training-loop kernels, noise-aware tolerances), quad-slot scatter A/B, full bench with extras
set -u
O=gpurun_out/r2b
mkdir -p $O
DIFFUS_TOL_REPORT=$O/tol.jsonl DIFFUS_TOL_CALIBRATE=1 timeout 1500 python -m pytest tests -m gpu -q -rf --durations=10 > $O/pytest_calibrate.log 2>&1
for s in trilinear nearest; do
  python benchmarks/experiments/scatter_step.py --sampler $s --poses 4096 --check >> $O/scatter.jsonl 2>> $O/scatter.err
  DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_scatter0.so python benchmarks/experiments/scatter_step.py --sampler $s --poses 4096 --check >> $O/scatter.jsonl 2>> $O/scatter.err
done
python bench.py --steps 100 > $O/bench_full.json 2> $O/bench_full.err
python bench.py --steps 100 --layout brick --no-extras --no-cpu-baseline > $O/bench_brick.json 2> $O/bench_brick.err
python bench.py --steps 100 --no-extras --no-cpu-baseline --e2e eager > $O/bench_eager.json 2> $O/bench_eager.err
python benchmarks/experiments/scatter_step.py --sampler trilinear --poses 1024 --iters 1 > $O/plain_scatter.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:render_bwd -s 2 -c 1 -o $O/prof_scatter_quads \
    python benchmarks/experiments/scatter_step.py --sampler trilinear --poses 1024 --iters 1 > $O/ncu_scatter.log 2>&1
ncu -i $O/prof_scatter_quads.ncu-rep --page raw --csv > $O/prof_scatter_quads.raw.csv 2>/dev/null
ncu -i $O/prof_scatter_quads.ncu-rep --page source --csv > $O/prof_scatter_quads.source.csv 2>/dev/null
rm -f $O/prof_scatter_quads.ncu-rep
ls -la $O
