#!/bin/bash
# round 2, GPU call 13: consolidated numbers of the current build (WIDE pose kernel, piecewise-linear MLP): tests with margins,
# full bench line, reference arm, all configs, launch list + ncu of the fused pose kernel
set -u
O=gpurun_out/r2m
mkdir -p $O
DIFFUS_TOL_REPORT=$O/tol.jsonl timeout 1500 python -m pytest tests -m gpu -q -rf --durations=8 > $O/pytest.log 2>&1
tail -4 $O/pytest.log
timeout 900 python bench.py --steps 100 > $O/bench_full.json 2> $O/bench_full.err
timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline --e2e eager > $O/bench_eager.json 2> $O/bench_eager.err
timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline --e2e graph > $O/bench_graph.json 2> $O/bench_graph.err
timeout 600 python benchmarks/run_configs.py --configs 1,2,3f,4,5 > $O/configs.jsonl 2> $O/configs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_bwd -s 3 -c 1 -o $O/prof_fused \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_fused.log 2>&1
ncu -i $O/prof_fused.ncu-rep --page raw --csv > $O/prof_fused.raw.csv 2>/dev/null
ncu -i $O/prof_fused.ncu-rep --page source --csv > $O/prof_fused.source.csv 2>/dev/null
rm -f $O/prof_fused.ncu-rep
ls $O
