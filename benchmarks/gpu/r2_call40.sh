#!/bin/bash
# round 2, GPU call 40: ncu of the final fused pose kernel (final build of the round), launch list, full bench, tests
set -u
O=gpurun_out/r2an
mkdir -p $O
timeout 900 python bench.py --steps 100 > $O/bench_full.json 2> $O/bench_full.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_bwd -s 3 -c 1 -o $O/prof_fused \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_fused.log 2>&1
ncu -i $O/prof_fused.ncu-rep --page raw --csv > $O/prof_fused.raw.csv 2>/dev/null
ncu -i $O/prof_fused.ncu-rep --page source --csv > $O/prof_fused.source.csv 2>/dev/null
rm -f $O/prof_fused.ncu-rep
timeout 600 python benchmarks/run_configs.py --configs 1,2,3f,4,5 > $O/configs.jsonl 2> $O/configs.err
timeout 300 python benchmarks/run_configs.py --configs 3f --layout texture >> $O/configs.jsonl 2>> $O/configs.err
DIFFUS_TOL_REPORT=$O/tol.jsonl timeout 1500 python -m pytest tests -m gpu -q -rf > $O/pytest.log 2>&1; tail -3 $O/pytest.log
