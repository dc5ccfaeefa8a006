#!/bin/bash
# round 2, GPU call 60: ncu of the final fused pose kernel (transposing final reduction, early pose load) + launch list
set -u
O=gpurun_out/r2bh
mkdir -p $O
timeout 600 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/bench_short.json 2> $O/bench_short.err; echo "plain rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_bwd -s 3 -c 1 -o $O/prof_fused \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_fused.log 2>&1
ncu -i $O/prof_fused.ncu-rep --page raw --csv > $O/prof_fused.raw.csv 2>/dev/null
ncu -i $O/prof_fused.ncu-rep --page source --csv > $O/prof_fused.source.csv 2>/dev/null
rm -f $O/prof_fused.ncu-rep
ls -la $O
