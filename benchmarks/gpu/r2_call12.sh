#!/bin/bash
# round 2, GPU call 12: piecewise-linear MLP after the incremental table build + joint accumulation
set -u
O=gpurun_out/r2l
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_training.py -m gpu -q -rf -k "mlp or trainer or training" > $O/pytest_mlp.log 2>&1
tail -6 $O/pytest_mlp.log
for n in 65536 1048576 16777216 134217728; do python benchmarks/experiments/mlp_paths.py --n $n >> $O/sweep.jsonl 2>> $O/sweep.err; done
cat $O/sweep.jsonl; tail -3 $O/sweep.err
python benchmarks/experiments/mlp_paths.py --iters 1 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mlp_pwl -c 4 -s 4 -o $O/prof_pwl python benchmarks/experiments/mlp_paths.py --iters 1 > $O/ncu.log 2>&1
ncu -i $O/prof_pwl.ncu-rep --page raw --csv > $O/prof_pwl.raw.csv 2>/dev/null
ncu -i $O/prof_pwl.ncu-rep --page source --csv > $O/prof_pwl.source.csv 2>/dev/null
rm -f $O/prof_pwl.ncu-rep
