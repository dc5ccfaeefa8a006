#!/bin/bash
# round 2, GPU call 23: launch list of the config-4 training step (both samplers)
set -u
O=gpurun_out/r2w
mkdir -p $O
for s in trilinear nearest; do
  timeout 300 python benchmarks/experiments/config4_step.py --sampler $s > $O/plain_$s.json 2>&1 && cat $O/plain_$s.json &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$s.csv \
      python benchmarks/experiments/config4_step.py --sampler $s --steps 2 > $O/ncu_$s.log 2>&1
done
ls $O
