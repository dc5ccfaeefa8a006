#!/bin/bash
# round 2, GPU call 51: loads in flight in the piecewise-linear MLP kernels, vector loads in reduce_sum: MLP / training tests, timings
set -u
O=gpurun_out/r2ay
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "mlp or training or trainer or impedance or train_step or slice" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 300 python benchmarks/experiments/mlp_paths.py > $O/mlp_paths.jsonl 2> $O/mlp_paths.err; cut -c1-300 $O/mlp_paths.jsonl | tail -8
for s in trilinear nearest; do
  timeout 300 python benchmarks/experiments/config4_step.py --sampler $s --steps 10 >> $O/config4.jsonl 2>> $O/config4.err
done
cat $O/config4.jsonl
