#!/bin/bash
# round 2, GPU call 45: trilinear quad-slot scatter with per-axis key components (shipped) vs eight slot keys vs two samples per trip
set -u
O=gpurun_out/r2as
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_training.py -m gpu -q -x -k "volume or fused_mse or randomised or edge_shapes or batched or training or trainer or scatter" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
for lib in shipped slotkeys axis_u2; do
  if [ $lib != shipped ]; then export DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_$lib.so; else unset DIFFUS_B200_LIB; fi
  timeout 300 python benchmarks/experiments/scatter_step.py --sampler trilinear --poses 4096 --iters 5 --check >> $O/scatter.jsonl 2>> $O/scatter.err
  DIFFUS_CONFIG4_GATHER=texture timeout 300 python benchmarks/experiments/config4_step.py >> $O/config4_$lib.jsonl 2>> $O/config4.err
done
unset DIFFUS_B200_LIB
cat $O/scatter.jsonl; tail -n 3 $O/config4_*.jsonl | cut -c1-400
