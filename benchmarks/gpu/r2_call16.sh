#!/bin/bash
# round 2, GPU call 16: volume-gradient scatter variants (branch-region vs predicated slot miss), parity of both
set -u
O=gpurun_out/r2p
mkdir -p $O
for s in trilinear nearest; do
  timeout 300 python benchmarks/experiments/scatter_step.py --sampler $s --poses 4096 --check >> $O/scatter.jsonl 2>> $O/scatter.err
done
DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_scpred.so timeout 300 python benchmarks/experiments/scatter_step.py --sampler trilinear --poses 4096 --check >> $O/scatter.jsonl 2>> $O/scatter.err
cat $O/scatter.jsonl | cut -c1-330
timeout 1500 python -m pytest tests -m gpu -q -rf > $O/pytest.log 2>&1; tail -4 $O/pytest.log
