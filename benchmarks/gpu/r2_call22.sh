#!/bin/bash
# round 2, GPU call 22: WIDE sweep in the fused volume-gradient kernels (config 4) vs the two-sub-segment sweep
set -u
O=gpurun_out/r2v
mkdir -p $O
for lib in shipped novgwide; do
  if [ $lib = shipped ]; then unset DIFFUS_B200_LIB; else export DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_$lib.so; fi
  for s in trilinear nearest; do
    timeout 300 python benchmarks/experiments/scatter_step.py --sampler $s --poses 4096 --check >> $O/scatter.jsonl 2>> $O/scatter.err
  done
  timeout 600 python bench.py --steps 20 --no-cpu-baseline --config5-poses 0 > $O/bench_$lib.json 2> $O/bench_$lib.err
  python -c "import json; d=json.load(open('$O/bench_$lib.json')); print('$lib', {k:(round(v['ms_per_step'],3), v['loss_last']) for k,v in d['config4'].items() if isinstance(v,dict)}, d['nccl_parity']['ok'])"
done
cut -c1-260 $O/scatter.jsonl
unset DIFFUS_B200_LIB
timeout 900 python -m pytest tests -m gpu -q -rf > $O/pytest.log 2>&1; tail -3 $O/pytest.log
