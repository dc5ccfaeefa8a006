#!/bin/bash
# round 2, GPU call 41: one-CTA-per-ray kernel for four-pass rays (config 5): parity tests, A/B against the multi-pass kernel
set -u
O=gpurun_out/r2ao
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "four_pass or config5 or stress" > $O/pytest_coop.log 2>&1; tail -5 $O/pytest_coop.log
for c in 1 0; do
  DIFFUS_COOP=$c timeout 600 python benchmarks/experiments/config5_step.py --poses 1024 > $O/config5_coop$c.json 2> $O/config5_coop$c.err
  python -c "import json; d=json.load(open('$O/config5_coop$c.json')); print('coop=$c', d['ms_per_step'], d['gsamples_per_s'], d['hbm_frac_at_36B'])"
done
