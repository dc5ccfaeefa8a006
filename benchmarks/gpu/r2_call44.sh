#!/bin/bash
# round 2, GPU call 44: copy-free GraphedFanPoseStep (results written into the static block), unrolled target copies in the COOP kernel
set -u
O=gpurun_out/r2ar
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "graphed or fan or multi_pass" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 600 python bench.py --steps 200 --no-extras --no-cpu-baseline > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.err
python -c "
import json; d=json.load(open('$O/bench.json'))
print('value', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], 'launches', d['gpu_launches'])"
timeout 600 python benchmarks/experiments/config5_step.py --poses 1024 > $O/config5.json 2> $O/config5.err
python -c "import json; d=json.load(open('$O/config5.json')); print('config5', d['ms_per_step'], d['gsamples_per_s'], d['hbm_frac_at_36B'])"
