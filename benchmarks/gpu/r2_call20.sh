#!/bin/bash
# round 2, GPU call 20: FusedTrainer gathering through the texture unit (config 4), PSF kernel test, full suite
set -u
O=gpurun_out/r2t
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -rf > $O/pytest.log 2>&1; tail -6 $O/pytest.log
for g in brick texture; do
  DIFFUS_CONFIG4_GATHER=$g timeout 600 python bench.py --steps 20 --no-cpu-baseline --config5-poses 0 > $O/bench_$g.json 2> $O/bench_$g.err
  python -c "import json; d=json.load(open('$O/bench_$g.json')); print('$g', {k:(round(v['ms_per_step'],3), v['loss_first'], v['loss_last']) for k,v in d['config4'].items() if isinstance(v,dict)})"
done
