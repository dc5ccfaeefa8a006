#!/bin/bash
# round 2, GPU call 43: full GPU suite + default bench + reference arm + launch list on the build with the one-CTA-per-ray kernel
set -u
O=gpurun_out/r2aq
mkdir -p $O
DIFFUS_TOL_REPORT=$O/tol.jsonl timeout 1500 python -m pytest tests -m gpu -q -rf > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 900 python bench.py --steps 100 > $O/bench_full.json 2> $O/bench_full.err; tail -c 300 $O/bench_full.err
python -c "
import json; d=json.load(open('$O/bench_full.json'))
print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['config5']['ms_per_step'], d['config5']['hbm_frac_at_36B'], {k: v['ms_per_step'] for k, v in d['config4'].items() if isinstance(v, dict)}, d['strong']['ms_per_step'] if 'strong' in d else None)"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
