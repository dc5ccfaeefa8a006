#!/bin/bash
# round 2, GPU call 53: the driver's round-end sequence on HEAD: GPU suite, smoke, reference arm, default bench
set -u
O=gpurun_out/r2ba
mkdir -p $O
DIFFUS_TOL_REPORT=$O/tol.jsonl timeout 1500 python -m pytest tests -m gpu -q -rf > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"; cut -c1-200 $O/bench_reference.json
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$O/bench_default.json'))
print(d['steps'], d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['config5']['ms_per_step'], {k:v['ms_per_step'] for k,v in d['config4'].items() if isinstance(v,dict)}, d['gpu_launches'], d['clocks'])"
