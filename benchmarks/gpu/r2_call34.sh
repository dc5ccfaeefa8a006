#!/bin/bash
# round 2, GPU call 34: lane-major staging of the target row (float4 accesses) in the WIDE fused pose kernel
set -u
O=gpurun_out/r2ah
mkdir -p $O
python benchmarks/experiments/compare_libs.py run $O/lm.npz > $O/cmp.txt 2>&1
DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_nolm.so python benchmarks/experiments/compare_libs.py run $O/nolm.npz >> $O/cmp.txt 2>&1
python benchmarks/experiments/compare_libs.py diff $O/lm.npz $O/nolm.npz >> $O/cmp.txt 2>&1; cat $O/cmp.txt
timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline > $O/bench_lm.json 2> $O/bench_lm.err
DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_nolm.so timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline > $O/bench_nolm.json 2> $O/bench_nolm.err
timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline > $O/bench_lm2.json 2> $O/bench_lm2.err
python -c "
import json
for f in ['lm','nolm','lm2']:
    d=json.load(open('$O/bench_%s.json'%f)); print(f, round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['loss'])
"
timeout 1500 python -m pytest tests -m gpu -q -rf > $O/pytest.log 2>&1; tail -4 $O/pytest.log
