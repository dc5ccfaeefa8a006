#!/bin/bash
# round 2, GPU call 31: target row read from global memory (no shared-memory staging): 4 CTAs in the 164 KB carveout
set -u
O=gpurun_out/r2ae
mkdir -p $O
python benchmarks/experiments/compare_libs.py run $O/tgt.npz > $O/cmp.txt 2>&1
DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_notgt.so python benchmarks/experiments/compare_libs.py run $O/notgt.npz >> $O/cmp.txt 2>&1
python benchmarks/experiments/compare_libs.py diff $O/tgt.npz $O/notgt.npz >> $O/cmp.txt 2>&1; cat $O/cmp.txt
timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline > $O/bench_tgt.json 2> $O/bench_tgt.err
DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_notgt.so timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline > $O/bench_notgt.json 2> $O/bench_notgt.err
for pct in 71 85; do DIFFUS_CARVEOUT_PCT=$pct timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline > $O/bench_tgt_$pct.json 2> /dev/null; done
python -c "
import json
for f in ['tgt','notgt','tgt_71','tgt_85']:
    d=json.load(open('$O/bench_%s.json'%f)); print(f, round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['loss'])
"
