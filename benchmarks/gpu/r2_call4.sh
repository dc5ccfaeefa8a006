#!/bin/bash
# round 2, GPU call 4: collapsed-cell / tld4-order gather (A/B against call 3's 0.624 ms), rolled sub-segment loop variant, tests
set -u
O=gpurun_out/r2d
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rf > $O/pytest.log 2>&1
timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline > $O/bench_texture.json 2> $O/bench_texture.err
DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_hrolled.so timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline > $O/bench_hrolled.json 2> $O/bench_hrolled.err
timeout 300 python bench.py --steps 100 --layout brick --no-extras --no-cpu-baseline > $O/bench_brick.json 2> $O/bench_brick.err
timeout 600 python benchmarks/run_configs.py --configs 3f,5 > $O/configs.jsonl 2> $O/configs.err
tail -3 $O/pytest.log
