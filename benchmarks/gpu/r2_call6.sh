#!/bin/bash
# round 2, GPU call 6: single 512-column reverse sweep (WideGeo) at 4 and 5 CTAs per SM vs the shipped two-sub-segment sweep; tests
set -u
O=gpurun_out/r2f
mkdir -p $O
timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline > $O/bench_shipped.json 2> $O/bench_shipped.err
for v in wide4 wide5; do
  DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_$v.so timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline > $O/bench_$v.json 2> $O/bench_$v.err
done
timeout 1500 python -m pytest tests -m gpu -q -rf > $O/pytest.log 2>&1
DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_wide4.so timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q -rf -k "config2 or config3 or fused" > $O/pytest_wide4.log 2>&1
tail -8 $O/pytest.log; tail -4 $O/pytest_wide4.log
