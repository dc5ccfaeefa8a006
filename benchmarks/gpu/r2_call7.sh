#!/bin/bash
# round 2, GPU call 7: variants of the WIDE kernel (gather batch depth, rays per CTA)
set -u
O=gpurun_out/r2g
mkdir -p $O
timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline > $O/bench_shipped.json 2> $O/bench_shipped.err
for v in w4gb2pipe w4gb2 w4wpb8 w4wpb2 w3; do
  DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_$v.so timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline > $O/bench_$v.json 2> $O/bench_$v.err
done
