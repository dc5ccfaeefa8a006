#!/bin/bash
# round 2, GPU call 5: software-pipelined texture gathers A/B (shipped = 2 tiles x 2 stages), parity tests of the new cell code
set -u
O=gpurun_out/r2e
mkdir -p $O
timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline > $O/bench_pipe2.json 2> $O/bench_pipe2.err
for v in nopipe pipe1 gb4; do
  DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_$v.so timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline > $O/bench_$v.json 2> $O/bench_$v.err
done
timeout 1500 python -m pytest tests -m gpu -q -rf > $O/pytest.log 2>&1
tail -8 $O/pytest.log
