#!/bin/bash
# round 2, GPU call 18: host-time profile of the reference-signature calls
set -u
O=gpurun_out/r2r
mkdir -p $O
timeout 600 python benchmarks/experiments/host_overhead.py > $O/host.txt 2> $O/host.err; tail -3 $O/host.err
