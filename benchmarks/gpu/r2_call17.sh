#!/bin/bash
# round 2, GPU call 17: gather-locality probe of the fused pose step
set -u
O=gpurun_out/r2q
mkdir -p $O
timeout 600 python benchmarks/experiments/locality_probe.py > $O/locality.jsonl 2> $O/locality.err; cat $O/locality.jsonl; tail -2 $O/locality.err
