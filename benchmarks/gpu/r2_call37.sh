#!/bin/bash
# round 2, GPU call 37: prefix-only forward pre-pass without the attenuation table (config 5)
set -u
O=gpurun_out/r2ak
mkdir -p $O
timeout 600 python benchmarks/experiments/config5_step.py --poses 1024 --layout texture | cut -c1-420
timeout 1500 python -m pytest tests -m gpu -q -rf > $O/pytest.log 2>&1; tail -4 $O/pytest.log
