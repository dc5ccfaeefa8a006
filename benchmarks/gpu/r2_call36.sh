#!/bin/bash
# round 2, GPU call 36: one-pass attenuation table in the multi-pass backward kernel (config 5 inside the 196 KB carveout)
set -u
O=gpurun_out/r2aj
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rf > $O/pytest.log 2>&1; tail -4 $O/pytest.log
timeout 600 python benchmarks/experiments/config5_step.py --poses 1024 --layout texture | cut -c1-420
timeout 600 python benchmarks/experiments/config5_step.py --poses 1024 --layout brick | cut -c1-420
timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['loss'])"
