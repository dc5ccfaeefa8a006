#!/bin/bash
# round 2, GPU call 38: immediate-offset addressing of the target copies and the gather stores
set -u
O=gpurun_out/r2al
mkdir -p $O
python benchmarks/experiments/compare_libs.py run $O/new.npz 2>&1 | tail -1
timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['loss'], d['roofline']['frac'])"
timeout 1500 python -m pytest tests -m gpu -q -rf > $O/pytest.log 2>&1; tail -4 $O/pytest.log
timeout 300 python benchmarks/experiments/config4_step.py --sampler trilinear; timeout 300 python benchmarks/experiments/config4_step.py --sampler nearest
