#!/bin/bash
# round 2, GPU call 3 (re-run after the container was replaced): GPU tests with tolerance margins, full bench line with
# extras, reference arm, launch list, ncu of the fused pose kernel and the volume-gradient scatter kernel
set -u
O=gpurun_out/r2c
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
DIFFUS_TOL_REPORT=$O/tol.jsonl DIFFUS_TOL_CALIBRATE=1 timeout 1500 python -m pytest tests -m gpu -q -rf --durations=10 > $O/pytest_calibrate.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_strict.log 2>&1
timeout 600 python bench.py --steps 100 > $O/bench_full.json 2> $O/bench_full.err
timeout 300 python bench.py --steps 100 --layout brick --no-extras --no-cpu-baseline > $O/bench_brick.json 2> $O/bench_brick.err
timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline --e2e eager > $O/bench_eager.json 2> $O/bench_eager.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_reference.json 2> $O/bench_reference.err
for s in trilinear nearest; do
  timeout 300 python benchmarks/experiments/scatter_step.py --sampler $s --poses 4096 --check >> $O/scatter.jsonl 2>> $O/scatter.err
  timeout 300 python benchmarks/experiments/scatter_step.py --sampler $s --poses 4096 --no-grad >> $O/scatter.jsonl 2>> $O/scatter.err
done
timeout 600 python benchmarks/run_configs.py --configs 1,2,3f,4,5 > $O/configs.jsonl 2> $O/configs.err
# launch list of the bench command (after it exited 0 without ncu above)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_launches.log 2>&1
export_rep() {
  ncu -i $1.ncu-rep --page raw --csv > $1.raw.csv 2>/dev/null
  ncu -i $1.ncu-rep --page source --csv > $1.source.csv 2>/dev/null
  rm -f $1.ncu-rep
}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_bwd -s 3 -c 1 -o $O/prof_fused \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_fused.log 2>&1
export_rep $O/prof_fused
timeout 300 python benchmarks/experiments/scatter_step.py --sampler trilinear --poses 1024 --iters 1 > $O/plain_scatter.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_bwd -s 2 -c 1 -o $O/prof_scatter \
    python benchmarks/experiments/scatter_step.py --sampler trilinear --poses 1024 --iters 1 > $O/ncu_scatter.log 2>&1
export_rep $O/prof_scatter
ls -la $O
