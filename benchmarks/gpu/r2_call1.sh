#!/bin/bash
# round 2, GPU call 1: full GPU test-suite with tolerance margins recorded, red.add probe, texture-layout A/B, ncu captures
set -u
O=gpurun_out/r2a
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
DIFFUS_TOL_REPORT=$O/tol.jsonl DIFFUS_TOL_CALIBRATE=1 timeout 1500 python -m pytest tests -m gpu -q -rf --durations=15 > $O/pytest_calibrate.log 2>&1
./benchmarks/micro/red_probe > $O/red_probe.md 2>&1
for layout in brick texture; do
  python bench.py --steps 100 --no-cpu-baseline --layout $layout > $O/bench_$layout.json 2> $O/bench_$layout.err
done
for v in texgb1 texgb4; do
  DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_$v.so python bench.py --steps 100 --no-cpu-baseline --layout texture > $O/bench_$v.json 2> $O/bench_$v.err
done
python benchmarks/run_configs.py --configs 3f --layout texture > $O/cfg3f_texture.jsonl 2>&1
python benchmarks/run_configs.py --configs 3f --layout brick > $O/cfg3f_brick.jsonl 2>&1
for s in trilinear nearest; do
  python benchmarks/experiments/scatter_step.py --sampler $s --poses 4096 >> $O/scatter.jsonl 2>> $O/scatter.err
  python benchmarks/experiments/scatter_step.py --sampler $s --poses 4096 --no-grad >> $O/scatter.jsonl 2>> $O/scatter.err
done
# ncu: the texture fused kernel, and the scatter kernel (each after the same command exited 0 without ncu)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --layout texture > $O/plain_tex.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:render_bwd -s 3 -c 1 -o $O/prof_fused_texture \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --layout texture > $O/ncu_tex.log 2>&1
export_rep() {   # the .ncu-rep files are ~20 MB each and gpurun_out/ is capped at 64 MiB: keep the csv pages, drop the report
  ncu -i $1.ncu-rep --page raw --csv > $1.raw.csv 2>/dev/null
  ncu -i $1.ncu-rep --page source --csv > $1.source.csv 2>/dev/null
  rm -f $1.ncu-rep
}
export_rep $O/prof_fused_texture
for s in trilinear nearest; do
  python benchmarks/experiments/scatter_step.py --sampler $s --poses 1024 --iters 1 > $O/plain_scatter_$s.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:render_bwd -s 2 -c 1 -o $O/prof_scatter_$s \
      python benchmarks/experiments/scatter_step.py --sampler $s --poses 1024 --iters 1 > $O/ncu_scatter_$s.log 2>&1
  export_rep $O/prof_scatter_$s
done
ls -la $O
