#!/bin/bash
# round 2, GPU call 27: sweep of the preferred shared-memory carveout of the fused pose kernel
set -u
O=gpurun_out/r2aa
mkdir -p $O
for pct in 50 58 60 65 70 72 75 79 82 86 90 100; do
  DIFFUS_CARVEOUT_PCT=$pct timeout 300 python bench.py --steps 100 --no-extras --no-cpu-baseline > $O/bench_$pct.json 2> $O/bench_$pct.err
  python -c "import json; d=json.load(open('$O/bench_$pct.json')); print($pct, round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))"
done
DIFFUS_CARVEOUT_PCT=79 timeout 300 ncu --metrics launch__shared_mem_config_size,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,gpu__time_duration.sum -k regex:render_bwd -s 3 -c 1 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline 2>&1 | grep -E "shared_mem_config|occupancy_limit|warps_active|hit_rate|duration" 
DIFFUS_CARVEOUT_PCT=86 timeout 300 ncu --metrics launch__shared_mem_config_size,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,gpu__time_duration.sum -k regex:render_bwd -s 3 -c 1 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline 2>&1 | grep -E "shared_mem_config|occupancy_limit|warps_active|hit_rate|duration"
