#!/bin/bash
# round 2, GPU call 14: new tests (MLP input gradient, singular-interface ray), red.add probe, forward sweep on TEXTURE,
# config 5 under ncu (DRAM bytes per launch of both kernels)
set -u
O=gpurun_out/r2n
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -rf -k "mlp or singular or zero_over_zero" > $O/pytest_new.log 2>&1; tail -5 $O/pytest_new.log
./benchmarks/micro/red_probe > $O/red_probe.txt 2>&1; cat $O/red_probe.txt
timeout 300 python benchmarks/run_configs.py --configs 3f --layout texture > $O/cfg3f_texture.jsonl 2>&1; cat $O/cfg3f_texture.jsonl | cut -c1-200
timeout 600 python benchmarks/experiments/config5_step.py --poses 1024 --layout texture > $O/config5_texture.json 2> $O/config5.err; cat $O/config5_texture.json | cut -c1-600
timeout 600 python benchmarks/experiments/config5_step.py --poses 1024 --layout brick > $O/config5_brick.json 2>> $O/config5.err; cat $O/config5_brick.json | cut -c1-300
timeout 300 python benchmarks/experiments/config5_step.py --poses 256 --iters 1 > /dev/null 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_ -s 4 -c 2 -o $O/prof_config5 \
    python benchmarks/experiments/config5_step.py --poses 256 --iters 1 > $O/ncu_config5.log 2>&1
ncu -i $O/prof_config5.ncu-rep --page raw --csv > $O/prof_config5.raw.csv 2>/dev/null
ncu -i $O/prof_config5.ncu-rep --page source --csv > $O/prof_config5.source.csv 2>/dev/null
rm -f $O/prof_config5.ncu-rep
