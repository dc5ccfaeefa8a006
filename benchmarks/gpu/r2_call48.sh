#!/bin/bash
# round 2, GPU call 48: strong-scaling shard sizes on one GPU (what each rank of an N-GPU sweep of 1024 poses runs), launch list, full bench
set -u
O=gpurun_out/r2av
mkdir -p $O
timeout 600 python - > $O/strong_shards.jsonl 2> $O/strong_shards.err <<'PY'
import json, torch, bench
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
for world in (1, 2, 4, 8):
    rec = bench.strong_scaling_record(dev, 0, world, torch.cuda.synchronize, "texture")
    print(json.dumps({"world": world, "poses_per_gpu": rec["poses_per_gpu"], "ms_op_calls": rec["ms_per_step_op_calls"], "ms_graph": rec["ms_per_step_cuda_graph"]}), flush=True)
PY
cat $O/strong_shards.jsonl; tail -3 $O/strong_shards.err
timeout 900 python bench.py --steps 100 > $O/bench_full.json 2> $O/bench_full.err; tail -c 300 $O/bench_full.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_launches.log 2>&1
python -c "
import json; d=json.load(open('$O/bench_full.json'))
print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['config5']['ms_per_step'], d['gpu_launches'])"
