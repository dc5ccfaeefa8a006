#!/bin/bash
# round 2, GPU call 49: what each rank of an N-GPU strong-scaled sweep of 1024 poses runs, timed on one GPU (no collective is on that path)
set -u
O=gpurun_out/r2aw
mkdir -p $O
timeout 600 python - > $O/strong_shards.jsonl 2> $O/strong_shards.err <<'PY'
import json, torch, bench
bench.max_over_ranks = lambda values, dev, world: values        # one process: the per-rank time itself
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
for world in (1, 2, 4, 8, 8, 1):
    rec = bench.strong_scaling_record(dev, 0, world, torch.cuda.synchronize, "texture")
    print(json.dumps({"world": world, "poses_per_gpu": rec["poses_per_gpu"], "ms_op_calls": rec["ms_per_step_op_calls"], "ms_graph": rec["ms_per_step_cuda_graph"]}), flush=True)
PY
cat $O/strong_shards.jsonl; tail -3 $O/strong_shards.err
