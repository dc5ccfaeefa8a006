#!/bin/bash
# round 2, GPU call 50: vector loads in the loss sum of the merged reduction: full GPU suite, bench, strong shards
set -u
O=gpurun_out/r2ax
mkdir -p $O
DIFFUS_TOL_REPORT=$O/tol.jsonl timeout 1500 python -m pytest tests -m gpu -q -rf > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 900 python bench.py --steps 200 > $O/bench_full.json 2> $O/bench_full.err; tail -c 300 $O/bench_full.err
python -c "
import json; d=json.load(open('$O/bench_full.json'))
print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config5']['ms_per_step'], d['gpu_launches'])"
timeout 600 python - > $O/strong_shards.jsonl 2> $O/strong_shards.err <<'PY'
import json, torch, bench
bench.max_over_ranks = lambda values, dev, world: values        # one process: the per-rank time itself
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
for world in (1, 2, 4, 8, 8, 1):
    rec = bench.strong_scaling_record(dev, 0, world, torch.cuda.synchronize, "texture")
    print(json.dumps({"world": world, "poses_per_gpu": rec["poses_per_gpu"], "ms_op_calls": rec["ms_per_step_op_calls"], "ms_graph": rec["ms_per_step_cuda_graph"]}), flush=True)
PY
cat $O/strong_shards.jsonl
