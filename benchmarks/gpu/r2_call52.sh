#!/bin/bash
# round 2, GPU call 52 (8 GPUs): scaling records of the final build at N = 8, 4, 1 (the driver's launch line)
set -u
O=gpurun_out/r2az
mkdir -p $O
port=29560
for n in 8 4; do
  port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 100 --warmup 3 > $O/bench_n$n.json 2> $O/bench_n$n.err
  echo "n=$n rc=$?"; wc -c $O/bench_n$n.json
done
timeout 600 python bench.py --gpus 1 --steps 100 --warmup 3 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err; echo "n=1 rc=$?"
python - <<'PY'
import json
for n in (1, 4, 8):
    d = json.load(open(f"gpurun_out/r2az/bench_n{n}.json"))
    c4 = d["config4"]
    print(n, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["strong"]["ms_per_step_cuda_graph"], c4["trilinear"]["ms_per_step"], c4["nearest"]["ms_per_step"], d["nccl_parity"]["ok"])
PY
