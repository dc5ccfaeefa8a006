#!/bin/bash
# round 2, GPU call 35 (8 GPUs): scaling records of the final build at N = 8, 4, 2, 1
set -u
O=gpurun_out/r2ai
mkdir -p $O
port=29540
for n in 8 4 2; do
  port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 100 --warmup 3 > $O/bench_n$n.json 2> $O/bench_n$n.err
  echo "n=$n rc=$?"; wc -c $O/bench_n$n.json
done
timeout 600 python bench.py --gpus 1 --steps 100 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "n=1 rc=$?"
