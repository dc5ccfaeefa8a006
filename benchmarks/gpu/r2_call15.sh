#!/bin/bash
# round 2, GPU call 15 (8 GPUs): both bench arms under torchrun exactly as the driver launches them
set -u
O=gpurun_out/r2o
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 100 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err
echo "rc=$?"; wc -c $O/bench_n8.json; tail -3 $O/bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 8 --steps 1 --warmup 0 --no-cpu-config1 > $O/bench_ref_n8.json 2> $O/bench_ref_n8.err
echo "rc=$?"; wc -c $O/bench_ref_n8.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 100 --warmup 3 > $O/bench_n4.json 2> $O/bench_n4.err
echo "rc=$?"; wc -c $O/bench_n4.json
