#!/bin/bash
# round 2, GPU call 9 (2 GPUs): the NCCL test GPUTEST cannot run on one GPU, both bench arms under torchrun at N = 2
set -u
O=gpurun_out/r2i
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -rs > $O/pytest_multi_gpu.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --no-cpu-config1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err
tail -3 $O/pytest_multi_gpu.log; wc -c $O/*.json; tail -5 $O/bench_n2.err
