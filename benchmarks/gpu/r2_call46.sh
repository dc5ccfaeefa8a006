#!/bin/bash
# round 2, GPU call 46 (2 GPUs): the driver's launch line at N = 2 on the final build, plus the 2-rank NCCL test
set -u
O=gpurun_out/r2at
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
echo "rc=$?"; wc -l $O/bench_n2.json
python -c "
import json; d=json.load(open('$O/bench_n2.json'))
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['strong'].get('ms_per_step_cuda_graph'), {k:(v.get('ms_per_step') if isinstance(v,dict) else None) for k,v in d['config4'].items()}, d['nccl_parity'])"
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q > $O/pytest_multi.log 2>&1; tail -2 $O/pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; echo "ref rc=$?"; cut -c1-300 $O/bench_ref_n2.json
