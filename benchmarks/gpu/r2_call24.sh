#!/bin/bash
# round 2, GPU call 24 (2 GPUs): regression check of the torchrun bench after the FusedTrainer / kernel changes; NCCL test
set -u
O=gpurun_out/r2x
mkdir -p $O
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -rs > $O/pytest_multi_gpu.log 2>&1; tail -2 $O/pytest_multi_gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 100 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2x/bench_n2.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
print(d["strong"]["ms_per_step_cuda_graph"], {k:(round(v["ms_per_step"],3), v["loss_first"], v["loss_last"]) for k,v in d["config4"].items() if isinstance(v,dict)}, d["nccl_parity"])
PY
