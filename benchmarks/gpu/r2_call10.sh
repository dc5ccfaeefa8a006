#!/bin/bash
# round 2, GPU call 10: the MLP as a piecewise-linear table -- parity tests, forward / backward timings of the three paths, config 4
set -u
O=gpurun_out/r2j
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_training.py -m gpu -q -rf -k "mlp or trainer or training" > $O/pytest_mlp.log 2>&1
tail -15 $O/pytest_mlp.log
timeout 600 python benchmarks/run_configs.py --configs 4 > $O/configs4.jsonl 2> $O/configs4.err
cat $O/configs4.jsonl | cut -c1-400; tail -3 $O/configs4.err
timeout 1500 python -m pytest tests -m gpu -q -rf > $O/pytest.log 2>&1; tail -5 $O/pytest.log
