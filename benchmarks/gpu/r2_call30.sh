#!/bin/bash
# round 2, GPU call 30: consolidated run of the carveout + 4-tile build: tests, bench, configs, forward batch depth
set -u
O=gpurun_out/r2ad
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rf > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 900 python bench.py --steps 100 > $O/bench_full.json 2> $O/bench_full.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2ad/bench_full.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["loss"])
print({k:round(v["ms_per_step"],3) for k,v in d["config4"].items() if isinstance(v,dict)}, d["nccl_parity"]["ok"], d["config5"]["ms_per_step"], d["strong"]["ms_per_step_cuda_graph"])
PY
timeout 300 python benchmarks/run_configs.py --configs 3f --layout texture 2>&1 | cut -c1-140
DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_fwdgb2.so timeout 300 python benchmarks/run_configs.py --configs 3f --layout texture 2>&1 | cut -c1-140
timeout 300 python benchmarks/run_configs.py --configs 3f --layout brick 2>&1 | cut -c1-140
DIFFUS_B200_LIB=$PWD/diffus_b200/variants/libdiffus_fwdgb2.so timeout 300 python benchmarks/run_configs.py --configs 3f --layout brick 2>&1 | cut -c1-140
