#!/bin/bash
# round 2, GPU call 21: full bench line of the current build (texture gathers in config 4), reference arm
set -u
O=gpurun_out/r2u
mkdir -p $O
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -2 $O/bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2u/bench_default.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"], d["clocks"])
print({k:(round(v["ms_per_step"],3)) for k,v in d["config4"].items() if isinstance(v,dict)}, d["nccl_parity"]["ok"], d["config5"]["ms_per_step"], d["strong"]["ms_per_step_cuda_graph"])
r=json.load(open("gpurun_out/r2u/bench_reference.json")); print(r["value"], r["steps"], r.get("config1_forward"))
PY
