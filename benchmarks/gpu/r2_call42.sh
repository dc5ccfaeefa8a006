#!/bin/bash
# round 2, GPU call 42: one-CTA-per-ray kernel for rays of 2..4 passes: parity tests, config-5 timing + ncu, 1024/1536-sample sweeps A/B
set -u
O=gpurun_out/r2ap
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "multi_pass or config5 or stress or fused_mse or batched_poses or randomised" > $O/pytest_coop.log 2>&1; tail -5 $O/pytest_coop.log
timeout 600 python benchmarks/experiments/config5_step.py --poses 1024 > $O/config5.json 2> $O/config5.err
python -c "import json; d=json.load(open('$O/config5.json')); print('config5', d['ms_per_step'], d['gsamples_per_s'], d['hbm_frac_at_36B'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_bwd -s 1 -c 1 -o $O/prof_config5 \
    python benchmarks/experiments/config5_step.py --poses 256 --iters 1 > $O/ncu_config5.log 2>&1
ncu -i $O/prof_config5.ncu-rep --page raw --csv > $O/prof_config5.raw.csv 2>/dev/null
ncu -i $O/prof_config5.ncu-rep --page source --csv > $O/prof_config5.source.csv 2>/dev/null
rm -f $O/prof_config5.ncu-rep
# 256^3 sweeps with 1024- and 1536-sample rays (2 and 3 passes): one CTA per ray vs the pre-pass + multi-pass kernel
for c in 1 0; do
  DIFFUS_COOP=$c timeout 600 python - > $O/sweep_coop$c.json 2> $O/sweep_coop$c.err <<'PY'
import json, torch, bench
from diffus_b200 import PreparedVolume, ops, render_frames
from diffus_b200._lib import SAMPLER_TRILINEAR
from diffus_b200.phantoms import layered_phantom, pose_sweep
dev = torch.device("cuda", 0)
vol = layered_phantom(256, 0).to(dev)
pv = PreparedVolume(vol, "texture")
out = {}
for S in (1024, 1536, 2048):
    P = 512
    s, d = pose_sweep(P, 128, 256, seed=3)
    s, d = s.to(dev), d.to(dev)
    with torch.no_grad():
        tgt = render_frames(pv, s + torch.tensor([1.5, 0.0, -1.0], device=dev), d, S, 1e-4, sampler="trilinear")
    def step():
        ops.render_mse_impl(pv.volume, pv.bricks, [256] * 3, s, d, tgt, S, 0, 1e-4, SAMPLER_TRILINEAR, False, False, True, False)
    ms = bench.timed_steps(step, 3, 10, torch.cuda.synchronize)
    out[S] = {"ms": ms, "gsamples_per_s": P * 128 * S / ms / 1e6}
print(json.dumps(out))
PY
  cat $O/sweep_coop$c.json; tail -2 $O/sweep_coop$c.err
done
