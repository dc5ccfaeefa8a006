#!/usr/bin/env python
"""Benchmark of the B-mode renderer hot path (BASELINE.json metric: frames/s and Gsamples/s, forward+backward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--poses P]

Workload (the batched pose sweep, with config 2's loss): P probe
poses per GPU x 128 rays x 512 samples over one synthetic MRI-shaped 256^3 impedance
volume, trilinear sampler, one step = render forward -> MSE against target frames ->
backward to d loss/d source and d loss/d directions of every pose.  Poses shard across
ranks with a full volume replica each and no data-path collective (weak scaling).

One JSON line on stdout (rank 0).  `value` is device-timed with everything resident in HBM;
`e2e` goes through the public API with the step's poses arriving from pinned host memory
and the loss + pose gradients read back to the host inside the timed region (the volume and
the target frames are uploaded once: they are the scene and the dataset of a pose-recovery
run, the poses are what changes per step).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

# stdout carries exactly one JSON line: NCCL's banner ("NCCL version ...", printed to its debug file -- stdout by default --
# when the box sets NCCL_DEBUG) and any other NCCL log go to stderr.  Set before torch / NCCL read the environment.
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import torch  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bmode_frames_per_s_fwd_bwd"
UNIT = "frames/s"
N_RAYS, N_SAMPLES, VOL_N = 128, 512, 256
ALPHA = 1e-4
OPENING_ANGLE = math.radians(60.0)     # pose_sweep's fan aperture
BYTES_PER_SAMPLE_FWD = 36      # 8 trilinear gathers x 4 B + 4 B frame write
BYTES_PER_SAMPLE_BWD = 36      # re-gather 8 x 4 B + read d loss/d frame 4 B (pose gradients only)
BYTES_PER_SAMPLE_FUSED = 36    # fused step: 8 gathers x 4 B + 4 B target read (frames are not written)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 100 ms while the GPU is under the benchmark load."""
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.mark = index, [], None, 0

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def begin_region(self):
        self.mark = len(self.rows)

    def count(self):
        return len(self.rows) - self.mark

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[self.mark:]:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(P, layout):
    """DRAM bytes per launch of the fused kernel from the committed ncu capture, if it is the same workload."""
    try:
        with open(os.path.join(ROOT, "profiles", "dram_traffic.json")) as f:
            entries = json.load(f)["render_bwd_kernel_fused_mse"]
        for t in (entries if isinstance(entries, list) else [entries]):
            if (t["poses"], t["rays"], t["samples"], t["layout"]) == (P, N_RAYS, N_SAMPLES, layout):
                return t["bytes_per_launch"]
    except Exception:
        pass
    return None


def ncu_counter(P, layout, key):
    """Another per-launch counter of the committed ncu capture (same file as ncu_traffic)."""
    try:
        with open(os.path.join(ROOT, "profiles", "dram_traffic.json")) as f:
            entries = json.load(f)["render_bwd_kernel_fused_mse"]
        for t in (entries if isinstance(entries, list) else [entries]):
            if (t["poses"], t["rays"], t["samples"], t["layout"]) == (P, N_RAYS, N_SAMPLES, layout):
                return t.get(key)
    except Exception:
        pass
    return None


def build_scene(device, n_poses, seed, return_params=False):
    from diffus_b200.phantoms import intensity_to_impedance, mri_phantom, pose_sweep
    vol = intensity_to_impedance(mri_phantom(VOL_N, "t1", seed=0))
    sources, dirs, median, hint = pose_sweep(n_poses, N_RAYS, VOL_N, seed=seed, return_params=True)
    if return_params:
        return vol, sources, dirs, median, hint
    return vol, sources, dirs


def run_ours(args):
    import torch.distributed as dist
    from diffus_b200 import PreparedVolume, ops, render_frames, render_mse_loss
    from diffus_b200.graphs import GraphedFanPoseStep, GraphedPoseStep
    from diffus_b200._lib import SAMPLER_TRILINEAR

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout must carry exactly one JSON line, and NCCL prints its banner ("NCCL version ...") to stdout when the
        # communicator is created: point fd 1 at stderr while the process group comes up (eagerly, plus one collective)
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    P = args.poses
    vol_h, src_h, dir_h, med_h, hint_h = build_scene(dev, P, seed=1000 + rank, return_params=True)   # every rank: its own pose shard
    vol = PreparedVolume(vol_h.to(dev), args.layout) if args.layout != "linear" else vol_h.to(dev)
    src_pin, dir_pin = src_h.pin_memory(), dir_h.pin_memory()
    src_d, dir_d = src_h.to(dev), dir_h.to(dev)
    with torch.no_grad():
        shift = torch.tensor([1.5, 0.0, -1.0], device=dev)
        target = render_frames(vol, src_d + shift, dir_d, N_SAMPLES, ALPHA, 0, sampler="trilinear")
    n_elem = target.numel()
    samples_per_step = P * N_RAYS * N_SAMPLES
    bricks = vol.bricks if isinstance(vol, PreparedVolume) else None
    vol_t = vol.volume if isinstance(vol, PreparedVolume) else vol
    dims = list(vol_t.shape)

    # ---- device-resident step: the fused forward + MSE + backward op (what autograd calls) ----
    def step_device(ev=None):
        if ev:
            ev[0].record()
        loss, _, _, gs, gd = ops.render_mse(vol_t, bricks, dims, src_d, dir_d, target, N_SAMPLES, 0, ALPHA,
                                            SAMPLER_TRILINEAR, False, False, True, False)
        if ev:
            ev[1].record()
        return loss, gs, gd

    # ---- end-to-end step through the public API: host poses in, loss + gradients out ----
    out_src = torch.empty((P, 3), dtype=torch.float32).pin_memory()
    out_dir = torch.empty((P, N_RAYS, 3), dtype=torch.float32).pin_memory()
    out_loss = torch.empty((), dtype=torch.float32).pin_memory()

    # A pose sweep repeats the same shapes every step: the user-facing call for that is a CUDA-graph wrapper of the
    # fused step (diffus_b200.graphs); eager autograd costs ~0.15 ms more.  "fan" (default) drives the step with the
    # pose PARAMETERS north_star names -- source, median direction, in-plane hint, aperture (9 floats per pose in, 9
    # gradient floats + the loss out; the fans are built and their gradient folded back on the device) -- "graph" and
    # "eager" pass explicit (P,R,3) direction tensors like the reference's plot_beam_frame.
    gstep = fstep = None
    poses_pin = torch.stack([src_h, med_h, hint_h]).contiguous().pin_memory()       # (3, P, 3): one H2D copy per step
    out_pin = torch.empty((9 * P + 1,), dtype=torch.float32).pin_memory()            # gradients + loss: one D2H copy
    if args.e2e == "graph":
        gstep = GraphedPoseStep(vol, target, N_RAYS, N_SAMPLES, ALPHA)
    elif args.e2e == "fan":
        fstep = GraphedFanPoseStep(vol, target, N_RAYS, N_SAMPLES, OPENING_ANGLE, ALPHA)

    def step_e2e():
        if fstep is not None:
            out_pin.copy_(fstep.packed(poses_pin), non_blocking=True)   # pose parameters in, gradients + loss out
        elif gstep is not None:
            loss, gs, gd = gstep(src_pin, dir_pin)              # H2D of the poses into the graph's static inputs
            out_loss.copy_(loss, non_blocking=True)
            out_src.copy_(gs, non_blocking=True)
            out_dir.copy_(gd, non_blocking=True)
        else:
            s = src_pin.to(dev, non_blocking=True).requires_grad_(True)
            d = dir_pin.to(dev, non_blocking=True).requires_grad_(True)
            loss = render_mse_loss(vol, s, d, target, N_SAMPLES, ALPHA, 0, sampler="trilinear")
            loss.backward()
            out_loss.copy_(loss.detach(), non_blocking=True)
            out_src.copy_(s.grad, non_blocking=True)
            out_dir.copy_(d.grad, non_blocking=True)
        torch.cuda.current_stream().synchronize()        # the caller needs the numbers on the host
        return out_loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W, K = max(args.warmup, 3), args.steps
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                  # nvidia-smi needs a moment to come up: start before the warm-up
    for _ in range(W):
        step_device()
    barrier()
    sampler.begin_region()
    launches0 = ops.launch_count()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(K)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(K):
        loss, gs, gd = step_device(evs[i])
    t1.record()
    barrier()
    launches = ops.launch_count() - launches0
    ms_total = t0.elapsed_time(t1)
    step_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in evs)     # the fused kernel (+ 2 tiny reductions)

    for _ in range(W):
        step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step_e2e()
    e1.record()
    barrier()
    e2e_ms_total = e0.elapsed_time(e1)
    if rank == 0:
        # a short run can end before nvidia-smi delivers a sample: keep the same load on (untimed) until it has
        t_hold = time.perf_counter()
        while sampler.proc is not None and sampler.count() < 3 and time.perf_counter() - t_hold < 4.0:
            for _ in range(20):
                step_device()
            torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None

    times = torch.tensor([ms_total, e2e_ms_total, step_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms_total, step_ms = times.tolist()
    ms_per_step = ms_total / K
    frames_per_s = world * P / (ms_per_step * 1e-3)
    e2e_frames_per_s = world * P / (e2e_ms_total / K * 1e-3)

    if rank == 0:
        peak, peak_src = load_peaks()
        dom = "render_bwd_kernel<trilinear, pose_grad, LOSS_MSE> (fused forward + MSE + backward)"
        dom_bytes = samples_per_step * BYTES_PER_SAMPLE_FUSED
        achieved = dom_bytes / (step_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": frames_per_s, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "gsamples_per_s": world * samples_per_step / (ms_per_step * 1e-3) / 1e9,
            "config": {
                "workload": f"config3 pose sweep fwd+bwd: {P} poses/GPU x {N_RAYS} rays x {N_SAMPLES} samples, "
                            f"{VOL_N}^3 MRI-shaped impedance volume, trilinear, MSE vs target frames, "
                            "gradients to every pose's source and directions",
                "poses_per_gpu": P, "rays": N_RAYS, "samples": N_SAMPLES, "volume": f"{VOL_N}^3 f32",
                "volume_layout": args.layout, "parallelism": f"pose-sharded x{world}, volume replicated",
                "l2": "no explicit flush: inputs exceed L2 -- every step streams its own 268 MB of target frames per "
                      "1024 poses (>> 126 MB L2, marked evict-first); the 64 MiB volume is meant to stay L2-resident",
            },
            "e2e": {"value": e2e_frames_per_s, "unit": UNIT,
                    "h2d_bytes_per_step": int(poses_pin.numel() * 4) if args.e2e == "fan"
                                          else int(src_pin.numel() * 4 + dir_pin.numel() * 4),
                    "d2h_bytes_per_step": int(out_pin.numel() * 4) if args.e2e == "fan"
                                          else int(out_src.numel() * 4 + out_dir.numel() * 4 + 4),
                    "ms_per_step": e2e_ms_total / K,
                    "api": {"fan": "GraphedFanPoseStep (CUDA-graph replay: fans from pose parameters -> fused step -> gradient "
                                   "folded back onto the pose parameters)",
                            "graph": "GraphedPoseStep (CUDA-graph replay of the fused step, explicit (P,R,3) directions)",
                            "eager": "render_mse_loss + loss.backward() (eager autograd, explicit directions)"}[args.e2e],
                    "note": "public API call per step; pose parameters (fan: source, median direction, in-plane hint; else "
                            "source + explicit directions) from pinned host memory each step, loss and pose gradients copied "
                            "back to pinned host memory, stream synchronised every step; volume and target frames resident"},
            "gpu_launches": launches,
            "kernels_ms": {"fused_step(render_bwd_kernel+reduce_rays+reduce_sum)": step_ms},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(P, args.layout), "peak_source": peak_src,
                         "bytes_per_sample": BYTES_PER_SAMPLE_FUSED,
                         "bytes_note": "SURVEY 8(d) counts 72 B/sample for forward+backward done as two passes "
                                       "(2 x (8 gathers x 4 B + 4 B)); the fused kernel gathers once and reads the "
                                       "target once, so its own algorithmic traffic is 36 B/sample",
                         "frac_vs_two_pass_bytes": samples_per_step * 72 / (step_ms * 1e-3) / 1e9 / peak},
            "clocks": clocks,
            "loss": float(loss),
            "loss_e2e": float(out_pin[-1]) if args.e2e == "fan" else float(out_loss),
        }
        # the L2 gather roof from our own microbenchmark (random 32-byte sectors over a 64 MiB buffer), and
        # the same over 512 MiB (HBM random sectors: what a volume copy that does not fit L2 would run at)
        try:
            l2 = ops.gather_probe(64, device=dev)
            hbm = ops.gather_probe(512, device=dev)
            l1_miss_sectors = ncu_counter(P, args.layout, "l2_read_sectors_per_launch")
            roof = {"what": "random 32-byte-sector reads, 8 loads in flight per thread, 148 x 8 x 256 threads",
                    "l2_resident_64MiB": {"sectors_per_s": l2["sectors_per_s"], "gb_per_s": l2["gb_per_s"]},
                    "hbm_512MiB": {"sectors_per_s": hbm["sectors_per_s"], "gb_per_s": hbm["gb_per_s"]}}
            if l1_miss_sectors:
                rate = l1_miss_sectors / (step_ms * 1e-3)
                roof["kernel_l2_sector_reads_per_s"] = rate
                roof["frac_of_l2_gather_roof"] = rate / l2["sectors_per_s"]
                roof["note"] = ("the kernel's L2 -> L1 sector reads per launch are the ncu count "
                                "(lts__t_sectors_srcunit_tex_op_read) of the committed capture; 83 % of its gather sectors "
                                "hit L1 and never reach L2")
            line["roofline"]["gather_roof"] = roof
        except Exception as exc:                       # the probe must never take the headline number down with it
            line["roofline"]["gather_roof"] = {"error": str(exc)}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.cpu_rays)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_sample_step(vol, src, dirs, target):
    """One bounded sample of the workload on the CPU with the reference's own algorithm (dense solves)."""
    from oracle import port
    s = src.clone().requires_grad_(True)
    d = dirs.clone().requires_grad_(True)
    _, _, _, f = port.plot_beam_frame(vol, s, d, N_SAMPLES, ALPHA, sampler="trilinear", propagation="dense")
    loss = torch.nn.functional.mse_loss(f, target)
    loss.backward()
    return float(loss.detach())


def cpu_scene(n_rays):
    from oracle import port
    vol, src, dirs = build_scene(None, 1, seed=1000)
    lo = N_RAYS // 2 - n_rays // 2
    d = dirs[0, lo:lo + n_rays].contiguous()
    s = src[0]
    with torch.no_grad():
        _, _, _, target = port.plot_beam_frame(vol, s + torch.tensor([1.5, 0.0, -1.0]), d, N_SAMPLES, ALPHA,
                                               sampler="trilinear", propagation="closed_form")
    return vol, s, d, target.float()


def bounded_cpu_rays(n_rays):
    """The dense per-depth solves keep ~2 GB of autograd state per ray at 512 samples: stay under a quarter of free RAM."""
    try:
        import psutil
        return max(1, min(n_rays, int(psutil.virtual_memory().available / 2**30 / 4 / 2)))
    except Exception:
        return min(n_rays, 4)


def cpu_baseline(n_rays):
    """Reference algorithm (oracle port: one dense solve per depth + torch autograd) on a bounded sample."""
    n_rays = bounded_cpu_rays(n_rays)
    vol, s, d, target = cpu_scene(n_rays)
    t = time.perf_counter()
    cpu_sample_step(vol, s, d, target)
    dt = time.perf_counter() - t
    frames = n_rays / N_RAYS
    return {"value": frames / dt, "unit": UNIT, "cores": torch.get_num_threads(), "host_cpus": os.cpu_count(),
            "kind": "port",
            "sample": f"{n_rays} of {N_RAYS} rays of one pose, all {N_SAMPLES} samples, forward+backward with the "
                      f"reference's per-depth dense solves (oracle/port.py propagation='dense') in {dt:.1f} s; "
                      "frames/s = (rays/128)/time", "seconds": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if os.environ.get("OMP_NUM_THREADS") and not os.environ.get("DIFFUS_REF_CHILD"):
        # torchrun pins OMP_NUM_THREADS=1; the reference arm is entitled to every host thread, and
        # torch.set_num_threads() after start-up makes the batched LAPACK solves crawl, so re-exec clean
        env = {k: v for k, v in os.environ.items() if k not in ("OMP_NUM_THREADS", "MKL_NUM_THREADS")}
        env["DIFFUS_REF_CHILD"] = "1"
        out = subprocess.run([sys.executable, os.path.abspath(__file__), *sys.argv[1:]], env=env, capture_output=True, text=True)
        sys.stdout.write(out.stdout)
        sys.stderr.write(out.stderr[-2000:])
        sys.stdout.flush()
        return
    K, W = args.steps, args.warmup
    args.cpu_rays = bounded_cpu_rays(args.cpu_rays)
    vol, s, d, target = cpu_scene(args.cpu_rays)
    for _ in range(min(W, 1)):
        cpu_sample_step(vol, s, d, target)
    # every step is a bounded sample; the whole arm is bounded too (the GPU arm's default K is sized for a
    # 1 ms step, the CPU sample takes seconds): stop after K steps or ~100 s, whichever comes first
    t = time.perf_counter()
    done = 0
    while done < K and (done == 0 or time.perf_counter() - t < 100.0):
        cpu_sample_step(vol, s, d, target)
        done += 1
    dt = (time.perf_counter() - t) / done
    K_req, K = K, done
    value = (args.cpu_rays / N_RAYS) / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": K, "steps_requested": K_req, "warmup": min(W, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"config3 pose sweep fwd+bwd sample: {args.cpu_rays} of {N_RAYS} rays x {N_SAMPLES} samples of one "
                               f"pose per step, {VOL_N}^3 volume, trilinear, MSE vs target frame (CPU, reference algorithm)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{args.cpu_rays}/{N_RAYS} rays of one frame per step, dense per-depth solves + autograd"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--poses", type=int, default=1024, help="poses per GPU per step")
    ap.add_argument("--layout", default="brick", choices=["linear", "brick", "quad", "texture"])
    ap.add_argument("--cpu-rays", type=int, default=16, help="rays in the bounded CPU sample (~2 GB and ~0.5 s each)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e", default="fan", choices=["fan", "graph", "eager"], help="public API used by the end-to-end loop")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
