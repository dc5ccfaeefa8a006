"""Timing experiment: the config-4 render + scatter kernel with the gradient volume (a) separate, (b) aliased onto a
second copy that is NOT gathered (control), (c) aliased onto the gathered Z bricks themselves (atomics hit the lines the
gathers just fetched; corrupts Z, timing only), (d) no volume gradient at all."""
import ctypes as C, sys, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
from diffus_b200 import ops, _lib, PreparedVolume, render_frames
from diffus_b200._lib import DiffusRenderBwdArgs
from diffus_b200.phantoms import intensity_to_impedance, mri_phantom, pose_sweep
dev = torch.device("cuda:0")
lib = _lib.load()
vol = intensity_to_impedance(mri_phantom(256, "t2")).to(dev)
pv = PreparedVolume(vol, "brick")
P = 4096
s, d = pose_sweep(P, 128, 256, seed=2)
s, d = s.to(dev), d.to(dev)
for sampler, name in ((0, "nearest"), (1, "trilinear")):
    with torch.no_grad():
        tgt = render_frames(pv, s + 0.5, d, 512, 1e-4, sampler=name)
    for mode in ("separate", "alias_Z", "none"):
        bricks = pv.bricks.clone()
        b = DiffusRenderBwdArgs()
        ops._fill_render_args(b.fwd, vol, bricks, [256] * 3, s, d, 512, 0, 1e-4, sampler, False)
        n = tgt.numel()
        loss = torch.empty((1,), device=dev)
        gvol = torch.zeros_like(bricks)
        b.fwd.frame = None; b.fwd.seg_prefix = None; b.grad_frame = None
        b.grad_volume = {"separate": gvol.data_ptr(), "alias_Z": bricks.data_ptr(), "none": None}[mode]
        b.grad_sources = None; b.grad_directions = None
        b.target, b.grad_scale, b.loss_scale, b.loss = tgt.data_ptr(), 2.0 / n, 1.0 / n, loss.data_ptr()
        wbytes = lib.diffus_render_bwd_workspace_bytes(C.byref(b))
        ws = torch.empty((max(wbytes, 1),), dtype=torch.uint8, device=dev)
        b.workspace, b.workspace_bytes = ws.data_ptr(), wbytes
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        for _ in range(2):
            _lib.check(lib.diffus_render_backward(C.byref(b), st), "bwd")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            _lib.check(lib.diffus_render_backward(C.byref(b), st), "bwd")
        e1.record(); torch.cuda.synchronize()
        print(name, mode, f"{e0.elapsed_time(e1) / 5:.3f} ms per 4096 poses", flush=True)
