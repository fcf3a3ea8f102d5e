"""Developer check (GPU): forward grouping, FFMA (mode 0) vs warp-MMA bf16x3 (mode 3): error vs fp64, time per layer."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from equi_articulated_pose_b200 import ops, so3_constants as C
from equi_articulated_pose_b200.lib import call, ptr
from oracle import so3 as O
import numpy as np

dev = torch.device("cuda:0")
anchors = torch.from_numpy(np.ascontiguousarray(C.get_anchors(60))).float()
params = O.backbone_params(input_num=1024)
layers = [l['args'] for blk in params['backbone'] for l in blk['blocks']] if isinstance(params, dict) and 'backbone' in params else None

def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

g = torch.Generator().manual_seed(0)
cases = [(8, 512, 512, 16, 64, 0.2828, 0.04), (8, 512, 256, 32, 64, 0.4, 0.08), (8, 256, 256, 16, 128, 0.4, 0.08),
         (8, 256, 128, 32, 128, 0.5657, 0.16), (8, 128, 128, 16, 256, 0.5657, 0.16), (8, 128, 64, 32, 256, 0.8, 0.32),
         (2, 100, 100, 9, 32, 0.4, 0.08), (2, 200, 50, 20, 96, 0.5, 0.1)]
kp = torch.from_numpy(np.ascontiguousarray(C.kernel_points(24) if hasattr(C, 'kernel_points') else C.get_kernel_points())).float() if False else None
for (b, n, p, nn, ci, radius, sigma) in cases:
    xyz = O.synthetic_cloud(b, n, 7).permute(0, 2, 1).contiguous().to(dev)
    sxyz = xyz[:, :, :p].contiguous()
    idx = ops.ball_query(sxyz, xyz, radius, nn)
    base = torch.randn(24, 3, generator=g); base = base / base.norm(dim=1, keepdim=True) * torch.rand(24, 1, generator=g) * 0.7 * radius
    rk = torch.einsum('aij,kj->aki', anchors, base).contiguous().to(dev)
    feats = torch.randn(b, n, 60, ci, generator=g).to(dev)
    outs = {}
    for mode in (0, 3):
        gg = torch.empty(b, p, 60, 24 * ci, device=dev)
        fn = lambda: call("vgtkb_inter_group_forward", dev, b, n, p, nn, 60, 24, ci, ptr(xyz), ptr(sxyz), ptr(idx), ptr(rk),
                          float(sigma), ptr(feats), ptr(gg), mode)
        ms = t(fn)
        outs[mode] = (gg, ms)
    dg = torch.randn(b, p, 60, 24 * ci, generator=g).to(dev)
    bouts = {}
    for mode in (0, 3):
        gx = torch.zeros(b, n, 60, ci, device=dev)
        fn = lambda: call("vgtkb_inter_group_backward", dev, b, n, p, nn, 60, 24, ci, ptr(xyz), ptr(sxyz), ptr(idx), ptr(rk),
                          float(sigma), ptr(dg), ptr(gx), mode)
        ms = t(fn)
        gx.zero_(); fn(); torch.cuda.synchronize()
        bouts[mode] = (gx.clone(), ms)
    berr = float((bouts[3][0].double() - bouts[0][0].double()).abs().max() / bouts[0][0].abs().max())
    print(f"   bwd: ffma {bouts[0][1]:.3f} ms, mma {bouts[3][1]:.3f} ms, relerr {berr:.2e}")
    ref = outs[0][0].double()
    err = float((outs[3][0].double() - ref).abs().max() / ref.abs().max())
    gb = outs[0][0].numel() * 4 / 1e9
    print(f"b{b} n{n} p{p} nn{nn} ci{ci}: ffma {outs[0][1]:.3f} ms, mma {outs[3][1]:.3f} ms ({gb / outs[3][1] * 1e3:.0f} GB/s write), "
          f"relerr(mma vs ffma) {err:.2e}", flush=True)
