"""dev: inter conv forward / data-gradient timings on the backbone's stride-1 and strided layer shapes under the timing
experiment bits (VGTKB_DBG_GROUP, VGTKB_DBG set by the caller)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from equi_articulated_pose_b200 import lib, ops, so3_constants
lib.load()
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, n=5):
    for _ in range(2): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]

torch.manual_seed(0)
tag = f"GROUP={os.environ.get('VGTKB_DBG_GROUP','0')} GEMM={os.environ.get('VGTKB_DBG','0')}"
out = []
anchors = torch.from_numpy(so3_constants.get_anchors()).float().to(dev)
kp = torch.from_numpy(so3_constants.scaled_kernel_points(0.3)).float().to(dev)      # [24,3]
rk = torch.einsum('aij,kj->aki', anchors, kp).contiguous()                           # [60,24,3]
for name, b, n, p, nn, ci, co in [("0.1", 8, 512, 512, 16, 64, 64), ("1.0", 8, 512, 256, 32, 64, 128), ("2.1", 8, 128, 128, 16, 256, 256)]:
    xyz = torch.randn(b, 3, n, device=dev) * 0.3
    sxyz = xyz[:, :, :p].contiguous()
    idx = torch.randint(0, n, (b, p, nn), device=dev, dtype=torch.int32)
    feats = torch.randn(b, n, 60, ci, device=dev, requires_grad=True)
    w = (torch.randn(co, 24 * ci, device=dev) / (24 * ci) ** 0.5).requires_grad_(True)
    y = ops.InterConvFn.apply(feats, w, xyz, sxyz, idx, rk, 0.05)
    gy = torch.randn_like(y)
    t_f = timeit(lambda: ops.InterConvFn.apply(feats, w, xyz, sxyz, idx, rk, 0.05))
    def bwd():
        feats.grad = None; w.grad = None
        y.backward(gy, retain_graph=True)
    t_b = timeit(bwd)
    out.append(f"{name}: fwd {t_f*1e3:6.0f} bwd {t_b*1e3:6.0f}")
    del y, gy, feats, w
print(tag, " | ".join(out))
