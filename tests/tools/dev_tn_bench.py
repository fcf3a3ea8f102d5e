"""dev: weight-gradient (TN) contraction timings on the backbone's shapes, old vs new work decomposition (same process)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from equi_articulated_pose_b200 import lib, ops
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

shapes = [("inter 0.1", 64, 1536, 245760), ("inter 1.0", 128, 1536, 122880), ("inter 1.1", 128, 3072, 122880),
          ("inter 2.0", 256, 3072, 61440), ("inter 2.1", 256, 6144, 61440), ("inter 3.0", 256, 6144, 30720)]
for name, co, kc, rows in shapes:
    gy = torch.randn(rows, co, device=dev)
    g = torch.randn(rows, kc, device=dev)
    hi, lo = ops.split_bf16(g)
    del g
    res = {}
    for mode in ("0", "1"):
        os.environ["VGTKB_TN_DECOMP"] = mode
        res[mode] = timeit(lambda: ops.gemm_tn_presplit(gy, hi, lo))
    flops = 2.0 * co * kc * rows * 3
    print(f"{name}: old {res['0']*1e3:7.1f} us  new {res['1']*1e3:7.1f} us  ({flops/res['1']/1e9:6.0f} TF/s bf16-equivalent, G read {rows*kc*4/res['1']/1e9:5.2f} TB/s)")
    del hi, lo, gy
table = None
for name, pts, c in [("intra C64", 4096, 64), ("intra C128", 2048, 128), ("intra C256", 1024, 256), ("intra C256 s", 512, 256)]:
    import equi_articulated_pose_b200 as pkg
    pkg.install()
    import vgtk.so3conv as sptk
    conv = sptk.IntraSO3Conv(c, c).to(dev)
    t, inv, _ = conv.tables()
    x = torch.randn(pts, 60, c, device=dev)
    y = torch.randn(pts * 60, c, device=dev)
    xh, xl = ops.split_bf16(x)
    yh, yl = ops.split_bf16(y)
    res = {}
    for mode in ("0", "1"):
        os.environ["VGTKB_TN_DECOMP"] = mode
        res[mode] = timeit(lambda: ops.gather_gemm_tn_planes(xh, xl, t, y, yh, yl))
    flops = 2.0 * pts * 60 * 12 * c * c * 3
    print(f"{name}: old {res['0']*1e3:7.1f} us  new {res['1']*1e3:7.1f} us  ({flops/res['1']/1e9:6.0f} TF/s bf16-equivalent)")
