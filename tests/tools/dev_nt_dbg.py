"""dev: decompose the plane-fed NT contraction: full kernel / no MMAs / no loads / neither (VGTKB_DBG, set by the caller)."""
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

tag = f"DBG={os.environ.get('VGTKB_DBG','0')} CHUNK={os.environ.get('VGTKB_CHUNK_KB','-')}"
out = []
for name, co, kc, rows in [("0.1", 64, 1536, 245760), ("1.1", 128, 3072, 122880), ("2.1", 256, 6144, 61440)]:
    hi = torch.zeros(rows, kc, dtype=torch.bfloat16, device=dev); lo = torch.zeros_like(hi)
    w = torch.randn(co, kc, device=dev)
    t = timeit(lambda: ops.gemm_nt_presplit(hi, lo, w))
    out.append(f"{name}:{t*1e3:6.0f}us")
    del hi, lo
import equi_articulated_pose_b200 as pkg
pkg.install()
import vgtk.so3conv as sptk
for name, pts, c in [("iC64", 4096, 64), ("iC128", 2048, 128)]:
    conv = sptk.IntraSO3Conv(c, c).to(dev)
    t, inv, _ = conv.tables()
    xh = torch.zeros(pts, 60, c, dtype=torch.bfloat16, device=dev); xl = torch.zeros_like(xh)
    w = torch.randn(c, 12 * c, device=dev)
    tt = timeit(lambda: ops.gather_gemm_nt_planes(xh, xl, t, w))
    out.append(f"{name}:{tt*1e3:6.0f}us")
print(tag, "  ".join(out))
