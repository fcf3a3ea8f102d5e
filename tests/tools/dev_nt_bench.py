"""dev: plane-fed contractions (forward NT, weight-gradient TN, intra gather-GEMMs) on the backbone's shapes: time per
launch, bf16-equivalent TF/s, operand stream rate, and a sampled fp64 check of every result."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from equi_articulated_pose_b200 import lib, ops
lib.load()
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def check(name, got, a_rows, b_rows, idx):
    """got[idx] vs fp64 a_rows @ b_rows.T"""
    ref = a_rows.double() @ b_rows.double().t()
    err = (got[idx].double() - ref).abs().max().item() / ref.abs().max().item()
    flag = "OK " if err < 3e-5 else "BAD"
    print(f"    {flag} {name}: max err / max ref = {err:.2e}")
    return err < 3e-5


ok = True
torch.manual_seed(0)
shapes = [("inter 0.1", 64, 1536, 245760), ("inter 1.0", 128, 1536, 122880), ("inter 1.1", 128, 3072, 122880),
          ("inter 2.0", 256, 3072, 61440), ("inter 2.1", 256, 6144, 61440), ("inter 3.0", 256, 6144, 30720)]
for name, co, kc, rows in shapes:
    g = torch.randn(rows, kc, device=dev)
    w = torch.randn(co, kc, device=dev) / kc ** 0.5
    gy = torch.randn(rows, co, device=dev)
    hi, lo = ops.split_bf16(g)
    idx = torch.randint(0, rows, (64,), device=dev)
    g_s = g[idx].clone()
    jdx = torch.randint(0, kc, (64,), device=dev)
    g_c = g[:, jdx].clone()
    del g
    t_nt = timeit(lambda: ops.gemm_nt_presplit(hi, lo, w))
    out = ops.gemm_nt_presplit(hi, lo, w)
    ok &= check("nt", out, g_s, w, idx)
    t_tn = timeit(lambda: ops.gemm_tn_presplit(gy, hi, lo))
    dw = ops.gemm_tn_presplit(gy, hi, lo)               # [co, kc]
    ok &= check("tn", dw.t().contiguous(), g_c.t().contiguous(), gy.t().contiguous(), jdx)
    flops = 2.0 * co * kc * rows * 3
    print(f"{name}: nt {t_nt*1e3:7.1f} us ({flops/t_nt/1e9:6.0f} TF/s, G {rows*kc*4/t_nt/1e9:5.2f} TB/s)   "
          f"tn {t_tn*1e3:7.1f} us ({flops/t_tn/1e9:6.0f} TF/s, G {rows*kc*4/t_tn/1e9:5.2f} TB/s)")
    del hi, lo, gy, out, dw

import equi_articulated_pose_b200 as pkg
pkg.install()
import vgtk.so3conv as sptk
for name, pts, c in [("intra C64", 4096, 64), ("intra C128", 2048, 128), ("intra C256", 1024, 256), ("intra C256 s", 512, 256)]:
    conv = sptk.IntraSO3Conv(c, c).to(dev)
    t, inv, _ = conv.tables()
    x = torch.randn(pts, 60, c, device=dev)
    y = torch.randn(pts * 60, c, device=dev)
    w = torch.randn(c, 12 * c, device=dev) / (12 * c) ** 0.5
    xh, xl = ops.split_bf16(x)
    yh, yl = ops.split_bf16(y)
    t_nt = timeit(lambda: ops.gather_gemm_nt_planes(xh, xl, t, w))
    out = ops.gather_gemm_nt_planes(xh, xl, t, w).reshape(pts, 60, c)
    # reference on a few (point, anchor) rows
    pi = torch.randint(0, pts, (16,), device=dev)
    ai = torch.randint(0, 60, (16,), device=dev)
    rows_ref = torch.stack([x[p, t.view(60, 12)[a].long()].reshape(-1) for p, a in zip(pi.tolist(), ai.tolist())])
    ref = rows_ref.double() @ w.double().t()
    got = out[pi, ai].double()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    print(f"    {'OK ' if err < 3e-5 else 'BAD'} gather nt: {err:.2e}")
    ok &= err < 3e-5
    t_tn = timeit(lambda: ops.gather_gemm_tn_planes(xh, xl, t, y, yh, yl))
    flops = 2.0 * pts * 60 * 12 * c * c * 3
    print(f"{name}: nt {t_nt*1e3:7.1f} us ({flops/t_nt/1e9:6.0f} TF/s)   tn {t_tn*1e3:7.1f} us ({flops/t_tn/1e9:6.0f} TF/s)")
print("ALL OK" if ok else "FAILURES")
