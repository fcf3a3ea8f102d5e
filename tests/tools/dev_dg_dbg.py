"""dev: data-gradient contraction dG = gy W of the inter conv (wide output, short reduction) under the timing experiment
bits (VGTKB_DBG set by the caller)."""
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

tag = f"DBG={os.environ.get('VGTKB_DBG','0')} TMA_EPI={os.environ.get('VGTKB_TMA_EPILOGUE','1')}"
out = []
for name, co, kc, rows in [("0.1", 64, 1536, 245760), ("1.1", 128, 3072, 122880), ("2.1", 256, 6144, 61440), ("1.0", 128, 1536, 122880)]:
    gy = torch.randn(rows, co, device=dev)
    hi, lo = ops.split_bf16(gy)
    wt = torch.randn(kc, co, device=dev)
    t = timeit(lambda: ops.gemm_nt_presplit(hi, lo, wt))
    out.append(f"{name}:{t*1e3:6.0f}us ({rows*kc*4/t/1e9:4.2f} TB/s)")
print(tag, "  ".join(out))
