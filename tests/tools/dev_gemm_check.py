"""Developer check (GPU): tcgen05 GEMMs vs fp64, error + CUDA-event timing per shape."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from equi_articulated_pose_b200 import ops

def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
which = sys.argv[1] if len(sys.argv) > 1 else "nt"
shapes = [(256, 64, 32), (128, 64, 64), (1000, 24, 192), (4096, 128, 768), (8 * 512 * 60, 64, 24), (8 * 512 * 60, 64, 1536),
          (8 * 128 * 60, 256, 6144), (8 * 128 * 60, 256, 3072), (8*128*60, 6144, 256), (999, 130, 36), (300, 257, 40)]
for (M, N, K) in shapes:
    A = torch.randn(M, K, generator=g).to(dev); B = torch.randn(N, K, generator=g).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    if which == "nt":
        ref = (A[:2048].double() @ B.double().t() + bias.double())
        for mode in (1, 3, 2):
            out = ops.gemm_nt(A, B, bias, mode=mode)
            err = float((out[:2048].double() - ref).abs().max() / ref.abs().max())
            ms = t(lambda: ops.gemm_nt(A, B, bias, mode=mode))
            print(f"nt M={M} N={N} K={K} mode={mode} relerr={err:.2e} {ms:.3f} ms {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    else:
        D = torch.randn(M, N, generator=g).to(dev)
        ref = D.double().t() @ A.double()
        for mode in (1, 3, 2):
            out = ops.gemm_tn(D, A, mode=mode)
            err = float((out.double() - ref).abs().max() / ref.abs().max())
            ms = t(lambda: ops.gemm_tn(D, A, mode=mode))
            print(f"tn R={M} M={N} N={K} mode={mode} relerr={err:.2e} {ms:.3f} ms {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
