"""Plane-operand contractions (vgtkb_split_bf16 + vgtkb_gemm_nt_presplit / vgtkb_gemm_tn_presplit, csrc/gemm_tc.cu template PRE:
the activation operand arrives as bf16 hi / lo planes and is loaded by TMA in the MMA's layout) against mode 3 of
vgtkb_gemm_nt / vgtkb_gemm_tn, whose arithmetic they reproduce (NT: bit-identical -- same operand split, same MMA order,
same chunked accumulation; TN: identical up to the order of the red.global.add partial sums).  These kernels are the
default path of vgtkb_inter_conv_* and of the intra-conv gather-GEMMs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(4096, 64, 1536), (1000, 128, 3072), (61440, 256, 6144), (300, 256, 64), (129, 72, 128)])
def test_presplit_contraction_matches_mode3(M, N, K):
    from equi_articulated_pose_b200 import lib, ops
    lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(dev)
    b = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    hi, lo = ops.split_bf16(a)
    assert torch.equal(hi.float(), a.to(torch.bfloat16).float())                       # round to nearest even
    assert torch.equal(lo.float(), (a - hi.float()).to(torch.bfloat16).float())
    ref = ops.gemm_nt(a, b, bias, mode=3)
    out = ops.gemm_nt_presplit(hi, lo, b, bias)
    assert torch.equal(out, ref)
    truth = (a.double() @ b.double().t() + bias.double()).float()
    assert float((out - truth).abs().max() / truth.abs().max()) < 2e-5


@pytest.mark.parametrize("M,N,R", [(64, 1536, 24576), (128, 3072, 12288), (256, 6144, 6144), (64, 128, 4096), (256, 192, 1000)])
def test_presplit_weight_gradient_matches_mode3(M, N, R):
    """C = a^T (b_hi + b_lo): the wide operand as planes; same arithmetic as mode 3 of vgtkb_gemm_tn up to the order of the
    red.global.add partial sums over the R splits (so: tight tolerance, not bit equality)."""
    from equi_articulated_pose_b200 import lib, ops
    lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M + N + R)
    a = torch.randn(R, M, generator=g).to(dev)
    b = torch.randn(R, N, generator=g).to(dev)
    hi, lo = ops.split_bf16(b)
    ref = ops.gemm_tn(a, b, mode=3)
    out = ops.gemm_tn_presplit(a, hi, lo)
    truth = (a.double().t() @ b.double()).float()
    scale = float(truth.abs().max())
    assert float((out - ref).abs().max()) < 2e-6 * scale
    assert float((out - truth).abs().max()) < 2e-5 * scale
