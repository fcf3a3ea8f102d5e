"""Legacy S^2 ZPConv (BASELINE config 1b: IntraZPConv with 12 direction anchors, N=128, batch 1) against the
fixture produced by the reference's own module (tests/golden/make_golden.py -> ref_intrazp_small.npz)."""
import os

import numpy as np
import pytest
import torch

from tests.helpers import GOLD, rel_err


def _module():
    import equi_articulated_pose_b200 as eap
    eap.install()
    import vgtk.spconv as zptk
    torch.manual_seed(0)
    return zptk, zptk.IntraZPConv(dim_in=32, dim_out=32, kernel_size=3, aperture=1.0, sigma=0.1, anchor_nn=6, anchor_in=12)


def test_constants_match_reference():
    """anchors, angular kNN table and linear kernel weights are built on the host at construction."""
    g = np.load(os.path.join(GOLD, "ref_intrazp_small.npz"))
    _, m = _module()
    assert np.array_equal(m.anchor_out.numpy(), g["anchors"])
    assert np.array_equal(m.kernels.numpy(), g["kernels"])
    assert np.array_equal(m.intra_idx.numpy(), g["intra_idx"])
    assert np.allclose(m.intra_w.numpy(), g["intra_w"], atol=1e-6)
    assert tuple(m.basic_conv.W.shape) == g["W"].shape and tuple(m.basic_conv.bias.shape) == g["bias"].shape
    assert np.array_equal(m.basic_conv.W.detach().numpy(), g["W"])          # same init stream as the reference


@pytest.mark.gpu
def test_config1b_intra_zpconv_fwd_bwd():
    g = np.load(os.path.join(GOLD, "ref_intrazp_small.npz"))
    zptk, m = _module()
    dev = torch.device("cuda:0")
    m = m.to(dev)
    with torch.no_grad():
        m.basic_conv.W.copy_(torch.from_numpy(g["W"]))
        m.basic_conv.bias.copy_(torch.from_numpy(g["bias"]))
    f = torch.from_numpy(g["feats"]).to(dev).requires_grad_(True)
    out = m(zptk.SphericalPointCloud(torch.zeros(1, 3, 128, device=dev), f, None)).feats
    assert tuple(out.shape) == (1, 32, 128, 12)
    assert rel_err(out, torch.from_numpy(g["out"])) < 1e-4
    (out * torch.from_numpy(g["grad_out"]).to(dev)).sum().backward()
    assert rel_err(f.grad, torch.from_numpy(g["grad_feats"])) < 1e-4
    assert rel_err(m.basic_conv.W.grad, torch.from_numpy(g["grad_W"])) < 1e-4
    assert rel_err(m.basic_conv.bias.grad, torch.from_numpy(g["grad_bias"])) < 1e-4


@pytest.mark.gpu
def test_zpconv_slot_kernels_vs_torch():
    """the four vgtk.cuda.zpconv entry points against index arithmetic in torch (fp64)."""
    import equi_articulated_pose_b200 as eap
    eap.install()
    import vgtk.cuda.zpconv as Z
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(5)
    b, c, nq, p, a, k, ann = 2, 5, 17, 9, 12, 3, 4
    idx = torch.randint(0, nq, (b, p, a, k, ann), generator=gen, dtype=torch.int32)
    w = torch.rand(b, p, a, k, ann, generator=gen)
    feats = torch.randn(b, c, nq, a, generator=gen, dtype=torch.float64, requires_grad=True)
    ar = torch.arange(a).view(1, 1, a, 1, 1).expand_as(idx)
    br = torch.arange(b).view(b, 1, 1, 1, 1).expand_as(idx)
    gath = feats.permute(0, 3, 2, 1)[br, ar, idx.long()]                      # [b,p,a,k,ann,c]
    ref = (gath * w.double().unsqueeze(-1)).sum(4).permute(0, 4, 3, 1, 2)      # [b,c,k,p,a]
    go = torch.randn(ref.shape, generator=gen, dtype=torch.float64)
    (ref * go).sum().backward()
    out = Z.inter_zpconv_forward(idx.to(dev), w.to(dev), feats.detach().float().to(dev))
    assert rel_err(out, ref) < 1e-5
    gf = Z.inter_zpconv_backward(idx.to(dev), w.to(dev), go.float().to(dev), nq)
    assert rel_err(gf, feats.grad) < 1e-5
    # intra
    ain, aout = 12, 7
    idx2 = torch.randint(0, ain, (aout, ann), generator=gen, dtype=torch.int32)
    w2 = torch.rand(aout, k, ann, generator=gen)
    f2 = torch.randn(b, c, p, ain, generator=gen, dtype=torch.float64, requires_grad=True)
    ref2 = torch.einsum('bcpan,akn->bckpa', f2[..., idx2.long()], w2.double())
    go2 = torch.randn(ref2.shape, generator=gen, dtype=torch.float64)
    (ref2 * go2).sum().backward()
    out2 = Z.intra_zpconv_forward(idx2.to(dev), w2.to(dev), f2.detach().float().to(dev))
    assert rel_err(out2, ref2) < 1e-5
    gf2 = Z.intra_zpconv_backward(idx2.to(dev), w2.to(dev), go2.float().to(dev), ain)
    assert rel_err(gf2, f2.grad) < 1e-5
