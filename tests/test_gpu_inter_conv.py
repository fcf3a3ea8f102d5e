"""vgtkb_inter_conv_forward / _backward (InterSO3Conv in one call per direction, grouped tensor only as bf16 operand
planes) against (1) the two-step path vgtkb_inter_group_* + vgtkb_gemm_* it replaces -- forward bit-identical: same
fp32 grouping accumulators, same hi/lo split, same MMA order -- and (2) the fp64 evaluation of the reference's
expressions (oracle/so3.py: inter_so3conv_grouping_anchor + inter_zpconv_grouping_naive + BasicSO3Conv, reference
vgtk/vgtk/so3conv/functional.py:2508-2549, vgtk/vgtk/spconv/functional.py:375-406, vgtk/vgtk/so3conv/modules.py:48-55)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    from equi_articulated_pose_b200 import lib
    lib.load()
    return torch.device("cuda:0")


def _case(dev, b, n, p, nn, ci, co, radius, sigma, seed, a=60):
    from oracle import so3 as O
    from equi_articulated_pose_b200 import ops, so3_constants as C
    import equi_articulated_pose_b200 as pkg
    pkg.install()
    import vgtk.so3conv.functional as L
    g = torch.Generator().manual_seed(seed)
    xyz = O.synthetic_cloud(b, n, seed).permute(0, 2, 1).contiguous().to(dev)
    sxyz = xyz[:, :, :p].contiguous()
    idx = ops.ball_query(sxyz, xyz, radius, nn)
    anchors = torch.from_numpy(C.get_anchors(a)).to(dev)
    kernels = torch.from_numpy(C.scaled_kernel_points(0.7 * radius, 1)).to(dev)
    rk = L.rotated_kernels(anchors, kernels)
    k = rk.shape[1]
    feats = torch.randn(b, n, a, ci, generator=g).to(dev)
    w_kc = (torch.randn(co, k * ci, generator=g) / (k * ci) ** 0.5).to(dev)
    gy = torch.randn(b * p * a, co, generator=g).to(dev)
    return xyz, sxyz, idx, rk, feats, w_kc, gy, sigma


SHAPES = [
    # b, n, p, nn, ci, co, radius, sigma
    (2, 128, 128, 16, 64, 64, 0.45, 0.08),      # layer 0.1 geometry (stride 1, nn = 16)
    (2, 128, 64, 32, 64, 128, 0.5, 0.1),        # strided, nn = 32
    (1, 96, 96, 12, 32, 72, 0.5, 0.1),          # ragged: nn not a multiple of 8, co = 72, ci = 32
    (2, 64, 32, 20, 128, 256, 0.7, 0.2),        # wide output
    (8, 512, 512, 16, 64, 64, 0.2828, 0.04),    # config 2 layer 0.1 at full size
    (2, 128, 128, 40, 64, 64, 0.6, 0.15),       # 33..48 neighbours: three k-steps (general warp-MMA kernels)
    (2, 256, 256, 64, 64, 128, 0.7, 0.2),       # model 38: 64 neighbours, four k-steps
]


@pytest.mark.parametrize("b,n,p,nn,ci,co,radius,sigma", SHAPES)
def test_inter_conv_matches_two_step_and_fp64(dev, b, n, p, nn, ci, co, radius, sigma):
    from equi_articulated_pose_b200 import ops
    xyz, sxyz, idx, rk, feats, w_kc, gy, sigma = _case(dev, b, n, p, nn, ci, co, radius, sigma, 11 * n + co)
    a, k = rk.shape[0], rk.shape[1]
    assert ops.inter_conv_supported(b, n, p, nn, a, k, ci, co)

    f1 = feats.clone().requires_grad_(True)
    w1 = w_kc.clone().requires_grad_(True)
    out1 = ops.InterConvFn.apply(f1, w1, xyz, sxyz, idx, rk, sigma)
    out1.backward(gy)

    f2 = feats.clone().requires_grad_(True)
    w2 = w_kc.clone().requires_grad_(True)
    g2 = ops.InterGroupFn.apply(f2, xyz, sxyz, idx, rk, sigma)
    out2 = ops.LinearFn.apply(g2.view(b * p * a, k * ci), w2, None)
    out2.backward(gy)

    if nn <= 32:
        assert torch.equal(out1, out2), "forward must be bit-identical to the two-step bf16x3 path"
        assert torch.equal(f1.grad, f2.grad) or float((f1.grad - f2.grad).abs().max()) <= 1e-6 * float(f2.grad.abs().max())
        assert float((w1.grad - w2.grad).abs().max()) <= 3e-6 * float(w2.grad.abs().max())
    else:       # the two-step path groups 33..64 neighbours with the exact FFMA kernels: same values to bf16x3 accuracy
        assert float((out1 - out2).abs().max()) <= 3e-5 * float(out2.abs().max())
        assert float((f1.grad - f2.grad).abs().max()) <= 5e-5 * float(f2.grad.abs().max())
        assert float((w1.grad - w2.grad).abs().max()) <= 5e-5 * float(w2.grad.abs().max())

    if b * p * a * k * ci <= 40_000_000:     # fp64 ground truth of the reference expressions (materialised weights)
        w = ops.inter_weights(xyz, sxyz, idx, rk, sigma).double()                      # [b,p,a,k,nn]
        fd = feats.double().requires_grad_(True)
        wd = w_kc.double().requires_grad_(True)
        gathered = torch.stack([fd[i][idx[i].long()] for i in range(b)])               # [b,p,nn,a,ci]
        G = torch.einsum('bpnac,bpakn->bpakc', gathered, w).reshape(b * p * a, k * ci)
        ref = G @ wd.t()
        ref.backward(gy.double())
        s = float(ref.abs().max())
        assert float((out1.double() - ref).abs().max()) <= 3e-5 * s
        assert float((f1.grad.double() - fd.grad).abs().max()) <= 5e-5 * float(fd.grad.abs().max())
        assert float((w1.grad.double() - wd.grad).abs().max()) <= 5e-5 * float(wd.grad.abs().max())


@pytest.mark.parametrize("b,n,p,nn,ci,co,radius,sigma", [SHAPES[0], SHAPES[1], SHAPES[3], SHAPES[6]])
def test_inter_conv_single_pass_bf16(dev, b, n, p, nn, ci, co, radius, sigma):
    """Contraction mode 4 (single-pass bf16, BASELINE config 3): only the hi plane of G exists; results within bf16 accuracy
    of the fp32-parity mode (stated tolerance: 1e-2 of the tensor maximum per conv, forward and both gradients)."""
    from equi_articulated_pose_b200 import ops
    xyz, sxyz, idx, rk, feats, w_kc, gy, sigma = _case(dev, b, n, p, nn, ci, co, radius, sigma, 7 * n + co)
    res = {}
    for mode in (3, 4):
        prev = ops.get_gemm_mode()
        ops.set_gemm_mode(mode)
        try:
            f = feats.clone().requires_grad_(True)
            w = w_kc.clone().requires_grad_(True)
            out = ops.InterConvFn.apply(f, w, xyz, sxyz, idx, rk, sigma)
            out.backward(gy)
            res[mode] = (out.detach(), f.grad, w.grad)
        finally:
            ops.set_gemm_mode(prev)
    for t3, t4 in zip(res[3], res[4]):
        e = float((t4 - t3).abs().max() / t3.abs().max())
        assert 1e-5 < e < 1e-2, e          # really a different arithmetic, and within the stated tolerance


def test_inter_conv_module_path_is_default(dev):
    """InterSO3Conv.forward takes the one-call path for the shapes it covers and its outputs equal the two-step path."""
    import equi_articulated_pose_b200 as pkg
    pkg.install()
    import vgtk.so3conv as sptk
    import vgtk.spconv as zptk
    from oracle import so3 as O
    from equi_articulated_pose_b200 import lib, ops
    torch.manual_seed(3)
    conv = sptk.InterSO3Conv(64, 64, 1, 1, 0.45, 0.08, 16).to(dev)
    xyz = O.synthetic_cloud(2, 128, 5).permute(0, 2, 1).contiguous().to(dev)
    feats = torch.randn(2, 128, 60, 64, device=dev).permute(0, 3, 1, 2)
    lib.PROFILE = []
    _, _, _, y = conv(zptk.SphericalPointCloud(xyz, feats, None))
    names = [r[0] for r in lib.PROFILE]
    lib.PROFILE = None
    assert "vgtkb_inter_conv_forward" in names and "vgtkb_inter_group_forward" not in names
    prev = ops.get_gemm_mode()
    try:
        ops.set_gemm_mode(1)                 # 3xTF32: the two-step path
        _, _, _, y1 = conv(zptk.SphericalPointCloud(xyz, feats, None))
    finally:
        ops.set_gemm_mode(prev)
    assert float((y.feats - y1.feats).abs().max()) <= 2e-5 * float(y1.feats.abs().max())


def test_operand_planes_block_matches_fp32_operands(dev):
    """A separable block with the norm kernels writing operand planes (default) against the same block with the planes
    switched off (every contraction converts its fp32 operand itself): same arithmetic, so outputs and gradients agree to
    the order of the atomics; and the planes are actually consumed (no silent fallback)."""
    from equi_articulated_pose_b200 import blocks, ops
    from oracle import so3 as O
    import vgtk.spconv as zptk
    params = {'dim_in': 64, 'dim_out': 64, 'kernel_size': 1, 'stride': 1, 'radius': 0.45, 'sigma': 0.08, 'n_neighbor': 16,
              'lazy_sample': True, 'dropout_rate': 0.0, 'multiplier': 2, 'activation': 'leaky_relu', 'pooling': None,
              'kanchor': 60, 'norm': 'BatchNorm2d'}
    torch.manual_seed(1)
    blk = blocks.SeparableSO3ConvBlock(params).to(dev).train()
    xyz = O.synthetic_cloud(2, 128, 9).permute(0, 2, 1).contiguous().to(dev)
    f0 = torch.randn(2, 128, 60, 64, device=dev)
    res = []
    for use in (True, False):
        ops._USE_PLANES = use
        ops.clear_planes()
        ops.PLANE_STATS.update(hit=0, miss=0)
        try:
            blk.zero_grad(set_to_none=True)
            f = f0.clone().requires_grad_(True)
            _, _, _, y = blk(zptk.SphericalPointCloud(xyz, f.permute(0, 3, 1, 2), None), None, None)
            y.feats.square().mean().backward()
            res.append((y.feats.detach().clone(), f.grad.clone(), {k: p.grad.clone() for k, p in blk.named_parameters()},
                        dict(ops.PLANE_STATS)))
        finally:
            ops._USE_PLANES = True
    (y1, g1, p1, st1), (y0, g0, p0, _) = res
    assert st1["hit"] >= 3 and st1["miss"] == 0, st1       # x planes (intra fwd), gy planes (intra bwd, inter bwd)
    assert float((y1 - y0).abs().max()) <= 1e-6 * float(y0.abs().max())
    assert float((g1 - g0).abs().max()) <= 2e-5 * float(g0.abs().max())
    for k in p0:
        assert float((p1[k] - p0[k]).abs().max()) <= 2e-5 * float(p0[k].abs().max()) + 1e-9, k


@pytest.mark.parametrize("mode,planes", [(3, False), (4, False), (3, True)])
def test_inter_conv_backward_stays_inside_its_workspace(dev, mode, planes):
    """vgtkb_inter_conv_backward with a workspace of EXACTLY the documented size (max(rows*co, 2*k*ci*co) floats) followed by a
    canary: few rows, so the weight split is the larger term.  Regression: without planes of grad_out the dG contraction
    split W^T with the lo plane at float offset k*ci*co of its scratch and overran the workspace by k*ci*co floats (silent
    inside torch's cached pool; an illegal address otherwise)."""
    from equi_articulated_pose_b200 import ops
    from equi_articulated_pose_b200.lib import call, ptr
    b, n, p, nn, ci, co = 1, 32, 16, 16, 64, 128
    xyz, sxyz, idx, rk, feats, w_kc, gy, sigma = _case(dev, b, n, p, nn, ci, co, 0.8, 0.2, 3)
    a, k = rk.shape[0], rk.shape[1]
    rows, kc = b * p * a, k * ci
    g_hi = torch.empty((rows, kc), dtype=torch.bfloat16, device=dev)
    g_lo = torch.empty((rows, kc), dtype=torch.bfloat16, device=dev)
    out = torch.empty((rows, co), dtype=torch.float32, device=dev)
    ws_f = torch.empty(co * kc, dtype=torch.float32, device=dev)
    call("vgtkb_inter_conv_forward", dev, b, n, p, nn, a, k, ci, co, ptr(xyz), ptr(sxyz), ptr(idx), ptr(rk), float(sigma),
         ptr(feats), ptr(w_kc), ptr(g_hi), ptr(g_lo if mode == 3 else None), ptr(ws_f), ptr(out), mode)
    size = max(rows * co, 2 * kc * co)
    assert size == 2 * kc * co                               # the case the regression needs
    pad = 2 * kc * co
    buf = torch.full((size + pad,), 12345.0, dtype=torch.float32, device=dev)
    gy_hi, gy_lo = ops.split_bf16(gy) if planes else (None, None)
    dg = torch.empty((rows, kc), dtype=torch.float32, device=dev)
    gx = torch.empty((b, n, a, ci), dtype=torch.float32, device=dev)
    gw = torch.empty((co, kc), dtype=torch.float32, device=dev)
    call("vgtkb_inter_conv_backward", dev, b, n, p, nn, a, k, ci, co, ptr(xyz), ptr(sxyz), ptr(idx), ptr(rk), float(sigma),
         ptr(w_kc), ptr(g_hi), ptr(g_lo if mode == 3 else None), ptr(gy), ptr(gy_hi), ptr(gy_lo), ptr(dg), ptr(gx), ptr(gw),
         ptr(buf[:size]), mode)
    torch.cuda.synchronize()
    assert bool((buf[size:] == 12345.0).all()), "vgtkb_inter_conv_backward wrote past its workspace"
    assert bool(torch.isfinite(gx).all()) and bool(torch.isfinite(gw).all())
