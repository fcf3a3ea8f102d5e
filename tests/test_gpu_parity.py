"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI of
libvgtkb200.so, against (1) the CPU oracle on the same seeded inputs, (2) the committed golden
fixtures produced by the reference's own Python, and (3) -- when oracle/_ref was built -- the
reference's own CUDA kernels recompiled for sm_100a.

Bars: integer / index outputs bit-exact; fp32 tensors max|d|/max|ref| <= 1e-4."""
import os

import numpy as np
import pytest
import torch

from tests.helpers import GOLD, rel_err, run_blocks_case, small_params, build_backbone

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from equi_articulated_pose_b200 import lib
    lib.load()                        # fails loudly when the .so is missing: no fallback
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def ops(dev):
    from equi_articulated_pose_b200 import ops as o
    return o


@pytest.fixture(scope="module")
def ref_mods(dev):
    """The reference's own kernels (oracle/_ref), or None when they were not built."""
    from oracle import build_ref
    return {n: build_ref.load_ref(n) for n in ("vgtk_ref_grouping", "vgtk_ref_gathering", "chamfer_ref")}


def _cloud(b, n, seed, kind="shell"):
    from oracle import so3 as O
    if kind == "shell":
        return O.synthetic_cloud(b, n, seed).permute(0, 2, 1).contiguous()
    g = torch.Generator().manual_seed(seed)
    if kind == "grid":       # many exact ties in distance
        pts = torch.randint(-4, 5, (b, 3, n), generator=g).float() * 0.125
        return pts.contiguous()
    return (torch.rand(b, 3, n, generator=g) - 0.5).contiguous()


# ------------------------------------------------------------------------------ ball query
@pytest.mark.parametrize("b,n,m,ns,r,kind", [
    (2, 128, 64, 16, 0.3, "shell"),
    (8, 1024, 512, 32, 0.2, "shell"),          # config 2, layer 0.0
    (8, 512, 512, 16, 0.2828, "shell"),        # layer 0.1
    (3, 1000, 333, 32, 0.25, "uniform"),       # n % 4 == 0 but m odd
    (2, 1001, 77, 24, 0.25, "uniform"),        # unaligned rows -> non-TMA staging
    (1, 9000, 100, 64, 0.08, "uniform"),       # two shared-memory chunks
    (2, 256, 256, 1, 0.2, "shell"),            # nsample = 1
    (2, 256, 256, 8, 1e-4, "shell"),           # only the point itself
    (2, 256, 64, 300, 5.0, "shell"),           # everything hits, nsample > n
    (2, 512, 512, 16, 0.25, "grid"),           # d2 == r2 ties (strict <)
])
def test_ball_query_bit_exact(dev, ops, ref_mods, b, n, m, ns, r, kind):
    from oracle import cops
    xyz = _cloud(b, n, 100 + n + m, kind)
    q = xyz[:, :, :m].contiguous() if m <= n else _cloud(b, m, 5, kind)
    want = cops.ball_query(q.numpy(), xyz.numpy(), r, ns)
    got = ops.ball_query(q.to(dev), xyz.to(dev), r, ns)
    assert got.dtype == torch.int32 and tuple(got.shape) == (b, m, ns)
    assert np.array_equal(got.cpu().numpy(), want)
    if ref_mods["vgtk_ref_grouping"] is not None:
        ref = ref_mods["vgtk_ref_grouping"].ball_query(q.to(dev), xyz.to(dev), r, ns)
        assert torch.equal(ref, got)


def test_ball_query_empty(dev, ops):
    xyz = _cloud(2, 64, 1).to(dev)
    assert ops.ball_query(xyz[:, :, :0].contiguous(), xyz, 0.2, 8).shape == (2, 0, 8)
    assert ops.ball_query(xyz[:0], xyz[:0], 0.2, 8).shape == (0, 64, 8)


# ------------------------------------------------------------------------------ FPS
@pytest.mark.parametrize("b,n,m,kind", [
    (2, 64, 32, "shell"), (3, 380, 190, "uniform"), (8, 1024, 512, "shell"), (2, 1000, 500, "uniform"),
    (2, 4096, 512, "shell"), (1, 5000, 300, "uniform"), (2, 1024, 1024, "grid"), (2, 777, 50, "grid"),
    (1, 16384, 64, "uniform"), (2, 1, 1, "uniform"), (2, 3, 3, "uniform"),
])
def test_fps_bit_exact(dev, ops, ref_mods, b, n, m, kind):
    from oracle import cops
    xyz = _cloud(b, n, 300 + n, kind)
    if kind == "uniform" and n > 100:
        xyz[:, :, 5:40] *= 1e-2          # inside the |p|^2 <= 1e-3 skip ball
    want = cops.furthest_point_sampling(xyz.numpy(), m)
    got = ops.furthest_point_sampling(xyz.to(dev), m)
    assert np.array_equal(got.cpu().numpy(), want)
    if ref_mods["vgtk_ref_grouping"] is not None and n >= 2:
        ref = ref_mods["vgtk_ref_grouping"].furthest_point_sampling(xyz.to(dev), m)
        assert torch.equal(ref, got)


def test_fps_all_skipped(dev, ops):
    z = torch.zeros(2, 3, 128, device=dev)
    assert ops.furthest_point_sampling(z, 9).cpu().tolist() == [[0] * 9] * 2


# ------------------------------------------------------------------------------ gather
def test_gather_forward_backward(dev, ops, ref_mods):
    from oracle import cops
    g = torch.Generator().manual_seed(5)
    pts = torch.randn(3, 7, 129, generator=g)
    idx = torch.randint(0, 129, (3, 200), generator=g, dtype=torch.int32)
    out = ops.gather_points_forward(pts.to(dev), idx.to(dev))
    assert np.array_equal(out.cpu().numpy(), cops.gather_points_forward(pts.numpy(), idx.numpy()))
    go = torch.randn(3, 7, 200, generator=g)
    back = ops.gather_points_backward(go.to(dev), idx.to(dev), 129)
    want = cops.gather_points_backward(go.numpy(), idx.numpy(), 129)
    assert np.allclose(back.cpu().numpy(), want, atol=1e-5)
    if ref_mods["vgtk_ref_gathering"] is not None:
        assert torch.equal(ref_mods["vgtk_ref_gathering"].gather_points_forward(pts.to(dev), idx.to(dev)), out)


# ------------------------------------------------------------------------------ chamfer
@pytest.mark.parametrize("b,n,m", [(4, 64, 128), (2, 1000, 3000), (3, 2048, 2048), (1, 5, 4100), (8, 512, 1024)])
def test_chamfer_forward_bit_exact(dev, ops, ref_mods, b, n, m):
    from oracle import cops
    g = torch.Generator().manual_seed(b * 1000 + n)
    x, y = torch.randn(b, n, 3, generator=g), torch.randn(b, m, 3, generator=g)
    y[:, m // 2] = y[:, 1]                  # duplicate target: the lower index must win
    d1, d2, i1, i2 = ops.chamfer_forward(x.to(dev), y.to(dev))
    w1, w2, j1, j2 = cops.chamfer_forward(x.numpy(), y.numpy())
    assert np.array_equal(i1.cpu().numpy(), j1) and np.array_equal(i2.cpu().numpy(), j2)
    assert np.array_equal(d1.cpu().numpy(), w1) and np.array_equal(d2.cpu().numpy(), w2)
    if ref_mods["chamfer_ref"] is not None:
        r = ref_mods["chamfer_ref"].forward(x.to(dev), y.to(dev))
        assert torch.equal(r[0], d1) and torch.equal(r[1], d2) and torch.equal(r[2], i1) and torch.equal(r[3], i2)


def test_chamfer_backward(dev, ops):
    from oracle import cops
    g = torch.Generator().manual_seed(9)
    x, y = torch.randn(4, 300, 3, generator=g), torch.randn(4, 500, 3, generator=g)
    g1, g2 = torch.randn(4, 300, generator=g), torch.randn(4, 500, generator=g)
    _, _, i1, i2 = ops.chamfer_forward(x.to(dev), y.to(dev))
    gx, gy = ops.chamfer_backward(x.to(dev), y.to(dev), i1, i2, g1.to(dev), g2.to(dev))
    wx, wy = cops.chamfer_backward(x.numpy(), y.numpy(), i1.cpu().numpy(), i2.cpu().numpy(), g1.numpy(), g2.numpy())
    assert np.allclose(gx.cpu().numpy(), wx, atol=1e-5) and np.allclose(gy.cpu().numpy(), wy, atol=1e-5)


def test_chamfer_module_gradcheck_shape(dev):
    """The reference's only test (extensions/chamfer_dist/test.py:22-28) uses x(4,64,3), y(4,128,3)."""
    import equi_articulated_pose_b200 as eap
    eap.install()
    from extensions.chamfer_dist import ChamferDistance
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4, 64, 3, generator=g).to(dev).requires_grad_(True)
    y = torch.randn(4, 128, 3, generator=g).to(dev).requires_grad_(True)
    loss = ChamferDistance()(x, y)
    loss.backward()
    xd, yd = x.detach().double().requires_grad_(True), y.detach().double().requires_grad_(True)
    D = ((xd[:, :, None] - yd[:, None]) ** 2).sum(-1)
    ref = D.min(2)[0].mean() + D.min(1)[0].mean()
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-5
    assert rel_err(x.grad, xd.grad) < 1e-5 and rel_err(y.grad, yd.grad) < 1e-5


# ------------------------------------------------------------------------------ grouping vs reference fixtures
def test_inter_weights_and_grouping_match_reference_fixture(dev, ops):
    import equi_articulated_pose_b200 as eap
    eap.install()
    import vgtk.so3conv.functional as L
    from equi_articulated_pose_b200 import so3_constants as C
    g = np.load(os.path.join(GOLD, "ref_weights_small.npz"))
    anchors = torch.from_numpy(C.anchors_all()).to(dev)
    kern = torch.from_numpy(g["kernels"]).to(dev)
    w = L.inter_so3conv_grouping_anchor(torch.from_numpy(g["grouped_xyz"]).to(dev), anchors, kern, float(g["sigma"]))
    assert rel_err(w, torch.from_numpy(g["inter_w"])) < 1e-5
    # explicit-weights literal API
    import vgtk.spconv as zp
    G = zp.inter_zpconv_grouping_naive(torch.from_numpy(g["idx"]).to(dev), w, torch.from_numpy(g["feats"]).to(dev))
    assert rel_err(G, torch.from_numpy(g["grouped"])) < 1e-5


@pytest.mark.parametrize("ci,nn,k", [(64, 16, 24), (128, 32, 24), (8, 12, 24), (1, 32, 24), (3, 7, 5)])
def test_inter_group_forward_backward_vs_oracle(dev, ops, ci, nn, k):
    from oracle import so3 as O, cops
    from equi_articulated_pose_b200 import so3_constants as C
    import equi_articulated_pose_b200 as eap
    eap.install()
    import vgtk.so3conv.functional as L
    b, n, p, a = 2, 96, 48, 60
    g = torch.Generator().manual_seed(ci * 7 + nn)
    xyz = _cloud(b, n, 40 + ci)
    sxyz = xyz[:, :, :p].contiguous()
    idx = torch.from_numpy(cops.ball_query(sxyz.numpy(), xyz.numpy(), 0.6, nn))
    anchors = torch.from_numpy(C.anchors_all())
    kern = torch.from_numpy(C.scaled_kernel_points(0.7 * 0.6))[:k].contiguous()
    feats = torch.randn(b, ci, n, a, generator=g, requires_grad=True)
    gx = torch.gather(xyz, 2, idx.view(b, 1, -1).expand(-1, 3, -1).long()).view(b, 3, p, nn) - sxyz.unsqueeze(3)
    w = O.anchor_weights(gx, anchors, kern, 0.18)
    G_ref = O.inter_group_feats(idx, w, feats)                       # [b,c,k,p,a]
    go = torch.randn(G_ref.shape, generator=g)
    (G_ref * go).sum().backward()

    rk = L.rotated_kernels(anchors.to(dev), kern.to(dev))
    f_cl = feats.detach().permute(0, 2, 3, 1).contiguous().to(dev).requires_grad_(True)
    G = ops.InterGroupFn.apply(f_cl, xyz.to(dev), sxyz.to(dev), idx.to(dev), rk, 0.18)
    G_log = G.view(b, p, a, k, ci).permute(0, 4, 3, 1, 2)
    assert rel_err(G_log, G_ref) < 2e-5
    (G_log * go.to(dev)).sum().backward()
    assert rel_err(f_cl.grad.permute(0, 3, 1, 2), feats.grad) < 2e-5


def test_intra_conv_matches_reference_fixture(dev, ops):
    import equi_articulated_pose_b200 as eap
    eap.install()
    import vgtk.so3conv as sptk
    import vgtk.spconv as zptk
    g = np.load(os.path.join(GOLD, "ref_intra_small.npz"))
    conv = sptk.IntraSO3Conv(16, 24).to(dev)
    with torch.no_grad():
        conv.basic_conv.W.copy_(torch.from_numpy(g["W"]))
    f = torch.from_numpy(g["feats"]).to(dev)
    out = conv(zptk.SphericalPointCloud(torch.zeros(1, 3, 32, device=dev), f, None)).feats
    assert rel_err(out, torch.from_numpy(g["out"])) < FP32_TOL


def test_config1_intra_so3conv_forward(dev, ops):
    """BASELINE config 1: single IntraSO3Conv(64, 64) forward, N=128 points, the 12-wide
    anchor neighbourhood of the 60 SO(3) anchors, batch 1, weights from torch.manual_seed(0); vs the oracle."""
    import equi_articulated_pose_b200 as eap
    eap.install()
    import vgtk.so3conv as sptk
    import vgtk.spconv as zptk
    from oracle import so3 as O
    from equi_articulated_pose_b200 import so3_constants as C
    torch.manual_seed(0)
    conv = sptk.IntraSO3Conv(64, 64)
    g = torch.Generator().manual_seed(1000)
    feats = torch.randn(1, 64, 128, 60, generator=g)
    xyz = torch.rand(1, 3, 128, generator=g) - 0.5
    want = O.basic_conv(conv.basic_conv.W.detach(), O.intra_group_feats(torch.from_numpy(C.intra_idx()), feats))
    conv = conv.to(dev)
    out = conv(zptk.SphericalPointCloud(xyz.to(dev), feats.to(dev), None))
    assert tuple(out.feats.shape) == (1, 64, 128, 60) and torch.equal(out.xyz.cpu(), xyz)
    assert rel_err(out.feats, want) < FP32_TOL


@pytest.mark.parametrize("mode", [1, 3])
@pytest.mark.parametrize("pts,c,co", [(200, 64, 64), (129, 128, 64), (64, 64, 256), (300, 256, 128)])
def test_intra_conv_gather_gemm_fwd_bwd(dev, ops, mode, pts, c, co):
    """The fused intra conv (3-D TMA gather-GEMMs: forward, data gradient through the inverse table, weight
    gradient) against the oracle's materialised gather + matmul in fp64."""
    from oracle import so3 as O
    from equi_articulated_pose_b200 import so3_constants as C
    g = torch.Generator().manual_seed(pts + c)
    ii = torch.from_numpy(C.intra_idx())
    f = torch.randn(1, c, pts, 60, generator=g, dtype=torch.float64, requires_grad=True)
    W = (torch.randn(co, c * 12, generator=g, dtype=torch.float64) / (c * 12) ** 0.5).requires_grad_(True)
    out_ref = O.basic_conv(W, O.intra_group_feats(ii, f))                     # [1,co,pts,60]
    go = torch.randn(out_ref.shape, generator=g, dtype=torch.float64)
    (out_ref * go).sum().backward()

    t = ii.to(torch.int32).to(dev).contiguous()
    inv = torch.empty_like(t)
    cols = torch.arange(12, device=dev).view(1, 12).expand(60, 12)
    inv[t.long(), cols] = torch.arange(60, device=dev, dtype=torch.int32).view(60, 1).expand(60, 12)
    x = f.detach().float().permute(0, 2, 3, 1).reshape(pts, 60, c).contiguous().to(dev).requires_grad_(True)
    w_kc = W.detach().float().view(co, c, 12).transpose(1, 2).reshape(co, 12 * c).contiguous().to(dev).requires_grad_(True)
    prev = ops.get_gemm_mode()
    ops.set_gemm_mode(mode)
    try:
        rows = ops.IntraConvFn.apply(x, w_kc, t, inv)                          # [pts*60, co]
        (rows * go.float().permute(0, 2, 3, 1).reshape(pts * 60, co).to(dev)).sum().backward()
    finally:
        ops.set_gemm_mode(prev)
    tol = 5e-6 if mode == 1 else 3e-5
    assert rel_err(rows.view(pts, 60, co), out_ref[0].permute(1, 2, 0)) < tol
    assert rel_err(x.grad, f.grad[0].permute(1, 2, 0)) < tol
    gw_ref = W.grad.view(co, c, 12).transpose(1, 2).reshape(co, 12 * c)
    assert rel_err(w_kc.grad, gw_ref) < tol


def test_intra_group_backward(dev, ops):
    from oracle import so3 as O
    from equi_articulated_pose_b200 import so3_constants as C
    g = torch.Generator().manual_seed(3)
    ii = torch.from_numpy(C.intra_idx())
    f = torch.randn(2, 20, 9, 60, generator=g, requires_grad=True)
    G = O.intra_group_feats(ii, f)
    go = torch.randn(G.shape, generator=g)
    (G * go).sum().backward()
    y = f.detach().permute(0, 2, 3, 1).reshape(18, 60, 20).contiguous().to(dev).requires_grad_(True)
    Gd = ops.IntraGroupFn.apply(y, ii.int().to(dev))
    Gl = Gd.view(2, 9, 60, 12, 20).permute(0, 4, 3, 1, 2)
    assert torch.equal(Gl.cpu(), G.detach())
    (Gl * go.to(dev)).sum().backward()
    assert rel_err(y.grad.view(2, 9, 60, 20).permute(0, 3, 1, 2), f.grad) < 1e-6


# ------------------------------------------------------------------------------ GEMM / norm
@pytest.mark.parametrize("mode", [0, 1, 3])
@pytest.mark.parametrize("M,N,K", [(128 * 60, 64, 24), (3000, 64, 1536), (7680, 256, 3072), (999, 24, 192), (257, 130, 33),
                                   (5000, 64, 1), (3000, 24, 3)])
def test_gemm_nt_tn(dev, ops, mode, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    bias = torch.randn(N, generator=g)
    want = (A.double() @ B.double().t() + bias.double()).float()
    # mode 1 = 3xTF32 on tcgen05 with chunked round-to-nearest accumulation (~1e-6); mode 0 = FFMA
    # mode 3 = bf16x3 (two bf16 per operand: ~5e-6)
    tol = {0: 2e-5, 1: 5e-6, 3: 3e-5}[mode]
    got = ops.gemm_nt(A.to(dev), B.to(dev), bias.to(dev), mode=mode)
    assert rel_err(got, want) < tol
    D = torch.randn(M, N, generator=g)
    want_t = (D.double().t() @ A.double()).float()
    got_t = ops.gemm_tn(D.to(dev), A.to(dev), mode=mode)
    assert rel_err(got_t, want_t) < tol


@pytest.mark.parametrize("groups,rows,c,affine", [(1, 8 * 64 * 60, 64, True), (8, 64 * 60, 256, False), (3, 1000, 24, False), (1, 777, 7, True)])
def test_norm_act(dev, ops, groups, rows, c, affine):
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(rows + c)
    x = (torch.randn(groups, rows, c, generator=g) * 2 + 0.5).requires_grad_(True)
    gamma = (1 + 0.1 * torch.randn(c, generator=g)).requires_grad_(True) if affine else None
    beta = (0.1 * torch.randn(c, generator=g)).requires_grad_(True) if affine else None
    res = torch.randn(groups, rows, c, generator=g)
    xd = x.double()
    mean, var = xd.mean(1, keepdim=True), xd.var(1, unbiased=False, keepdim=True)
    yh = (xd - mean) / torch.sqrt(var + 1e-5)
    if affine:
        yh = yh * gamma.double() + beta.double()
    ref = F.leaky_relu(yh, 0.01) + res.double()
    go = torch.randn(ref.shape, generator=g)
    (ref * go.double()).sum().backward()
    xg = x.detach().to(dev).requires_grad_(True)
    gg = gamma.detach().to(dev).requires_grad_(True) if affine else None
    bg = beta.detach().to(dev).requires_grad_(True) if affine else None
    rm, rv = (torch.zeros(c, device=dev), torch.ones(c, device=dev)) if groups == 1 else (None, None)
    y = ops.norm_act(xg, gg, bg, res.to(dev), rm, rv)
    assert rel_err(y, ref) < 1e-5
    (y * go.to(dev)).sum().backward()
    assert rel_err(xg.grad, x.grad) < 5e-5
    if affine:
        assert rel_err(gg.grad, gamma.grad) < 5e-5 and rel_err(bg.grad, beta.grad) < 5e-5
    if groups == 1:
        assert rel_err(rm, 0.1 * mean.flatten().float()) < 1e-5


# ------------------------------------------------------------------------------ blocks
def test_blocks_match_reference_fixture(dev):
    err = run_blocks_case(dev)
    assert err["xyz"] == 0.0
    assert err["out"] < FP32_TOL and err["loss"] < FP32_TOL, err
    assert err["grad"] < 5e-4, err
    assert err["running_stats"] < 1e-5, err


def _oracle_run(n, b, seed, dtype, loss_kind="square"):
    """Classic backbone fwd+bwd with the oracle in `dtype` (fp64 = the ground truth both fp32
    implementations are measured against; indices always come from the fp32 coordinates)."""
    from oracle import so3 as O
    from equi_articulated_pose_b200 import so3_constants as C
    params = O.backbone_params(input_num=n)
    sd = O.init_backbone_state(params, seed=seed)
    pts = O.synthetic_cloud(b, n, seed + 1)
    sdo = {k: v.clone().to(dtype).requires_grad_(True) for k, v in sd.items()}
    for bi, blk in enumerate(params):           # BatchNorm running stats for the oracle
        for li, layer in enumerate(blk):
            co = layer['args']['dim_out']
            for pre in (f'backbone.{bi}.blocks.{li}.inter_conv.norm.', f'backbone.{bi}.blocks.{li}.norm.'):
                sdo[pre + 'running_mean'], sdo[pre + 'running_var'] = torch.zeros(co, dtype=dtype), torch.ones(co, dtype=dtype)
    xyz = pts.permute(0, 2, 1).contiguous().to(dtype)
    oxyz, of = O.backbone_forward(sdo, params, xyz, torch.ones(b, 1, n, 60, dtype=dtype),
                                  torch.from_numpy(C.anchors_all()).to(dtype), torch.from_numpy(C.intra_idx()),
                                  C.kernel_points_base(), training=True)
    loss = _loss(of, loss_kind)
    loss.backward()
    return params, sd, pts, sdo, of.detach(), oxyz, loss.detach()


def _loss(feats, kind):
    """'square' = the benchmark loss feats.square().mean(): nearly invariant under the last
    normalisation layers, so its gradients are a small residual of large cancelling terms (fp32 noise
    ~1e-2).  'proj' = projection on a fixed random tensor: well-conditioned gradients."""
    if kind == "square":
        return feats.square().mean()
    w = torch.randn(feats.shape, generator=torch.Generator().manual_seed(77)).to(feats.dtype).to(feats.device)
    return (feats * w).mean()


def _oracle_case(dev, n, b, seed, loss_kind="square"):
    params, sd, pts, sdo, of, oxyz, loss = _oracle_run(n, b, seed, torch.float32, loss_kind)
    net = build_backbone(params, sd, dev)
    net.train()
    out = net(pts.to(dev))
    l2 = _loss(out.feats, loss_kind)
    l2.backward()
    return sdo, of, oxyz, loss, net, out, l2


# Gradient bar of the whole-backbone tests: every tensor at most GRAD_RATIO x as far from the fp64 gradient as the fp32
# oracle (= the reference's arithmetic) is, or within GRAD_FLOOR of the tensor maximum.  No blanket 10 % floor: the
# per-tensor tables these tests write (gpurun_out/r2_grad_parity_*.json, committed under profiles/) carry the evidence.
# Measured at N=256, B=2 (profiles/r2_grad_parity_n256_*.json): the worst tensor sits at 4.8e-2 of its maximum in EVERY
# arithmetic (FFMA, 3xTF32, bf16x3: the last block's skip conv sees 2 x 16 points), the fp32 oracle itself at 4.0e-2; the
# FFMA fallback (mode 0, fp32 atomics over 256-row slices, not a default path) reaches 7.8e-2 on one tensor.  At the
# benchmarked size (8 x 1024 points) the floor is 1e-2: tests/test_gpu_bench_parity.py.
GRAD_RATIO, GRAD_FLOOR, GRAD_FLOOR_FFMA = 5.0, 5e-2, 1e-1


def _dump_grad_table(fname, rows):
    import json
    from tests.helpers import ROOT
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump([{"tensor": n, "scale": s, "e_gpu_over_scale": (eg / s if s > 0 else None), "e_ref_over_scale": (er / s if s > 0 else None)}
               for n, s, eg, er in rows], open(os.path.join(ROOT, "gpurun_out", fname), "w"), indent=1)


def _grad_errors(net, sdo, sd64):
    rows = []
    for name, p in net.named_parameters():
        truth = sd64[name].grad
        scale = float(truth.abs().max())
        e_ref = float((sdo[name].grad.double() - truth).abs().max())
        e_gpu = float((p.grad.double().cpu() - truth).abs().max())
        rows.append((name, scale, e_gpu, e_ref))
    return rows


@pytest.mark.parametrize("gemm_mode", [0, 1, 3])
@pytest.mark.parametrize("loss_kind", ["proj", "square"])
def test_classic_backbone_fwd_bwd_vs_oracle(dev, ops, loss_kind, gemm_mode):
    """Full 7-layer classic backbone (cls_so3net_pn defaults), N=256, B=2, fwd+bwd vs the oracle.

    Forward: 1e-4 against the fp32 oracle (= the reference's arithmetic) and the fp64 oracle.
    Backward: every gradient is measured against the fp64 evaluation of the same graph, normalised by
    the largest entry of its tensor.  fp32 gradients of this network have a noise floor of
    3e-3..4e-2 whatever the implementation (the fp32 CPU oracle = the reference's arithmetic is that
    far from fp64; so are the FFMA mode 0 and the 3xTF32 mode 1 with chunked RN accumulation), so the
    bar is "as close to fp64 as the fp32 reference": within 5x of its error or 3e-2 of the tensor maximum
    (GRAD_RATIO / GRAD_FLOOR above; the benchmark-size version of this test is tests/test_gpu_bench_parity.py)."""
    prev = ops.get_gemm_mode()
    ops.set_gemm_mode(gemm_mode)
    try:
        sdo, of, oxyz, loss, net, out, l2 = _oracle_case(dev, 256, 2, 11, loss_kind)
    finally:
        ops.set_gemm_mode(prev)
    assert torch.equal(out.xyz.cpu(), oxyz.float())
    assert rel_err(out.feats, of) < FP32_TOL
    assert abs(float(l2) - float(loss)) < 2e-4 * max(abs(float(loss)), 1e-3)
    _, _, _, sd64, of64, _, _ = _oracle_run(256, 2, 11, torch.float64, loss_kind)
    assert rel_err(out.feats, of64) < FP32_TOL
    rows = _grad_errors(net, sdo, sd64)
    gmax = max(scale for _, scale, _, _ in rows)
    _dump_grad_table(f"r2_grad_parity_n256_{loss_kind}_mode{gemm_mode}.json", rows)
    bad = []
    for name, scale, e_gpu, e_ref in rows:
        if scale < 1e-6 * gmax:   # structurally zero gradients (bias before BatchNorm, first skip branch)
            assert e_gpu < 1e-4 * gmax, (name, e_gpu)
        elif e_gpu > max(GRAD_RATIO * e_ref, (GRAD_FLOOR_FFMA if gemm_mode == 0 else GRAD_FLOOR) * scale):
            bad.append((name, e_gpu / scale, e_ref / scale))
    assert not bad, bad


def test_config2_shape_forward_vs_oracle(dev):
    """BASELINE config 2 shape (N=1024, A=60) for one cloud: forward within 1e-4 of the oracle."""
    params, sd, pts, sdo, of, oxyz, loss = _oracle_run(1024, 1, 5, torch.float32)
    net = build_backbone(params, sd, dev).train()
    with torch.no_grad():
        out = net(pts.to(dev))
    assert torch.equal(out.xyz.cpu(), oxyz)
    assert rel_err(out.feats, of) < FP32_TOL


def test_config2_full_size_equivariance(dev):
    """BASELINE config 2 (N=1024, A=60, B=8) at full size through a size-independent property:
    rotating the input by anchor rotation R_g permutes the anchor axis of the output features
    (SURVEY appendix C.3): f'[.., a] = f[.., pi_g(a)] with Rs[pi_g(a)] = R_g^T R_a."""
    from oracle import so3 as O
    from equi_articulated_pose_b200 import so3_constants as C
    params = O.backbone_params(input_num=1024)
    sd = O.init_backbone_state(params, seed=3)
    net = build_backbone(params, sd, dev)
    net.train()
    pts = O.synthetic_cloud(8, 1024, 2000).to(dev)
    Rs = torch.from_numpy(C.anchors_all()).to(dev)
    with torch.no_grad():
        f0 = net(pts).feats
        assert tuple(f0.shape) == (8, 256, 64, 60)
        for gidx in (7, 41):
            Rg = Rs[gidx]
            f1 = net(pts @ Rg.t()).feats                      # x' = R_g x
            target = torch.einsum('ji,ajk->aik', Rg, Rs)       # R_g^T R_a
            perm = (target.reshape(60, 1, 9) - Rs.reshape(1, 60, 9)).square().sum(-1).argmin(1)
            err = rel_err(f1, f0[..., perm])
            # ball-query membership can flip for pairs exactly at the radius; allow a loose bound
            assert err < 5e-3, (gidx, err)


def test_config4_dense_scan_with_chamfer(dev, ops):
    """BASELINE config 4: classic backbone at N=4096 (first stride 8: 4096 -> 512), batch 32, loss
    feats.square().mean() + ChamferDistance(Y, X) with Y = X[:, :1024] + noise (requires grad).  Full size,
    so the checks are size-independent: the first-layer FPS / ball-query indices of two clouds are bit-exact
    against the oracle, the output has the reference shape walk, everything is finite, the chamfer gradient
    equals 2 (y - nn(y)) / (B n) + the reverse-direction term on a sampled row, and a second run is bit-identical
    in the forward (no atomics on the forward path)."""
    import equi_articulated_pose_b200 as eap
    eap.install()
    from extensions.chamfer_dist import ChamferDistance
    from oracle import so3 as O, cops
    B, N = 32, 4096
    params = O.backbone_params(input_num=N)
    assert params[0][0]['args']['stride'] == 8 and params[0][0]['args']['n_neighbor'] == 32
    net = build_backbone(params, O.init_backbone_state(params, seed=4), dev).train()
    pts = O.synthetic_cloud(B, N, 4000)
    X = pts.to(dev)
    g = torch.Generator().manual_seed(4)
    Y = (pts[:, :1024] + 0.01 * torch.randn(B, 1024, 3, generator=g)).to(dev).requires_grad_(True)
    out = net(X)
    assert tuple(out.feats.shape) == (B, 256, 64, 60) and tuple(out.xyz.shape) == (B, 3, 64)
    loss = out.feats.square().mean() + ChamferDistance()(Y, X)
    loss.backward()
    assert torch.isfinite(loss) and torch.isfinite(Y.grad).all()
    assert all(torch.isfinite(p.grad).all() for p in net.parameters())
    # index ops of the first layer, two clouds, against the oracle
    xyz2 = pts[:2].permute(0, 2, 1).contiguous()
    fps_ref = cops.furthest_point_sampling(xyz2.numpy(), 512)
    fps = ops.furthest_point_sampling(xyz2.to(dev), 512)
    assert np.array_equal(fps.cpu().numpy(), fps_ref)
    q = cops.gather_points_forward(xyz2.numpy(), fps_ref)
    a0 = params[0][0]['args']
    assert np.array_equal(ops.ball_query(torch.from_numpy(q).to(dev), xyz2.to(dev), a0['radius'], 32).cpu().numpy(),
                          cops.ball_query(q, xyz2.numpy(), a0['radius'], 32))
    # chamfer gradient on one cloud against the closed form
    with torch.no_grad():
        d = ((Y[0, :, None] - X[0, None]) ** 2).sum(-1)
        i1 = d.argmin(1)
        i2 = d.argmin(0)
        gref = 2 * (Y[0] - X[0][i1]) / (B * 1024)
        gref.index_add_(0, i2, -2 * (X[0] - Y[0][i2]) / (B * N))
    assert rel_err(Y.grad[0], gref) < 1e-4
    with torch.no_grad():
        assert torch.equal(net(X).feats, net(X).feats)


@pytest.mark.parametrize("pm", [0, 1])
def test_pose_grouping_non_identity_matches_reference_fixture(dev, ops, pm):
    """InterSO3PoseConv grouping with random per-point rotations (stride 1) against the reference's own output
    (tests/golden/ref_pose_group_small.npz), forward; backward against the oracle's autograd."""
    from oracle import so3 as O, cops
    from equi_articulated_pose_b200 import so3_constants as C
    import equi_articulated_pose_b200 as eap
    eap.install()
    import vgtk.so3conv.functional as L
    g = np.load(os.path.join(GOLD, "ref_pose_group_small.npz"))
    xyz, pose = torch.from_numpy(g["xyz"]), torch.from_numpy(g["pose"])
    feats = torch.from_numpy(g["feats"]).requires_grad_(True)
    anchors, kern = torch.from_numpy(C.anchors_all()), torch.from_numpy(g["kernels"])
    idx = torch.from_numpy(cops.ball_query(xyz.numpy(), xyz.numpy(), float(g["radius"]), int(g["nn"])))
    G_ref, _, pi_ref = O.pose_inter_group_feats(xyz, pose, feats, idx, anchors, kern, float(g["sigma"]), pm)
    go = torch.randn(G_ref.shape, generator=torch.Generator().manual_seed(3))
    (G_ref * go).sum().backward()

    rel, perm = ops.pose_neighbourhood(xyz.to(dev), pose.to(dev), idx.to(dev), anchors.to(dev), with_perm=pm != 0)
    if pm:
        assert torch.equal(perm.cpu().long(), pi_ref)
    rk = L.rotated_kernels(anchors.to(dev), kern.to(dev))
    x = feats.detach().permute(0, 2, 3, 1).contiguous().to(dev).requires_grad_(True)
    G = ops.PoseGroupFn.apply(x, idx.to(dev), rel, perm, rk, float(g["sigma"]))
    b, n, a, kc = G.shape
    G_log = G.view(b, n, a, 24, kc // 24).permute(0, 4, 3, 1, 2)
    assert rel_err(G_log, torch.from_numpy(g[f"grouped_pm{pm}"])) < 1e-5
    (G_log * go.to(dev)).sum().backward()
    assert rel_err(x.grad.permute(0, 3, 1, 2), feats.grad) < 1e-5


def test_pose_conv_module_non_identity(dev):
    """Module level: InterSO3PoseConv with rotated poses runs the general kernels and returns a pose-carrying cloud;
    with the identity pose it equals InterSO3Conv bit for bit."""
    import equi_articulated_pose_b200 as eap
    eap.install()
    import vgtk.so3conv as sptk
    import vgtk.spconv as zptk
    g = np.load(os.path.join(GOLD, "ref_pose_group_small.npz"))
    torch.manual_seed(0)
    conv = sptk.InterSO3PoseConv(5, 8, 1, 1, float(g["radius"]), float(g["sigma"]), int(g["nn"]), kanchor=60, permute_modes=1).to(dev)
    plain = sptk.InterSO3Conv(5, 8, 1, 1, float(g["radius"]), float(g["sigma"]), int(g["nn"]), kanchor=60).to(dev)
    plain.load_state_dict(conv.state_dict())
    xyz, feats = torch.from_numpy(g["xyz"]).to(dev), torch.from_numpy(g["feats"]).to(dev)
    eye = torch.eye(4, device=dev).repeat(1, xyz.shape[2], 1, 1)
    _, _, _, o_id = conv(zptk.SphericalPointCloudPose(xyz, feats, None, eye))
    _, _, _, o_pl = plain(zptk.SphericalPointCloud(xyz, feats, None))
    assert torch.equal(o_id.feats, o_pl.feats)
    idx_rot, _, _, o_rot = conv(zptk.SphericalPointCloudPose(xyz, feats, None, torch.from_numpy(g["pose"]).to(dev)))
    assert torch.equal(o_rot.pose, torch.from_numpy(g["pose"]).to(dev))
    # non-identity pose: the module output equals the oracle's pose grouping (pinned on the reference fixture by the test
    # above: rotated offsets R_p R_j^T (x_j - x_p), nearest-anchor permutation) followed by BasicSO3Conv with the same W
    from oracle import so3 as O, cops
    from equi_articulated_pose_b200 import so3_constants as C
    xyz_c, pose_c, feats_c = torch.from_numpy(g["xyz"]), torch.from_numpy(g["pose"]), torch.from_numpy(g["feats"])
    idx = torch.from_numpy(cops.ball_query(xyz_c.numpy(), xyz_c.numpy(), float(g["radius"]), int(g["nn"])))
    assert torch.equal(idx_rot.cpu(), idx)
    G_ref, _, _ = O.pose_inter_group_feats(xyz_c, pose_c, feats_c, idx, torch.from_numpy(C.anchors_all()), conv.kernels.cpu(),
                                           float(g["sigma"]), 1)
    want = O.basic_conv(conv.basic_conv.W.detach().cpu().double(), G_ref.double())
    assert tuple(o_rot.feats.shape) == tuple(want.shape)
    assert rel_err(o_rot.feats, want.float()) < 1e-4
    assert not torch.allclose(o_rot.feats, o_pl.feats)


@pytest.mark.parametrize("lazy,pm", [(0, 0), (0, 1), (1, 1)])
def test_pose_grouping_strided_matches_reference_fixture(dev, ops, lazy, pm):
    """Strided branch of the pose grouping (functional.py:931-1029) against the reference's own output
    (tests/golden/ref_pose_group_strided_small.npz): kernels forward, permutation table bit-exact vs the oracle, backward
    against the oracle's autograd; and the module (sampling + ball query + kernels + BasicSO3Conv) against oracle + fixture."""
    from oracle import so3 as O
    from equi_articulated_pose_b200 import so3_constants as C
    import equi_articulated_pose_b200 as eap
    eap.install()
    import vgtk.so3conv as sptk
    import vgtk.spconv as zptk
    import vgtk.so3conv.functional as L
    g = np.load(os.path.join(GOLD, "ref_pose_group_strided_small.npz"))
    xyz, pose = torch.from_numpy(g["xyz"]), torch.from_numpy(g["pose"])
    feats = torch.from_numpy(g["feats"]).requires_grad_(True)
    anchors, kern = torch.from_numpy(C.anchors_all()), torch.from_numpy(g["kernels"])
    stride, radius, sigma, nn_ = int(g["stride"]), float(g["radius"]), float(g["sigma"]), int(g["nn"])
    _, idx, sidx, sxyz = O.ball_grouping(xyz, stride, radius, nn_, bool(lazy))
    G_ref, _, pi_ref = O.pose_inter_group_feats(xyz, pose, feats, idx, anchors, kern, sigma, pm, sxyz, sidx)
    go = torch.randn(G_ref.shape, generator=torch.Generator().manual_seed(4))
    (G_ref * go).sum().backward()

    rel, perm = ops.pose_neighbourhood(xyz.to(dev), pose.to(dev), idx.to(dev).int(), anchors.to(dev), with_perm=pm != 0,
                                       sample_xyz=sxyz.contiguous().to(dev), sample_idx=sidx.to(dev).int())
    if pm:
        assert torch.equal(perm.cpu().long(), pi_ref)
    rk = L.rotated_kernels(anchors.to(dev), kern.to(dev))
    x = feats.detach().permute(0, 2, 3, 1).contiguous().to(dev).requires_grad_(True)
    G = ops.PoseGroupFn.apply(x, idx.to(dev).int(), rel, perm, rk, sigma)
    b, p, a, kc = G.shape
    assert p == xyz.shape[2] // stride
    G_log = G.view(b, p, a, 24, kc // 24).permute(0, 4, 3, 1, 2)
    assert rel_err(G_log, torch.from_numpy(g[f"grouped_lazy{lazy}_pm{pm}"])) < 1e-5
    (G_log * go.to(dev)).sum().backward()
    assert rel_err(x.grad.permute(0, 3, 1, 2), feats.grad) < 1e-5

    # module level: same sampling, same neighbourhoods, pose of the centres, BasicSO3Conv on top
    torch.manual_seed(0)
    conv = sptk.InterSO3PoseConv(5, 8, 1, stride, radius, sigma, nn_, lazy_sample=bool(lazy), kanchor=60, permute_modes=pm).to(dev)
    r_idx, r_w, s_idx, out = conv(zptk.SphericalPointCloudPose(xyz.to(dev), feats.detach().to(dev), None, pose.to(dev)))
    assert r_idx is None and torch.equal(s_idx.cpu().long(), torch.from_numpy(g[f"sample_idx_lazy{lazy}"]).long())
    assert torch.equal(out.xyz.cpu(), torch.from_numpy(g[f"new_xyz_lazy{lazy}"]))
    assert torch.equal(out.pose.cpu(), torch.from_numpy(g[f"sampled_pose_lazy{lazy}"]))
    want = O.basic_conv(conv.basic_conv.W.detach().cpu().double(), torch.from_numpy(g[f"grouped_lazy{lazy}_pm{pm}"]).double())
    assert rel_err(out.feats, want.float()) < 1e-4


# ------------------------------------------------------------------------------ anchor-orbit chamfer (model 38 loss)
def _orbit_case(b, a, m, n, seed):
    from oracle import so3 as O
    from equi_articulated_pose_b200 import so3_constants as C
    g = torch.Generator().manual_seed(seed)
    anchors = torch.from_numpy(np.ascontiguousarray(C.get_anchors(60)))[:a].float()
    # predicted rotation (random, orthonormalised) composed with the anchors, as the model builds glb_R
    q, _ = torch.linalg.qr(torch.randn(b, 3, 3, generator=g))
    rot = torch.matmul(q.unsqueeze(1), anchors.unsqueeze(0)).contiguous()
    trans = (torch.randn(b, a, 3, generator=g) * 0.05).contiguous()
    ori = O.synthetic_cloud(b, n, seed).permute(0, 2, 1).contiguous() * 0.5             # [B,3,N]
    canon = (torch.randn(b, 3, m, generator=g) * 0.3).contiguous()
    return canon, rot, trans, ori


@pytest.mark.parametrize("b,a,m,n", [(2, 60, 256, 512), (1, 7, 100, 3000), (3, 20, 513, 64)])
def test_anchor_chamfer_matches_unfused_reference_path(dev, ops, b, a, m, n):
    from oracle import so3 as O
    canon, rot, trans, ori = _orbit_case(b, a, m, n, 77 + m)
    want = O.anchor_orbit_chamfer(canon, rot, trans, ori)
    d1, d2, i1, i2 = ops.anchor_chamfer(canon.transpose(1, 2).contiguous().to(dev), rot.to(dev), trans.to(dev),
                                        ori.transpose(1, 2).contiguous().to(dev))
    # the transformed points differ from torch.matmul's by rounding only; the matching is the pinned chamfer kernel's
    for got, ref, gi, ri in ((d1, want["d1"], i1, want["i1"]), (d2, want["d2"], i2, want["i2"])):
        got, gi = got.cpu(), gi.cpu()
        assert torch.allclose(got, ref, rtol=1e-4, atol=1e-7)
        assert float((gi != ri).float().mean()) < 1e-3                       # near-ties only
    # unfused composition on the GPU through the bit-exact chamfer kernel: same transformed points -> identical
    y = want["transformed"].contiguous().view(b * a, m, 3).to(dev)
    e = ori.transpose(1, 2).unsqueeze(1).repeat(1, a, 1, 1).contiguous().view(b * a, n, 3).to(dev)
    u1, u2, _, _ = ops.chamfer_forward(y, e)
    assert rel_err(d1.view(b * a, m), u1) < 1e-5 and rel_err(d2.view(b * a, n), u2) < 1e-5
    # module: per-anchor means, orbit minimum and its index (SPConvNets/models/...38...py:437-450)
    import equi_articulated_pose_b200 as eap
    eap.install()
    from extensions.chamfer_dist import AnchorChamferDistance
    minn, orbit, r2o, o2r = AnchorChamferDistance()(canon.to(dev), rot.to(dev), trans.to(dev), ori.to(dev))
    assert rel_err(r2o, want["cd_r2o"]) < 1e-5 and rel_err(o2r, want["cd_o2r"]) < 1e-5
    assert torch.equal(orbit.cpu(), want["orbit"]) and rel_err(minn, want["minn"]) < 1e-5


def test_anchor_chamfer_gradients(dev, ops):
    canon, rot, trans, ori = _orbit_case(2, 12, 96, 160, 5)
    leaves = [t.to(dev).requires_grad_(True) for t in (canon.transpose(1, 2).contiguous(), rot, trans,
                                                       ori.transpose(1, 2).contiguous())]
    d1, d2, _, _ = ops.anchor_chamfer(*leaves)
    g = torch.Generator().manual_seed(3)
    w1, w2 = torch.rand(d1.shape, generator=g).to(dev), torch.rand(d2.shape, generator=g).to(dev)
    ((d1 * w1).sum() + (d2 * w2).sum()).backward()
    dl = [t.detach().double().requires_grad_(True) for t in leaves]
    y = torch.einsum('bajk,bmk->bamj', dl[1], dl[0]) + dl[2].unsqueeze(2)
    D = ((y[:, :, :, None] - dl[3][:, None, None]) ** 2).sum(-1)             # [B,A,M,N]
    ((D.min(3)[0] * w1.double()).sum() + (D.min(2)[0] * w2.double()).sum()).backward()
    for got, ref in zip(leaves, dl):
        assert rel_err(got.grad, ref.grad) < 2e-5


# ------------------------------------------------------------------------------ plain FPS / torch_cluster shim
@pytest.mark.parametrize("b,n,m,kind", [(4, 512, 128, "shell"), (2, 2048, 512, "uniform"), (3, 1000, 1000, "grid"),
                                        (1, 4096, 1024, "shell")])
def test_fps_plain_bit_exact_vs_oracle(dev, ops, b, n, m, kind):
    from oracle import cops
    xyz = _cloud(b, n, 31 + n, kind)
    xyz[:, :, 3] = 0.0                                   # a point at the origin is eligible here (it is not in the vgtk kernel)
    got = ops.fps_plain(xyz.to(dev), m).cpu().numpy()
    assert np.array_equal(got, cops.fps_plain(xyz.numpy(), m))


def test_torch_cluster_shim_matches_reference_wrapper_semantics(dev):
    """farthest_point_sampling of SPConvNets/models/model_util.py:183-200: flat indices, segment after segment."""
    import equi_articulated_pose_b200 as eap
    eap.install()
    import torch_cluster
    from oracle import cops
    bz, n, ns = 3, 600, 150
    pos = _cloud(bz, n, 5, "uniform").permute(0, 2, 1).contiguous()                      # [bz, N, 3]
    batch = torch.arange(bz).view(bz, 1).repeat(1, n).view(-1)
    idx = torch_cluster.fps(pos.view(-1, 3).to(dev), batch.to(dev), ratio=float(ns / n), random_start=False)
    want = cops.fps_plain(pos.permute(0, 2, 1).contiguous().numpy(), ns).astype(np.int64) + np.arange(bz)[:, None] * n
    assert idx.dtype == torch.int64 and np.array_equal(idx.cpu().numpy(), want.reshape(-1))


# ------------------------------------------------------------------------------ PointnetSO3Conv head
@pytest.mark.parametrize("na", [60, 1])
def test_pointnet_head_matches_reference_fixture(dev, na):
    """vgtk.so3conv.PointnetSO3Conv (drop-in, fused kernels) against the reference's own module, fwd + bwd."""
    import equi_articulated_pose_b200 as eap
    eap.install()
    import vgtk.so3conv as sptk
    g = np.load(os.path.join(GOLD, "ref_pointnet_small.npz"))
    t = {k[len(f'a{na}_'):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(f'a{na}_')}
    head = sptk.PointnetSO3Conv(t['feats'].shape[1], t['pooled'].shape[1], kanchor=na).to(dev)
    head.load_state_dict({'embed.weight': t['weight'], 'embed.bias': t['bias'], 'anchors': head.anchors.cpu()})
    feats = t['feats'].to(dev).requires_grad_(True)
    pooled = head(sptk.SphericalPointCloud(t['xyz'].to(dev), feats, None))
    assert pooled.shape == t['pooled'].shape and rel_err(pooled, t['pooled']) < FP32_TOL
    (pooled * t['grad_out'].to(dev)).sum().backward()
    assert rel_err(feats.grad, t['grad_feats']) < FP32_TOL
    assert rel_err(head.embed.weight.grad, t['grad_weight']) < FP32_TOL
    assert rel_err(head.embed.bias.grad, t['grad_bias']) < FP32_TOL
    head.return_raw = True
    with torch.no_grad():
        raw = head(sptk.SphericalPointCloud(t['xyz'].to(dev), feats.detach(), None))
    assert raw.shape == t['raw'].shape and rel_err(raw, t['raw']) < FP32_TOL
    raw_g = head(sptk.SphericalPointCloud(t['xyz'].to(dev), feats, None))      # differentiable raw path
    assert rel_err(raw_g, t['raw']) < FP32_TOL


def test_pointnet_head_backbone_size_vs_oracle(dev):
    """Config-2 tail size: feats [8, 256, 64, 60] -> [8, 128, 60]; first-maximum arg-max, xyz gradient."""
    import equi_articulated_pose_b200 as eap
    eap.install()
    import vgtk.so3conv as sptk
    from oracle import so3 as O
    g = torch.Generator().manual_seed(11)
    nb, c, npt, na, co = 8, 256, 64, 60, 128
    head = sptk.PointnetSO3Conv(c, co, kanchor=na).to(dev)
    feats = torch.randn(nb, c, npt, na, generator=g)
    xyz = (torch.rand(nb, 3, npt, generator=g) - 0.5)
    fd, xd = feats.to(dev).requires_grad_(True), xyz.to(dev).requires_grad_(True)
    out = head(sptk.SphericalPointCloud(xd, fd, None))
    gout = torch.randn(out.shape, generator=g)
    (out * gout.to(dev)).sum().backward()
    w64, b64 = head.embed.weight.detach().cpu().double().view(co, c + 3), head.embed.bias.detach().cpu().double()
    f64, x64 = feats.double().requires_grad_(True), xyz.double().requires_grad_(True)
    ref = O.pointnet_so3conv(w64, b64, head.anchors.cpu().double(), x64, f64)
    (ref * gout.double()).sum().backward()
    assert rel_err(out, ref) < FP32_TOL

    def flipped(got, want):
        """fraction of elements off by more than the tolerance: only the outputs whose two largest candidates are
        closer than the fp32 rounding of the embedding may pick the other point (a handful of 61440 maxima)"""
        return float(((got.detach().cpu().double() - want).abs() > FP32_TOL * float(want.abs().max())).double().mean())
    assert flipped(fd.grad, f64.grad) < 1e-3 and flipped(xd.grad, x64.grad) < 2e-2


# ------------------------------------------------------------------------------ grouping: tensor-core variants
@pytest.mark.parametrize("b,n,p,nn,ci,k", [
    (2, 128, 128, 16, 64, 24),      # stride-1 layer shape, one k-step
    (2, 200, 100, 32, 128, 24),     # stride-2 layer shape, two k-steps
    (1, 90, 90, 7, 32, 24),         # fewer neighbours than one k-step: padded slots carry weight 0
    (2, 150, 75, 17, 96, 24),       # nn just above one k-step, Ci = 3 chunks
    (1, 64, 64, 12, 160, 13),       # fewer kernel points than 24: tensor-map box of 13 rows
    (3, 256, 128, 32, 256, 24),
])
def test_inter_group_mma_variants_match_fp32_kernel(dev, b, n, p, nn, ci, k):
    """mode 3 (warp-level bf16x3 MMAs, tensor-map store / load) against mode 0 (exact fp32 FFMA kernel), forward and
    backward, through the C ABI; and against the oracle's torch restatement on the smallest case."""
    from equi_articulated_pose_b200 import ops, so3_constants as C
    from equi_articulated_pose_b200.lib import call, ptr
    from oracle import so3 as O
    g = torch.Generator().manual_seed(100 + nn + ci)
    anchors = torch.from_numpy(np.ascontiguousarray(C.get_anchors(60))).float()
    xyz = O.synthetic_cloud(b, n, 5 + n).permute(0, 2, 1).contiguous().to(dev)
    sxyz = xyz[:, :, :p].contiguous()
    radius, sigma = 0.45, 0.1
    idx = ops.ball_query(sxyz, xyz, radius, nn)
    base = torch.randn(k, 3, generator=g)
    base = base / base.norm(dim=1, keepdim=True) * torch.rand(k, 1, generator=g) * 0.7 * radius
    rk = torch.einsum('aij,kj->aki', anchors, base).contiguous().to(dev)
    feats = torch.randn(b, n, 60, ci, generator=g).to(dev)
    dg = torch.randn(b, p, 60, k * ci, generator=g).to(dev)
    out, gin = {}, {}
    for mode in (0, 3):
        gg = torch.full((b, p, 60, k * ci), float('nan'), device=dev)
        call("vgtkb_inter_group_forward", dev, b, n, p, nn, 60, k, ci, ptr(xyz), ptr(sxyz), ptr(idx), ptr(rk), sigma, ptr(feats),
             ptr(gg), mode)
        gx = torch.zeros(b, n, 60, ci, device=dev)
        call("vgtkb_inter_group_backward", dev, b, n, p, nn, 60, k, ci, ptr(xyz), ptr(sxyz), ptr(idx), ptr(rk), sigma, ptr(dg),
             ptr(gx), mode)
        out[mode], gin[mode] = gg, gx
    assert torch.isfinite(out[3]).all()
    assert rel_err(out[3], out[0]) < 3e-5 and rel_err(gin[3], gin[0]) < 3e-5
    if nn == 7:     # tie the fp32 kernel itself to the oracle's restatement of the reference einsum
        gxyz = torch.gather(xyz.cpu(), 2, idx.cpu().long().view(b, 1, p * nn).expand(b, 3, p * nn)).view(b, 3, p, nn) \
            - sxyz.cpu().unsqueeze(-1)
        w = torch.relu(1.0 - ((gxyz[:, :, :, None, None, :] - rk.cpu().permute(2, 0, 1)[None, :, None, :, :, None]) ** 2).sum(1) / sigma)
        f = feats.cpu()[torch.arange(b)[:, None, None], idx.cpu().long()]            # [b,p,nn,a,c]
        ref = torch.einsum('bpakn,bpnac->bpakc', w, f).reshape(b, p, 60, k * ci)
        assert rel_err(out[0], ref) < 1e-5
