"""Block-level behaviours outside the default benchmark configuration (GPU):
  * dropout > 0 in a separable block: the reference drops out relu(norm(intra)) and only then adds the skip branch
    (SPConvNets/utils/base_so3conv.py:59-64,210-217) -- the fused norm+act+residual pass must not be used then;
  * backward through an eval-mode (frozen statistics) BatchNorm, as torch's BatchNorm2d supports;
  * momentum=None BatchNorm keeps torch's cumulative moving average;
  * a reloaded state dict invalidates the cached rotated kernels / intra tables."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    from equi_articulated_pose_b200 import lib
    lib.load()
    return torch.device("cuda:0")


def _params(dropout):
    return {'dim_in': 64, 'dim_out': 64, 'kernel_size': 1, 'stride': 1, 'radius': 0.45, 'sigma': 0.08, 'n_neighbor': 16,
            'lazy_sample': True, 'dropout_rate': dropout, 'multiplier': 2, 'activation': 'leaky_relu', 'pooling': None,
            'kanchor': 60, 'norm': 'BatchNorm2d'}


def test_dropout_is_applied_before_the_skip_connection(dev):
    from equi_articulated_pose_b200 import blocks
    from oracle import so3 as O
    import vgtk.spconv as zptk
    torch.manual_seed(0)
    blk = blocks.SeparableSO3ConvBlock(_params(0.5)).to(dev).train()
    ref = blocks.SeparableSO3ConvBlock(_params(0.0)).to(dev).train()
    ref.load_state_dict(blk.state_dict())
    xyz = O.synthetic_cloud(2, 96, 3).permute(0, 2, 1).contiguous().to(dev)
    feats = torch.randn(2, 96, 60, 64, device=dev).permute(0, 3, 1, 2)
    # reference composition from the dropout-free block's parts, drawing the two dropout masks (inter block, intra block:
    # base_so3conv.py:125-131, 59-64) in the same order from the same seed:
    #   y1 = dropout(lrelu(BN(inter conv)));  out = dropout(lrelu(IN(intra conv(y1)))) + lrelu(BN(skip conv))
    from equi_articulated_pose_b200 import ops
    torch.manual_seed(123)
    _, _, _, y = ref.inter_conv(zptk.SphericalPointCloud(xyz, feats, None), None, None)
    r1, _ = blocks._rows(y.feats)
    y1 = zptk.SphericalPointCloud(y.xyz, blocks._unrows(F.dropout(r1, 0.5, True), 2, 96, 60), y.anchors)
    srows = feats.permute(0, 2, 3, 1).reshape(-1, 64)
    srows = ops.LinearFn.apply(srows, ref.skip_conv.weight.view(64, 64), ref.skip_conv.bias)
    skip = blocks._unrows(blocks._apply_norm(ref.norm, srows, 2, ref.slope), 2, 96, 60)
    r2, _ = blocks._rows(ref.intra_conv(y1).feats)                    # leaky_relu(IN(intra conv)), no residual
    want = blocks._unrows(F.dropout(r2, 0.5, True), 2, 96, 60) + skip
    torch.manual_seed(123)
    _, _, _, out = blk(zptk.SphericalPointCloud(xyz, feats, None), None, None)
    assert float((out.feats - want).abs().max()) <= 1e-5 * float(want.abs().max())
    # the skip path survives where dropout zeroed the intra branch: out == skip exactly there
    dropped = (out.feats - skip).abs() < 1e-12
    assert 0.3 < float(dropped.float().mean()) < 0.7


def test_eval_mode_batchnorm_backward(dev):
    from equi_articulated_pose_b200 import ops
    g = torch.Generator().manual_seed(4)
    x = torch.randn(1, 3000, 64, generator=g)
    gy = torch.randn(1, 3000, 64, generator=g)
    gamma, beta = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.1
    rm, rv = torch.randn(64, generator=g) * 0.2, torch.rand(64, generator=g) + 0.5
    xd = x.double().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    ref = F.leaky_relu(F.batch_norm(xd.permute(0, 2, 1), rm.double(), rv.double(), gd, bd, False, 0.1, 1e-5), 0.01).permute(0, 2, 1)
    ref.backward(gy.double())
    xg = x.to(dev).requires_grad_(True)
    gg, bg = gamma.to(dev).requires_grad_(True), beta.to(dev).requires_grad_(True)
    y = ops.norm_act(xg, gg, bg, None, rm.to(dev), rv.to(dev), 0.1, 1e-5, 0.01, use_running=True)
    y.backward(gy.to(dev))
    for got, want in ((y, ref), (xg.grad, xd.grad), (gg.grad, gd.grad), (bg.grad, bd.grad)):
        assert float((got.detach().cpu().double() - want.detach()).abs().max()) <= 2e-5 * float(want.detach().abs().max())


def test_momentum_none_is_a_cumulative_average(dev):
    from equi_articulated_pose_b200 import blocks
    torch.manual_seed(1)
    ours = blocks.FusedBatchNorm2d(16, momentum=None).to(dev).train()
    ref = torch.nn.BatchNorm2d(16, momentum=None).to(dev).train()
    for i in range(3):
        x = torch.randn(4, 16, 50, 6, device=dev) * (i + 1) + i
        ref(x)
        ours.forward_rows(x.permute(0, 2, 3, 1).reshape(-1, 16).contiguous(), 0.01)
    assert torch.allclose(ours.running_mean, ref.running_mean, atol=1e-5)
    assert torch.allclose(ours.running_var, ref.running_var, rtol=1e-4, atol=1e-5)
    assert int(ours.num_batches_tracked) == 3


def test_cached_tables_follow_a_reloaded_state_dict(dev):
    import equi_articulated_pose_b200 as pkg
    pkg.install()
    import vgtk.so3conv as sptk
    conv = sptk.InterSO3Conv(32, 32, 1, 1, 0.4, 0.08, 16).to(dev)
    rk0 = conv.rot_kernels().clone()
    sd = {k: v.clone() for k, v in conv.state_dict().items()}
    sd["kernels"] = sd["kernels"] * 2.0
    conv.load_state_dict(sd)
    assert torch.allclose(conv.rot_kernels(), 2.0 * rk0)
    intra = sptk.IntraSO3Conv(64, 64).to(dev)
    t0 = intra.tables()[0].clone()
    sd = {k: v.clone() for k, v in intra.state_dict().items()}
    sd["intra_idx"] = sd["intra_idx"].flip(1)
    intra.load_state_dict(sd)
    assert torch.equal(intra.tables()[0], t0.flip(1))


@pytest.mark.parametrize("stride", [1, 2])
def test_gradient_hand_over_between_skip_branch_and_inter_conv(dev, monkeypatch, stride):
    """ops.GradSlot: the skip branch deposits its input gradient and the inter conv's scatter adds onto it
    (vgtkb_inter_conv_backward mode | 256).  Same block, same inputs, hand-over on / off: identical outputs, input and
    parameter gradients equal up to the order of the floating-point additions; and the deposit really happened."""
    from equi_articulated_pose_b200 import blocks, ops
    from oracle import so3 as O
    import vgtk.spconv as zptk
    p = _params(0.0)
    p['stride'] = stride
    if stride > 1:
        p['n_neighbor'] = 32
    torch.manual_seed(1)
    blk = blocks.SeparableSO3ConvBlock(p).to(dev).train()
    xyz = O.synthetic_cloud(2, 128, 5).permute(0, 2, 1).contiguous().to(dev)
    base = torch.randn(2, 128, 60, 64, device=dev)
    taken = []
    orig = ops.GradSlot.deposit

    def spy(self, g):
        ok = orig(self, g)
        taken.append(ok)
        return ok
    monkeypatch.setattr(ops.GradSlot, "deposit", spy)
    res = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("VGTKB_GRAD_SLOT", flag)
        blk.zero_grad(set_to_none=True)
        x = base.clone().requires_grad_(True)
        _, _, _, out = blk(zptk.SphericalPointCloud(xyz, x.permute(0, 3, 1, 2), None), None, None)
        go = torch.randn(out.feats.shape, generator=torch.Generator().manual_seed(9)).to(dev)
        (out.feats * go).sum().backward()
        res[flag] = (out.feats.detach().clone(), x.grad.clone(), {k: v.grad.clone() for k, v in blk.named_parameters()})
    assert taken == [True]                                   # one deposit with the hand-over on, none with it off
    assert torch.equal(res["1"][0], res["0"][0])
    s = float(res["0"][1].abs().max())
    assert float((res["1"][1] - res["0"][1]).abs().max()) <= 2e-6 * s
    for k in res["0"][2]:            # (the weight gradients are split-R atomic sums: equal up to the addition order)
        a, b = res["1"][2][k], res["0"][2][k]
        assert float((a - b).abs().max()) <= 1e-5 * max(float(b.abs().max()), 1e-30), k
