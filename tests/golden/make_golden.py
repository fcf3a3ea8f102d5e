"""Generate the committed golden fixtures by RUNNING THE REFERENCE'S OWN PYTHON in place.

    python tests/golden/make_golden.py          (needs /root/reference; CPU only)

The reference ships no golden vectors for this path, so parity is
pinned on outputs of the reference itself, imported through oracle/ref_harness.py.
Writes (all small):
  equi_articulated_pose_b200/data/so3_constants.npz  anchors Rs [60,3,3] f32, intra_idx [60,12] i64,
                                                     kpsphere24 [24,3] f32 (+ icosahedron verts/faces)
  tests/golden/ref_blocks_small.npz   2 stacked separable blocks, fwd + bwd, train mode
  tests/golden/ref_intra_small.npz    one IntraSO3Conv (BASELINE config 1a, reduced size)
  tests/golden/ref_weights_small.npz  inter_so3conv_grouping_anchor + grouping einsum
  tests/golden/ref_intrazp_small.npz, ref_pose_group_small.npz, ref_pointnet_small.npz (`make_golden.py pointnet`)
  tests/golden/ref_pointnet2_small.npz  PointnetPP encoder-decoder (SPConvNets/models/PointNet2.py), fwd (`make_golden.py pointnet2`)
  tests/golden/ref_pose_group_strided_small.npz  strided branch of inter_so3poseconv_grouping_strided (`make_golden.py pose_strided`)
  tests/golden/ref_heads_small.npz  invariant heads of base_so3conv.py (InvOutBlockR / Pointnet / MVD / Ours / OursWithMask), fwd + bwd (`make_golden.py heads`)
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as H  # noqa: E402
from oracle import so3 as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "equi_articulated_pose_b200", "data")


def small_params():
    """Two separable layers in one BasicSO3ConvBlock + one strided block; N=64."""
    def layer(ci, co, stride, radius, sigma, nn, lazy):
        return {'type': 'separable_block', 'args': {
            'dim_in': ci, 'dim_out': co, 'kernel_size': 1, 'stride': stride, 'radius': radius, 'sigma': sigma,
            'n_neighbor': nn, 'lazy_sample': lazy, 'dropout_rate': 0.0, 'multiplier': 2,
            'activation': 'leaky_relu', 'pooling': None, 'kanchor': 60, 'norm': 'BatchNorm2d'}}
    return [[layer(1, 8, 2, 0.6, 0.18, 16, False), layer(8, 8, 1, 0.8, 0.32, 8, True)],
            [layer(8, 16, 2, 1.0, 0.5, 12, True)]]


def make_pointnet():
    """PointnetSO3Conv head (so3conv/modules.py:376-413), fwd + bwd, A = 60 and A = 1, raw and pooled."""
    H.import_blocks()
    import vgtk.so3conv as sptk
    out = {}
    for na in (60, 1):
        torch.manual_seed(0)
        g = torch.Generator().manual_seed(1004 + na)
        nb, c, npt, co = 2, 16, 40, 24
        head = sptk.PointnetSO3Conv(c, co, kanchor=na)
        feats = torch.randn(nb, c, npt, na, generator=g, requires_grad=True)
        xyz = torch.rand(nb, 3, npt, generator=g) - 0.5
        pooled = head(sptk.SphericalPointCloud(xyz, feats, None))
        gout = torch.randn(pooled.shape, generator=g)
        (pooled * gout).sum().backward()
        head.return_raw = True
        raw = head(sptk.SphericalPointCloud(xyz, feats.detach(), None))
        out.update({f'a{na}_weight': head.embed.weight.detach().numpy(), f'a{na}_bias': head.embed.bias.detach().numpy(),
                    f'a{na}_feats': feats.detach().numpy(), f'a{na}_xyz': xyz.numpy(), f'a{na}_pooled': pooled.detach().numpy(),
                    f'a{na}_raw': raw.detach().numpy(), f'a{na}_grad_out': gout.numpy(), f'a{na}_grad_feats': feats.grad.numpy(),
                    f'a{na}_grad_weight': head.embed.weight.grad.numpy(), f'a{na}_grad_bias': head.embed.bias.grad.numpy()})
    np.savez_compressed(os.path.join(GOLD, "ref_pointnet_small.npz"), **out)
    print("pointnet", out['a60_pooled'].shape, out['a1_pooled'].shape)


def make_pose_strided():
    """Strided branch of inter_so3poseconv_grouping_strided (so3conv/functional.py:931-1029) with random per-point
    rotations: stride 2, FPS sampling (lazy_sample False) and lazy sampling, permute_modes 0 and 1."""
    H.import_blocks()
    import vgtk.so3conv.functional as L
    from scipy.spatial.transform import Rotation
    anc = torch.from_numpy(L.get_anchors(60))
    g = torch.Generator().manual_seed(1005)
    nb, npt, nnb, cc = 1, 32, 6, 5
    pxyz = (torch.rand(nb, 3, npt, generator=g) - 0.5)
    rot = torch.from_numpy(Rotation.random(nb * npt, random_state=11).as_matrix().astype('float32')).view(nb, npt, 3, 3)
    ppose = torch.eye(4).repeat(nb, npt, 1, 1)
    ppose[:, :, :3, :3] = rot
    pfeats = torch.randn(nb, cc, npt, 60, generator=g)
    pk = torch.from_numpy(L.get_sphereical_kernel_points_from_ply(0.7 * 0.5, 1))
    outp = {'xyz': pxyz.numpy(), 'pose': ppose.numpy(), 'feats': pfeats.numpy(), 'kernels': pk.numpy(),
            'radius': np.float32(0.5), 'sigma': np.float32(0.12), 'nn': np.int32(nnb), 'stride': np.int32(2)}
    for lazy in (0, 1):
        for pm in ((0, 1) if lazy == 0 else (1,)):
            with contextlib.redirect_stdout(io.StringIO()):
                r = L.inter_so3poseconv_grouping_strided(pxyz, ppose, pfeats, 2, nnb, anc, pk, 0.5, 0.12, None, None, bool(lazy),
                                                         permute_modes=pm)
            outp[f'grouped_lazy{lazy}_pm{pm}'] = r[3].numpy()
            outp[f'sample_idx_lazy{lazy}'] = r[4].numpy()
            outp[f'new_xyz_lazy{lazy}'] = r[2].numpy()
            outp[f'sampled_pose_lazy{lazy}'] = r[5].numpy()
    np.savez_compressed(os.path.join(GOLD, "ref_pose_group_strided_small.npz"), **outp)
    print("pose grouping (strided)", r[3].shape)


HEAD_CASES = [
    # name, class, constructor kwargs, params overrides
    ("r_att", "InvOutBlockR", {}, {"pooling": "attention"}),
    ("r_max", "InvOutBlockR", {}, {"pooling": "max"}),
    ("pn_max", "InvOutBlockPointnet", {}, {"pooling": "max"}),
    ("mvd", "InvOutBlockMVD", {}, {}),
    ("ours_max", "InvOutBlockOurs", {"pooling_method": "max"}, {}),
    ("mask_att", "InvOutBlockOursWithMask", {"norm": 1, "pooling_method": "attention", "use_pointnet": True}, {}),
]


def head_inputs():
    g = torch.Generator().manual_seed(1006)
    nb, c, npt, na = 2, 16, 16, 60
    feats = torch.randn(nb, c, npt, na, generator=g)
    xyz = torch.rand(nb, 3, npt, generator=g) - 0.5
    mask = (torch.rand(nb, npt, generator=g) > 0.3).float()
    soft = torch.rand(nb, npt, generator=g)
    return feats, xyz, mask, soft


def head_params(over):
    p = {"dim_in": 16, "mlp": [32, 64], "fc": [64], "k": 8, "kanchor": 60, "temperature": 3.0}
    p.update(over)
    return p


def make_heads():
    """The invariant heads (SPConvNets/utils/base_so3conv.py:481-645, 766-840, 1013-1150), train mode, fwd + bwd on the
    reference's own modules; every tensor output contributes sum(out * fixed random tensor) to the loss."""
    M = H.import_blocks()
    import vgtk.so3conv as sptk
    feats0, xyz, mask, soft = head_inputs()
    out = {"feats": feats0.numpy(), "xyz": xyz.numpy(), "mask": mask.numpy(), "soft_mask": soft.numpy()}
    anc = torch.from_numpy(sptk.get_anchors(60))
    for name, cls, kw, over in HEAD_CASES:
        torch.manual_seed(100 + len(name))
        with contextlib.redirect_stdout(io.StringIO()):
            head = getattr(M, cls)(head_params(over), **kw)
        head.train()
        feats = feats0.clone().requires_grad_(True)
        if cls == "InvOutBlockR":
            res = head(feats)
        elif cls == "InvOutBlockOursWithMask":
            res = head(sptk.SphericalPointCloud(xyz, feats, anc), mask, soft_mask=soft)
        else:
            res = head(sptk.SphericalPointCloud(xyz, feats, anc))
        res = res if isinstance(res, tuple) else (res,)
        gg = torch.Generator().manual_seed(7)
        loss = 0
        for i, r in enumerate(res):
            w = torch.randn(r.shape, generator=gg)
            loss = loss + (r * w).sum()
            out[f"{name}_out{i}"] = r.detach().numpy()
        loss.backward()
        out[f"{name}_grad_feats"] = feats.grad.numpy()
        for k, v in head.state_dict().items():
            out[f"{name}_sd_{k}"] = v.numpy()
        for k, v in head.named_parameters():
            out[f"{name}_pg_{k}"] = v.grad.numpy() if v.grad is not None else np.zeros(0, np.float32)
    np.savez_compressed(os.path.join(GOLD, "ref_heads_small.npz"), **out)
    print("heads", [n for n, *_ in HEAD_CASES])


def make_pointnet2():
    """PointnetPP (SPConvNets/models/PointNet2.py:8-196), train-mode forward with return_global=True on two clouds of 600
    points, in_feat_dim = 6 (x = 3 extra channels).  Weights come from oracle.pointnet2.make_state (numpy RandomState, so the
    tests rebuild them); torch_cluster.fps is a stub over the oracle's plain FPS (dependency not vendored)."""
    import math
    import types
    from oracle import cops
    from oracle import pointnet2 as OP
    tc = types.ModuleType("torch_cluster")

    def fps(src, batch=None, ratio=0.5, random_start=True):
        assert not random_start
        nb = int(batch[-1]) + 1
        n = src.shape[0] // nb
        m = int(math.ceil(ratio * n))
        xyz = src[:, :3].float().reshape(nb, n, 3).permute(0, 2, 1).contiguous().numpy()
        idx = torch.from_numpy(cops.fps_plain(xyz, m)).long()
        return (idx + torch.arange(nb).view(nb, 1) * n).reshape(-1)
    tc.fps = fps
    sys.modules["torch_cluster"] = tc
    H.import_blocks()
    import importlib
    P = importlib.import_module("SPConvNets.models.PointNet2")
    net = P.PointnetPP(in_feat_dim=6)
    net.load_state_dict(OP.make_state(6, seed=7))
    net.train()
    g = torch.Generator().manual_seed(6001)
    pos = torch.rand(2, 600, 3, generator=g) - 0.5
    x = torch.randn(2, 600, 3, generator=g)
    taps = {}
    orig = net.sample_and_group

    def tapped(feat, p, n_samples, use_pos=True, k=64):
        r = orig(feat, p, n_samples, use_pos=use_pos, k=k)
        taps[f"topk_dist_{n_samples}"] = r[1].detach().numpy().copy()
        taps[f"pos_{n_samples}"] = r[2].detach().numpy().copy()
        return r
    net.sample_and_group = tapped
    out, glb, p_out = net(x.clone(), pos.clone(), return_global=True)
    res = {"pos": pos.numpy(), "x": x.numpy(), "out": out.detach().numpy(), "global_x": glb.detach().numpy(),
           "pos_out": p_out.numpy(), "seed": np.int64(7)}
    res.update(taps)
    for name in ("mlp_layers.0.2.1", "mlp_layers.2.2.1", "up_mlp_layers.2.2.1"):
        sd = net.state_dict()
        res["rm:" + name] = sd[name + ".running_mean"].numpy()
        res["rv:" + name] = sd[name + ".running_var"].numpy()
    try:                        # the reference masks the MLP output in place (:109): its own backward rejects that
        out.square().mean().backward()
        res["reference_backward"] = np.array("ok")
    except RuntimeError as e:
        res["reference_backward"] = np.array("RuntimeError: " + str(e).splitlines()[0][:160])
    np.savez_compressed(os.path.join(GOLD, "ref_pointnet2_small.npz"), **res)
    print("pointnet2", out.shape, glb.shape, str(res["reference_backward"]))


def main():
    torch.set_num_threads(os.cpu_count())
    if len(sys.argv) > 1 and sys.argv[1] == "heads":
        return make_heads()
    if len(sys.argv) > 1 and sys.argv[1] == "pose_strided":
        return make_pose_strided()
    if len(sys.argv) > 1 and sys.argv[1] == "pointnet2":
        return make_pointnet2()
    if len(sys.argv) > 1 and sys.argv[1] == "pointnet":      # only this fixture (the others stay untouched)
        return make_pointnet()
    M = H.import_blocks()
    import vgtk.so3conv as sptk
    import vgtk.so3conv.functional as L
    import vgtk.spconv as zptk
    import vgtk.pc as pctk

    os.makedirs(DATA, exist_ok=True)
    # ---- constants ------------------------------------------------------------------
    Rs = L.get_anchors(60)
    Ri = L.get_intra_idx()
    anchors_dir = os.path.join(H.REF, "vgtk", "vgtk", "data", "anchors")
    kp24 = pctk.load_ply(os.path.join(anchors_dir, "kpsphere24.ply")).astype("float32")
    v12, f12 = H.read_ply(os.path.join(anchors_dir, "sphere12.ply"))
    np.savez(os.path.join(DATA, "so3_constants.npz"), anchors=Rs.astype(np.float32), intra_idx=Ri.astype(np.int64),
             kpsphere24=kp24, ico_vertices=v12, ico_faces=f12)
    print("constants", Rs.shape, Ri.shape, kp24.shape)

    # ---- weights + grouping ---------------------------------------------------------
    g = torch.Generator().manual_seed(11)
    gxyz = (torch.rand(2, 3, 5, 7, generator=g) - 0.5) * 0.5
    kern = torch.from_numpy(L.get_sphereical_kernel_points_from_ply(0.7 * 0.4, 1))
    anc = torch.from_numpy(Rs)
    w = L.inter_so3conv_grouping_anchor(gxyz, anc, kern, 0.08)
    feats = torch.randn(2, 3, 9, 60, generator=g)
    idx = torch.randint(0, 9, (2, 5, 7), generator=g, dtype=torch.int32)
    G = zptk.inter_zpconv_grouping_naive(idx, w, zptk.functional.add_shadow_feature(feats))
    np.savez(os.path.join(GOLD, "ref_weights_small.npz"), grouped_xyz=gxyz.numpy(), kernels=kern.numpy(),
             sigma=np.float32(0.08), inter_w=w.numpy(), feats=feats.numpy(), idx=idx.numpy(), grouped=G.numpy())

    # ---- single intra conv (config 1a, reduced) -------------------------------------
    torch.manual_seed(0)
    conv = sptk.IntraSO3Conv(16, 24)
    f = torch.randn(1, 16, 32, 60, generator=g)
    x = zptk.SphericalPointCloud(torch.rand(1, 3, 32, generator=g) - 0.5, f, None)
    out = conv(x).feats
    np.savez(os.path.join(GOLD, "ref_intra_small.npz"), W=conv.basic_conv.W.detach().numpy(), feats=f.numpy(),
             out=out.detach().numpy())

    # ---- stacked separable blocks, fwd + bwd ----------------------------------------
    params = small_params()
    torch.manual_seed(1)
    with contextlib.redirect_stdout(io.StringIO()):
        backbone = torch.nn.ModuleList([M.BasicSO3ConvBlock(p) for p in params])
    with torch.no_grad():
        for n, p in backbone.named_parameters():
            if n.endswith('norm.weight'):
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))
            if n.endswith('norm.bias'):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
    backbone.train()
    pts = O.synthetic_cloud(2, 64, seed=2001)
    xyz = pts.permute(0, 2, 1).contiguous()
    x = zptk.SphericalPointCloud(xyz, sptk.get_occupancy_features(pts, 60, False), None)
    state = {('backbone.' + k): v.detach().clone().numpy() for k, v in backbone.state_dict().items()
             if not k.endswith(('anchors', 'kernels', 'intra_idx', 'num_batches_tracked'))}
    for blk in backbone:
        x = blk(x)
    loss = x.feats.square().mean()
    loss.backward()
    out = {'in_points': pts.numpy(), 'out_feats': x.feats.detach().numpy(), 'out_xyz': x.xyz.detach().numpy(),
           'loss': loss.detach().numpy()}
    for k, v in state.items():
        out['state/' + k] = v
    for n, p in backbone.named_parameters():
        out['grad/backbone.' + n] = p.grad.numpy()
    for k, v in backbone.state_dict().items():
        if k.endswith(('running_mean', 'running_var')):
            out['after/backbone.' + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLD, "ref_blocks_small.npz"), **out)
    print("blocks", x.feats.shape, float(loss))

    # ---- legacy S^2 IntraZPConv (BASELINE config 1b: A = 12 direction anchors, N = 128, batch 1) ------
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(1002)       # own stream: the fixtures above stay bit-identical
    zp = zptk.IntraZPConv(dim_in=32, dim_out=32, kernel_size=3, aperture=1.0, sigma=0.1, anchor_nn=6, anchor_in=12)
    fz = torch.randn(1, 32, 128, 12, generator=g, requires_grad=True)
    oz = zp(zptk.SphericalPointCloud(torch.rand(1, 3, 128, generator=g) - 0.5, fz, None)).feats
    goz = torch.randn(oz.shape, generator=g)
    (oz * goz).sum().backward()
    np.savez_compressed(os.path.join(GOLD, "ref_intrazp_small.npz"), W=zp.basic_conv.W.detach().numpy(),
                        bias=zp.basic_conv.bias.detach().numpy(), intra_idx=zp.intra_idx.numpy(), intra_w=zp.intra_w.numpy(),
                        anchors=zp.anchor_out.numpy(), kernels=zp.kernels.numpy(), feats=fz.detach().numpy(),
                        out=oz.detach().numpy(), grad_out=goz.numpy(), grad_feats=fz.grad.numpy(),
                        grad_W=zp.basic_conv.W.grad.numpy(), grad_bias=zp.basic_conv.bias.grad.numpy())
    print("intrazp", oz.shape)

    # ---- pose-aware inter grouping with NON-identity per-point rotations (no-stride branch) -------------
    from scipy.spatial.transform import Rotation
    g = torch.Generator().manual_seed(1003)
    nb, npt, nnb, cc = 1, 24, 6, 5
    pxyz = (torch.rand(nb, 3, npt, generator=g) - 0.5)
    rot = torch.from_numpy(Rotation.random(nb * npt, random_state=7).as_matrix().astype('float32')).view(nb, npt, 3, 3)
    ppose = torch.eye(4).repeat(nb, npt, 1, 1)
    ppose[:, :, :3, :3] = rot
    pfeats = torch.randn(nb, cc, npt, 60, generator=g)
    pk = torch.from_numpy(L.get_sphereical_kernel_points_from_ply(0.7 * 0.5, 1))
    outp = {'xyz': pxyz.numpy(), 'pose': ppose.numpy(), 'feats': pfeats.numpy(), 'kernels': pk.numpy(),
            'radius': np.float32(0.5), 'sigma': np.float32(0.12), 'nn': np.int32(nnb)}
    for pm in (0, 1):
        with contextlib.redirect_stdout(io.StringIO()):
            r = L.inter_so3poseconv_grouping_strided(pxyz, ppose, pfeats, 1, nnb, anc, pk, 0.5, 0.12, None, None, True,
                                                     permute_modes=pm)
        outp[f'grouped_pm{pm}'] = r[3].numpy()
    np.savez_compressed(os.path.join(GOLD, "ref_pose_group_small.npz"), **outp)
    print("pose grouping", r[3].shape)
    make_pointnet()


if __name__ == "__main__":
    main()
