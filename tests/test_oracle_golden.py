"""CPU suite: pin the oracle (oracle/) against the committed outputs of the reference's own
Python (tests/golden/, produced by tests/golden/make_golden.py) and against the derived
known-answer tests of SURVEY.md appendix C."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import cops, so3 as O
from equi_articulated_pose_b200 import so3_constants as C


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


# ---------------------------------------------------------------- constants (appendix C.1, C.2)
def test_anchor_fingerprints():
    Rs, Ri = C.anchors_all(), C.intra_idx()
    assert Rs.dtype == np.float32 and Rs.shape == (60, 3, 3)
    assert hashlib.sha256(Rs.tobytes()).hexdigest()[:16] == "a2c584147246900d"
    assert hashlib.sha256(Ri.astype(np.int64).tobytes()).hexdigest()[:16] == "d1c64694466d57e7"
    assert (Rs[29] == np.eye(3, dtype=np.float32)).all()
    assert Ri[0].tolist() == [10, 33, 13, 15, 56, 30, 59, 8, 44, 0, 45, 26]
    assert Ri[29].tolist() == [44, 31, 32, 13, 42, 43, 30, 14, 12, 29, 28, 27]
    assert int(Ri.sum()) == 21240


def test_anchor_group_structure():
    Rs, Ri = C.anchors_all().astype(np.float64), C.intra_idx()
    prod = np.einsum('aij,bjk->abik', Rs, Rs).reshape(3600, 9)
    d = ((prod[:, None, :] - Rs.reshape(1, 60, 9)) ** 2).sum(-1).min(1)
    assert d.max() < 1e-10                                     # closed under multiplication
    for k in range(12):
        assert sorted(Ri[:, k].tolist()) == list(range(60))    # every column is a permutation
    assert (Ri[:, 9] == np.arange(60)).all()                   # column 9 = identity
    assert all(len(set(r.tolist())) == 12 for r in Ri)


def test_anchor_derivation_matches_table():
    assert np.abs(C.derive_anchor_group() - C.anchors_all()).max() < 1e-6
    Ri = C.intra_idx()
    assert (C.derive_intra_idx(C.anchors_all().astype(np.float64), Ri[0]) == Ri).all()


def test_select_anchor():
    assert C.get_anchors(1).shape == (1, 3, 3) and (C.get_anchors(1)[0] == np.eye(3)).all()
    assert C.get_anchors(20).shape == (20, 3, 3) and C.get_anchors(40).shape == (40, 3, 3)


def test_kernel_points(golden_dir):
    g = _load(golden_dir, "ref_weights_small.npz")
    assert np.array_equal(O.scaled_kernel_points(C.kernel_points_base(), 0.4), g["kernels"])
    assert np.array_equal(C.scaled_kernel_points(0.7 * 0.4), g["kernels"])


# ---------------------------------------------------------------- opt_n_threads (appendix C.6)
def test_opt_n_threads_matches_integer_rule():
    for n, want in ((380, 256), (512, 512), (1000, 512), (1024, 1024), (4096, 1024), (1, 1), (3, 2)):
        assert cops.opt_n_threads(n) == want
    for n in range(1, 20000):
        p = 1
        while p * 2 <= n and p < 1024:
            p *= 2
        assert cops.opt_n_threads(n) == p, n


# ---------------------------------------------------------------- index ops: semantics by hand
def test_ball_query_padding_rules():
    xyz = np.zeros((1, 3, 6), np.float32)
    xyz[0, 0] = [0.0, 0.05, 0.5, 0.06, 0.07, 0.9]
    q = np.zeros((1, 3, 1), np.float32)
    # 4 hits (0,1,3,4) with nsample=8 -> cyclic repetition
    assert cops.ball_query(q, xyz, 0.1, 8)[0, 0].tolist() == [0, 1, 3, 4, 0, 1, 3, 4]
    # nsample=5: cnt == nsample-1 -> last slot stays 0
    assert cops.ball_query(q, xyz, 0.1, 5)[0, 0].tolist() == [0, 1, 3, 4, 0]
    # nsample=3: first three in index order
    assert cops.ball_query(q, xyz, 0.1, 3)[0, 0].tolist() == [0, 1, 3]
    # no hit at all -> zeros; strict '<'
    far = np.full((1, 3, 1), 5.0, np.float32)
    assert cops.ball_query(far, xyz, 0.1, 4)[0, 0].tolist() == [0, 0, 0, 0]
    q2 = np.zeros((1, 3, 1), np.float32); q2[0, 0, 0] = 0.25
    x2 = np.zeros((1, 3, 2), np.float32); x2[0, 0] = [0.0, 0.5]
    assert cops.ball_query(q2, x2, 0.25, 2)[0, 0].tolist() == [0, 0]      # d2 == r2 is not a hit


def test_fps_rules():
    g = np.random.default_rng(0)
    xyz = g.normal(size=(2, 3, 100)).astype(np.float32)
    idx = cops.furthest_point_sampling(xyz, 20)
    assert (idx[:, 0] == 0).all()
    # brute-force restatement without the block structure (valid when there are no exact ties)
    for b in range(2):
        P = xyz[b].T
        temp = np.full(100, 1e10, np.float32)
        old, out = 0, [0]
        for _ in range(19):
            d = ((P - P[old]) ** 2).sum(1).astype(np.float32)
            temp = np.minimum(temp, d)
            old = int(temp.argmax()); out.append(old)
        assert out == idx[b].tolist()
    # points with |p|^2 <= 1e-3 are never selected (except the forced start index 0)
    xyz[0, :, 5:50] *= 1e-3
    idx = cops.furthest_point_sampling(xyz, 40)
    assert not set(idx[0, 1:].tolist()) & set(range(5, 50))
    # all points skipped -> the tree returns index 0 forever
    z = np.zeros((1, 3, 16), np.float32)
    assert cops.furthest_point_sampling(z, 5)[0].tolist() == [0, 0, 0, 0, 0]


def test_fps_tie_break_is_bit_reversed_thread_order():
    # 8 points, block size 8: points 1 and 4 are equidistant from point 0 and everything else is closer.
    xyz = np.zeros((1, 3, 8), np.float32)
    xyz[0, 0] = [1.0, 3.0, 1.2, 1.3, -1.0, 1.1, 1.4, 1.5]     # |x - 1| = 2 for k=1 and k=4
    xyz[0, 1] = 0.5
    idx = cops.furthest_point_sampling(xyz, 2)
    # tree: stride 4 pairs (0,4)(1,5).. -> slot0 holds k=4, slot1 holds k=1; final stride 1: tie -> left -> k=4
    assert idx[0].tolist() == [0, 4]


def test_gather_and_chamfer_small():
    g = np.random.default_rng(1)
    pts = g.normal(size=(2, 3, 7)).astype(np.float32)
    idx = g.integers(0, 7, size=(2, 5)).astype(np.int32)
    out = cops.gather_points_forward(pts, idx)
    assert np.array_equal(out, np.take_along_axis(pts, idx[:, None, :].repeat(3, 1), 2))
    back = cops.gather_points_backward(out, idx, 7)
    ref = np.zeros_like(pts)
    for b in range(2):
        for j in range(5):
            ref[b, :, idx[b, j]] += out[b, :, j]
    assert np.allclose(back, ref)
    a = g.normal(size=(2, 9, 3)).astype(np.float32)
    c = g.normal(size=(2, 11, 3)).astype(np.float32)
    d1, d2, i1, i2 = cops.chamfer_forward(a, c)
    D = ((a[:, :, None, :] - c[:, None, :, :]) ** 2).sum(-1)
    assert np.array_equal(i1, D.argmin(2)) and np.array_equal(i2, D.argmin(1))
    assert np.allclose(d1, D.min(2), rtol=1e-6) and np.allclose(d2, D.min(1), rtol=1e-6)
    # duplicate targets: the lowest index wins
    c[:, 5] = c[:, 2]
    _, _, i1, _ = cops.chamfer_forward(a, c)
    assert not (i1 == 5).any()


def test_chamfer_backward_matches_autograd():
    g = torch.Generator().manual_seed(3)
    a = torch.randn(2, 6, 3, generator=g, requires_grad=True)
    c = torch.randn(2, 8, 3, generator=g, requires_grad=True)
    D = ((a[:, :, None] - c[:, None]) ** 2).sum(-1)
    w1, w2 = torch.randn(2, 6, generator=g), torch.randn(2, 8, generator=g)
    ((D.min(2)[0] * w1).sum() + (D.min(1)[0] * w2).sum()).backward()
    _, _, i1, i2 = cops.chamfer_forward(a.detach().numpy(), c.detach().numpy())
    g1, g2 = cops.chamfer_backward(a.detach().numpy(), c.detach().numpy(), i1, i2, w1.numpy(), w2.numpy())
    assert np.allclose(g1, a.grad.numpy(), atol=1e-5) and np.allclose(g2, c.grad.numpy(), atol=1e-5)


# ---------------------------------------------------------------- float path vs the reference's outputs
def test_anchor_weights_and_grouping_match_reference(golden_dir):
    g = _load(golden_dir, "ref_weights_small.npz")
    anchors = torch.from_numpy(C.anchors_all())
    w = O.anchor_weights(torch.from_numpy(g["grouped_xyz"]), anchors, torch.from_numpy(g["kernels"]), float(g["sigma"]))
    assert torch.allclose(w, torch.from_numpy(g["inter_w"]), atol=1e-6)
    G = O.inter_group_feats(torch.from_numpy(g["idx"]), w, torch.from_numpy(g["feats"]))
    assert torch.allclose(G, torch.from_numpy(g["grouped"]), atol=1e-5)


def test_intra_conv_matches_reference(golden_dir):
    g = _load(golden_dir, "ref_intra_small.npz")
    out = O.basic_conv(torch.from_numpy(g["W"]),
                       O.intra_group_feats(torch.from_numpy(C.intra_idx()), torch.from_numpy(g["feats"])))
    assert torch.allclose(out, torch.from_numpy(g["out"]), atol=1e-5)


def test_blocks_forward_backward_match_reference(golden_dir):
    import sys
    sys.path.insert(0, golden_dir)
    from make_golden import small_params
    g = _load(golden_dir, "ref_blocks_small.npz")
    sd = {k[len("state/"):]: torch.from_numpy(g[k]).clone() for k in g.files if k.startswith("state/")}
    for k in sd:
        if k.endswith(("W", "weight", "bias")) and "running" not in k:
            sd[k].requires_grad_(True)
    pts = torch.from_numpy(g["in_points"])
    xyz = pts.permute(0, 2, 1).contiguous()
    feats = torch.ones(pts.shape[0], 1, pts.shape[1], 60)
    oxyz, of = O.backbone_forward(sd, small_params(), xyz, feats, torch.from_numpy(C.anchors_all()),
                                  torch.from_numpy(C.intra_idx()), C.kernel_points_base(), training=True)
    ref = torch.from_numpy(g["out_feats"])
    assert torch.equal(oxyz, torch.from_numpy(g["out_xyz"]))
    assert (of - ref).abs().max() / ref.abs().max() < 1e-5
    loss = of.square().mean()
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    loss.backward()
    for k in g.files:
        if k.startswith("grad/"):
            name = k[len("grad/"):]
            gr, rr = sd[name].grad, torch.from_numpy(g[k])
            # the first skip branch normalises a constant tensor: its true gradient is 0 and both sides hold rounding noise ~1e-6
            assert (gr - rr).abs().max() <= 2e-4 * rr.abs().max() + 5e-6, name
        if k.startswith("after/"):
            name = k[len("after/"):]
            assert torch.allclose(sd[name], torch.from_numpy(g[k]), atol=1e-5), name


def test_pose_grouping_matches_reference(golden_dir):
    """Non-identity per-point rotations: relative rotation of the neighbour offsets and the anchor permutation
    (functional.py:1061-1261) against the reference's own outputs, permute_modes 0 and 1."""
    g = _load(golden_dir, "ref_pose_group_small.npz")
    xyz, pose, feats = torch.from_numpy(g["xyz"]), torch.from_numpy(g["pose"]), torch.from_numpy(g["feats"])
    idx, _ = O.ball_query(xyz, xyz, float(g["radius"]), int(g["nn"]))
    anchors = torch.from_numpy(C.anchors_all())
    for pm in (0, 1):
        G, _, pi = O.pose_inter_group_feats(xyz, pose, feats, idx, anchors, torch.from_numpy(g["kernels"]), float(g["sigma"]), pm)
        ref = torch.from_numpy(g[f"grouped_pm{pm}"])
        assert (G - ref).abs().max() / ref.abs().max() < 1e-5, pm
    assert not torch.equal(pi, torch.arange(60).view(1, 1, 1, 60).expand_as(pi))     # the permutation is not trivial here
    # identity pose reduces to the plain grouping
    eye = torch.eye(4).repeat(1, xyz.shape[2], 1, 1)
    G0, gx, pi0 = O.pose_inter_group_feats(xyz, eye, feats, idx, anchors, torch.from_numpy(g["kernels"]), float(g["sigma"]), 1)
    w = O.anchor_weights(gx, anchors, torch.from_numpy(g["kernels"]), float(g["sigma"]))
    assert torch.equal(pi0, torch.arange(60).view(1, 1, 1, 60).expand_as(pi0))
    assert torch.allclose(G0, O.inter_group_feats(idx, w, feats), atol=1e-6)


def test_pose_grouping_strided_matches_reference(golden_dir):
    """Strided branch (functional.py:931-1029, stride 2): centres by FPS / lazy sampling, their rotations pose[sample_idx],
    neighbours from the full cloud -- against the reference's own outputs (sample indices, centre poses, grouped features)."""
    g = _load(golden_dir, "ref_pose_group_strided_small.npz")
    xyz, pose, feats = torch.from_numpy(g["xyz"]), torch.from_numpy(g["pose"]), torch.from_numpy(g["feats"])
    anchors, kern = torch.from_numpy(C.anchors_all()), torch.from_numpy(g["kernels"])
    for lazy in (0, 1):
        _, idx, sidx, sxyz = O.ball_grouping(xyz, int(g["stride"]), float(g["radius"]), int(g["nn"]), bool(lazy))
        assert torch.equal(sidx.long(), torch.from_numpy(g[f"sample_idx_lazy{lazy}"]).long())
        assert torch.equal(sxyz, torch.from_numpy(g[f"new_xyz_lazy{lazy}"]))
        sp = torch.gather(pose, 1, sidx.long().view(*sidx.shape, 1, 1).expand(-1, -1, 4, 4))
        assert torch.equal(sp, torch.from_numpy(g[f"sampled_pose_lazy{lazy}"]))
        for pm in ((0, 1) if lazy == 0 else (1,)):
            G, _, _ = O.pose_inter_group_feats(xyz, pose, feats, idx, anchors, kern, float(g["sigma"]), pm, sxyz, sidx)
            ref = torch.from_numpy(g[f"grouped_lazy{lazy}_pm{pm}"])
            assert G.shape == ref.shape
            assert (G - ref).abs().max() / ref.abs().max() < 1e-5, (lazy, pm)


def test_anchor_orbit_chamfer_oracle_vs_fp64_bruteforce():
    """oracle.so3.anchor_orbit_chamfer (the unfused reference path of model 38's reconstruction loss, built on the pinned
    chamfer restatement) against a float64 brute-force evaluation."""
    from oracle import so3 as O
    g = torch.Generator().manual_seed(0)
    b, a, m, n = 2, 5, 40, 60
    q, _ = torch.linalg.qr(torch.randn(b, a, 3, 3, generator=g))
    canon, ori = torch.randn(b, 3, m, generator=g), torch.randn(b, 3, n, generator=g)
    tr = torch.randn(b, a, 3, generator=g) * 0.1
    r = O.anchor_orbit_chamfer(canon, q.contiguous(), tr, ori)
    y = torch.einsum('bajk,bkm->bamj', q.double(), canon.double()) + tr.double().unsqueeze(2)
    D = ((y[:, :, :, None] - ori.double().transpose(1, 2)[:, None, None]) ** 2).sum(-1)
    assert float((D.min(3)[0] - r['d1']).abs().max()) < 1e-5 and float((D.min(2)[0] - r['d2']).abs().max()) < 1e-5
    assert torch.equal(r['i1'].long(), D.min(3)[1]) and torch.equal(r['i2'].long(), D.min(2)[1])
    total = D.min(3)[0].mean(-1) + D.min(2)[0].mean(-1)
    assert torch.equal(r['orbit'], total.argmin(-1))
    single = O.anchor_orbit_chamfer(canon, q.contiguous(), tr, ori, glb_single_cd=1)
    assert torch.equal(single['orbit'], D.min(2)[0].mean(-1).argmin(-1))


def test_pointnet_head_oracle_vs_reference_fixture(golden_dir):
    """oracle.so3.pointnet_so3conv against the reference's own PointnetSO3Conv (tests/golden/make_golden.py pointnet)."""
    g = np.load(os.path.join(golden_dir, "ref_pointnet_small.npz"))
    anchors = torch.from_numpy(np.ascontiguousarray(C.get_anchors(60)))
    for na in (60, 1):
        t = {k[len(f'a{na}_'):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(f'a{na}_')}
        feats = t['feats'].clone().requires_grad_(True)
        w, b = t['weight'].clone().requires_grad_(True), t['bias'].clone().requires_grad_(True)
        anc = anchors if na == 60 else anchors[29:30]
        pooled = O.pointnet_so3conv(w, b, anc, t['xyz'], feats)
        raw = O.pointnet_so3conv(w, b, anc, t['xyz'], feats, return_raw=True)
        assert torch.allclose(pooled, t['pooled'], atol=2e-6) and torch.allclose(raw, t['raw'], atol=2e-6)
        (pooled * t['grad_out']).sum().backward()
        assert torch.allclose(feats.grad, t['grad_feats'], atol=1e-5)
        assert torch.allclose(w.grad.view_as(t['grad_weight']), t['grad_weight'], atol=1e-5)
        assert torch.allclose(b.grad, t['grad_bias'], atol=1e-5)


# ---------------------------------------------------------------- PointNet++ encoder-decoder
def test_pointnet2_oracle_vs_reference_fixture(golden_dir):
    """oracle.pointnet2.forward against the reference's own PointnetPP (tests/golden/make_golden.py pointnet2):
    sampled positions bit-exact, sorted neighbour distances to one ulp, features to fp32 rounding, running statistics."""
    from oracle import pointnet2 as OP
    g = np.load(os.path.join(golden_dir, "ref_pointnet2_small.npz"))
    sd = OP.make_state(6, seed=int(g["seed"]))
    taps, stats = {}, {}
    out, glb, pos_out = OP.forward(sd, torch.from_numpy(g["x"]), torch.from_numpy(g["pos"]), taps=taps, stats=stats)
    assert np.array_equal(taps["pos0"].numpy(), g["pos_512"]) and np.array_equal(taps["pos1"].numpy(), g["pos_128"])
    # torch.sqrt is correctly rounded on some hosts and 1 ulp off for ~0.5 % of the values on AVX-512 ones
    assert np.abs(taps["topk_dist0"].numpy() - g["topk_dist_512"]).max() <= 6e-8
    assert np.abs(taps["topk_dist1"].numpy() - g["topk_dist_128"]).max() <= 6e-8
    assert np.array_equal(pos_out.numpy(), g["pos_out"])
    assert torch.allclose(glb, torch.from_numpy(g["global_x"]), atol=1e-5, rtol=1e-5)
    assert torch.allclose(out, torch.from_numpy(g["out"]), atol=1e-5, rtol=1e-5)
    for name in ("mlp_layers.0.2.1", "mlp_layers.2.2.1", "up_mlp_layers.2.2.1"):
        rm, rv = stats[name[:-2]]
        assert np.allclose(rm.numpy(), g["rm:" + name], atol=1e-6) and np.allclose(rv.numpy(), g["rv:" + name], atol=1e-6)
    # the reference's own backward cannot run (in-place mask on the ReLU output, PointNet2.py:109): recorded, not assumed
    assert str(g["reference_backward"]).startswith("RuntimeError")


def test_pointnet2_oracle_gradient_defined():
    """The restatement masks out of place, so its backward exists; masked / non-maximal rows get no gradient."""
    from oracle import pointnet2 as OP
    y = torch.randn(2, 5, 8, 4).relu().requires_grad_(True)
    d = torch.rand(2, 5, 8)
    d[..., 0] = 0.0
    r = OP.max_pooling_with_r(y, d, 0.5)
    r.sum().backward()
    assert float(y.grad[(d > 0.5)].abs().max()) == 0.0
    assert torch.equal(y.grad.sum(2), torch.ones(2, 5, 4))


def test_pointnet2_c_oracle_matches_torch_restatement_and_fixture(golden_dir):
    """oracle_ops.c::oracle_knn / oracle_three_nn (plain C, explicit evaluation order, correctly rounded sqrt) against the
    torch evaluation the reference performs: same neighbours, squared-distance order bit-exact, distances equal to
    numpy's correctly rounded sqrt and within one ulp of torch.sqrt; and against the neighbour distances the reference's
    own module produced (fixture)."""
    g = torch.Generator().manual_seed(3)
    pos = torch.rand(3, 500, 3, generator=g) - 0.5
    cen = pos[:, torch.randperm(500, generator=g)[:40]].contiguous()
    idx, dist = cops.knn(pos.numpy(), cen.numpy(), 64)
    d2 = torch.sum((cen.unsqueeze(2) - pos.unsqueeze(1)) ** 2, dim=-1)                      # PointNet2.py:85
    ref_d2, ref_i = torch.topk(d2, k=64, dim=2, largest=False)
    assert np.array_equal(dist, np.sqrt(ref_d2.numpy()))
    assert np.abs(dist - torch.sqrt(ref_d2).numpy()).max() <= 6e-8
    uniq = np.ones(ref_d2.shape, bool)
    same = (ref_d2[..., 1:] == ref_d2[..., :-1]).numpy()
    uniq[..., 1:] &= ~same
    uniq[..., :-1] &= ~same
    uniq[..., -1] = False
    assert np.array_equal(idx[uniq], ref_i.numpy()[uniq])
    # ties resolve to the smaller index
    grid = (torch.randint(-3, 4, (1, 200, 3), generator=g).float() * 0.25).contiguous()
    gi, _ = cops.knn(grid.numpy(), grid[:, :8].contiguous().numpy(), 32)
    full = torch.sum((grid[:, :8].unsqueeze(2) - grid.unsqueeze(1)) ** 2, dim=-1).double()
    assert np.array_equal(gi, (full * 1e6 + torch.arange(200).double() * 1e-3).argsort(dim=2)[..., :32].numpy())
    # the reference module's own neighbour distances (level 1 of the fixture), to one ulp of its torch.sqrt
    f = np.load(os.path.join(golden_dir, "ref_pointnet2_small.npz"))
    _, fd = cops.knn(f["pos"], f["pos_512"], 64)
    assert np.abs(fd - f["topk_dist_512"]).max() <= 6e-8
    # three nearest neighbours + inverse-distance weights (interpolate_features :114-123)
    p1, p2 = torch.rand(2, 90, 3, generator=g) - 0.5, torch.rand(2, 300, 3, generator=g) - 0.5
    ti, tw = cops.three_nn(p1.numpy(), p2.numpy())
    dist3 = torch.norm(p2[:, :, None, :] - p1[:, None, :, :], dim=-1, p=2)
    rd, ri = dist3.topk(3, dim=-1, largest=False)
    assert np.array_equal(ti, ri.numpy())
    rec = 1.0 / (rd + 1e-8)
    assert np.abs(tw - (rec / rec.sum(2, keepdim=True)).numpy()).max() <= 2e-7          # torch.norm's sqrt: one ulp
    one_i, one_w = cops.three_nn(p1[:, :1].contiguous().numpy(), p2.numpy())
    assert (one_i == 0).all() and np.array_equal(one_w[..., 0], np.ones((2, 300), np.float32)) and (one_w[..., 1:] == 0).all()
