"""Shared parity helpers for the GPU tests, __graft_entry__.smoke() and bench.py's checker leg.
Everything here compares the CUDA path (through the C ABI) with the oracle / golden fixtures."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def rel_err(x, ref):
    """max|x-ref| / max|ref|  -- the per-tensor metric of BASELINE.md (fp32 bar: 1e-4)."""
    ref = ref.detach().float().cpu()
    x = x.detach().float().cpu()
    return float((x - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def small_params():
    sys.path.insert(0, GOLD)
    from make_golden import small_params as sp
    return sp()


def build_backbone(params, state=None, device="cuda"):
    """Our fused block stack with the reference's module/parameter names."""
    from equi_articulated_pose_b200 import blocks
    net = blocks.SO3Backbone(params)
    if state is not None:
        missing, unexpected = net.load_state_dict(state, strict=False)
        assert not unexpected, unexpected
        bad = [k for k in missing if not k.endswith(("anchors", "kernels", "intra_idx", "num_batches_tracked",
                                                        "running_mean", "running_var"))]
        assert not bad, bad
    return net.to(device)


def run_blocks_case(device, seed=2001):
    """The committed reference fixture (tests/golden/ref_blocks_small.npz: outputs of the
    reference's own Python, fwd + bwd, train mode) replayed on the CUDA path."""
    g = np.load(os.path.join(GOLD, "ref_blocks_small.npz"))
    sd = {k[len("state/"):]: torch.from_numpy(g[k]).clone() for k in g.files if k.startswith("state/")}
    net = build_backbone(small_params(), sd, device)
    net.train()
    pts = torch.from_numpy(g["in_points"]).to(device)
    out = net(pts)
    feats = out.feats
    loss = feats.square().mean()
    loss.backward()
    err = {"out": rel_err(feats, torch.from_numpy(g["out_feats"])),
           "xyz": float((out.xyz.cpu() - torch.from_numpy(g["out_xyz"])).abs().max()),
           "loss": abs(float(loss.detach()) - float(g["loss"])) / abs(float(g["loss"]))}
    worst = 0.0
    named = dict(net.named_parameters())
    for k in g.files:
        if k.startswith("grad/"):
            ref = torch.from_numpy(g[k])
            got = named[k[len("grad/"):]].grad
            # the first skip branch normalises a constant tensor: its true gradient is 0, both sides hold noise
            d, scale = float((got.cpu() - ref).abs().max()), float(ref.abs().max())
            e = d / scale if scale > 1e-4 else (0.0 if d < 2e-5 else 1.0)
            worst = max(worst, e)
    err["grad"] = worst
    stats = 0.0
    bufs = dict(net.named_buffers())
    for k in g.files:
        if k.startswith("after/"):
            stats = max(stats, float((bufs[k[len("after/"):]].cpu() - torch.from_numpy(g[k])).abs().max()))
    err["running_stats"] = stats
    return err


def check_index_ops(device, b, n, m, nsample, radius, seed):
    """ball query / FPS / gather bit-exact against oracle_ops.c on one seeded cloud."""
    from oracle import cops, so3 as O
    from equi_articulated_pose_b200 import ops
    pts = O.synthetic_cloud(b, n, seed)
    xyz = pts.permute(0, 2, 1).contiguous()
    fps_ref = cops.furthest_point_sampling(xyz.numpy(), m)
    fps = ops.furthest_point_sampling(xyz.to(device), m)
    assert np.array_equal(fps.cpu().numpy(), fps_ref), "FPS indices differ from the oracle"
    q = ops.gather_points_forward(xyz.to(device), fps)
    q_ref = cops.gather_points_forward(xyz.numpy(), fps_ref)
    assert np.array_equal(q.cpu().numpy(), q_ref), "gather differs from the oracle"
    bq_ref = cops.ball_query(q_ref, xyz.numpy(), radius, nsample)
    bq = ops.ball_query(q, xyz.to(device), radius, nsample)
    assert np.array_equal(bq.cpu().numpy(), bq_ref), "ball-query indices differ from the oracle"
