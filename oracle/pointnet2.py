"""CPU restatement (torch fp32) of the reference's PointNet++ encoder-decoder `PointnetPP`
(SPConvNets/models/PointNet2.py:8-196; helpers SPConvNets/models/model_util.py:93-118,148-156,183-200).

TEST INFRASTRUCTURE ONLY: imported by tests/, tests/golden/make_golden.py and nothing else.  Pinned on
tests/golden/ref_pointnet2_small.npz, the outputs of the reference's own module run on the CPU
(tests/test_oracle_golden.py::test_pointnet2_*).

Third-party arithmetic: the reference samples with torch_cluster.fps(random_start=False)
(torch-cluster==1.5.9, env.yaml:13; not vendored, no reference test pins it -> "parity unpinned" for the sampler
itself).  Its published algorithm is plain farthest-point sampling started at the first point of every batch
segment; oracle_ops.c::oracle_fps_plain restates it and both the fixture and this file use that.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import cops

N_SAMPLES = [512, 128, 1]                                   # PointNet2.py:19
MLPS = [[64, 64, 128], [128, 128, 256], [256, 512, 1024]]   # :22
UP_MLPS = [[256, 256], [256, 128], [128, 128, 128]]         # :25
RADIUS = [0.2, 0.4, None]                                   # :30


def layer_dims(in_feat_dim, n_layers=3):
    """-> (encoder [(n_in, dims)], decoder [(n_in, dims)]) exactly as PointnetPP.__init__ builds them (:22-40)."""
    mlps_in = [in_feat_dim, 128 + 3, 256 + 3]
    up_in = [1024 + 256, 256 + 128, 128 + in_feat_dim]
    enc = list(zip(mlps_in[:n_layers], MLPS[:n_layers]))
    dec = list(zip(up_in[-n_layers:], UP_MLPS[-n_layers:]))
    return enc, dec


def make_state(in_feat_dim, seed=0, n_layers=3):
    """A deterministic state dict with the reference's keys (mlp_layers.L.J.0.weight [Co,Ci,1,1], .0.bias,
    .1.weight/.bias/.running_mean/.running_var/.num_batches_tracked), drawn from numpy's RandomState so the
    fixture generator and the tests rebuild the same weights without storing them."""
    rs = np.random.RandomState(seed)
    enc, dec = layer_dims(in_feat_dim, n_layers)
    sd = {}
    for name, spec in (("mlp_layers", enc), ("up_mlp_layers", dec)):
        for li, (n_in, dims) in enumerate(spec):
            ci = n_in
            for j, co in enumerate(dims):
                p = f"{name}.{li}.{j}"
                sd[p + ".0.weight"] = torch.from_numpy((rs.standard_normal((co, ci, 1, 1)) * np.sqrt(2.0 / ci)).astype(np.float32))
                sd[p + ".0.bias"] = torch.from_numpy((rs.standard_normal(co) * 0.1).astype(np.float32))
                sd[p + ".1.weight"] = torch.from_numpy((1.0 + 0.2 * rs.standard_normal(co)).astype(np.float32))
                sd[p + ".1.bias"] = torch.from_numpy((0.1 * rs.standard_normal(co)).astype(np.float32))
                sd[p + ".1.running_mean"] = torch.zeros(co)
                sd[p + ".1.running_var"] = torch.ones(co)
                sd[p + ".1.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
                ci = co
    return sd


def farthest_point_sampling(pos, n_sampling):
    """model_util.py:183-200 over torch_cluster.fps: pos [B,N,3] -> flat global indices [B*n_sampling]."""
    b, n, _ = pos.shape
    xyz = pos[:, :, :3].float().permute(0, 2, 1).contiguous().numpy()
    idx = torch.from_numpy(cops.fps_plain(xyz, n_sampling)).long()
    return (idx + torch.arange(b).view(b, 1) * n).reshape(-1)


def sample_and_group(feat, pos, n_samples, k=64):
    """PointNet2.py:78-100 (use_pos=True)."""
    b, n = pos.shape[:2]
    fps_idx = farthest_point_sampling(pos, n_samples)
    sampled = pos.reshape(b * n, -1)[fps_idx].view(b, n_samples, -1)
    ppdist = torch.sqrt(torch.sum((sampled.unsqueeze(2) - pos.unsqueeze(1)) ** 2, dim=-1))
    topk_dist, topk_idx = torch.topk(ppdist, k=k, dim=2, largest=False)
    gather = lambda v: torch.gather(v.unsqueeze(1).expand(-1, n_samples, -1, -1), 2,
                                    topk_idx.unsqueeze(-1).expand(-1, -1, -1, v.shape[-1]))
    grouped = gather(pos) - sampled.unsqueeze(2)
    if feat is not None:
        grouped = torch.cat([grouped, gather(feat)], dim=-1)
    return grouped, topk_dist, topk_idx, sampled


def max_pooling_with_r(grouped_feat, ppdist, r=None):
    """PointNet2.py:102-112.  The reference overwrites the masked entries of the MLP output in place (which its own
    autograd then rejects in backward); the values are the same with torch.where, and the gradient is the natural one."""
    if r is not None:
        grouped_feat = torch.where((ppdist <= r).unsqueeze(-1), grouped_feat, torch.full_like(grouped_feat, -1e8))
    return torch.max(grouped_feat, dim=2)[0]


def interpolate_features(feat, p1, p2):
    """PointNet2.py:114-129."""
    dist = torch.norm(p2[:, :, None, :] - p1[:, None, :, :], dim=-1, p=2)
    kk = min(3, dist.size(-1))
    dist, idx = dist.topk(kk, dim=-1, largest=False)
    rec = 1.0 / (dist + 1e-8)
    w = rec / torch.sum(rec, dim=2, keepdim=True)
    near = torch.gather(feat.unsqueeze(1).expand(-1, p2.shape[1], -1, -1), 2, idx.unsqueeze(-1).expand(-1, -1, -1, feat.shape[-1]))
    return torch.sum(near * w[:, :, :, None], dim=2)


def apply_mlp(x, sd, prefix, n_blocks, training=True, stats=None):
    """model_util.py:148-156 over the blocks of construct_conv_modules (:93-118): x [B,S,k,C] -> [B,S,k,C'];
    each block = Conv2d 1x1 (bias) + BatchNorm2d (batch statistics in training mode) + ReLU."""
    x = x.permute(0, 3, 1, 2).contiguous()
    for j in range(n_blocks):
        p = f"{prefix}.{j}"
        x = F.conv2d(x, sd[p + ".0.weight"], sd[p + ".0.bias"])
        rm, rv = sd[p + ".1.running_mean"].clone(), sd[p + ".1.running_var"].clone()
        x = F.batch_norm(x, rm, rv, sd[p + ".1.weight"], sd[p + ".1.bias"], training, 0.1, 1e-5)
        if stats is not None:
            stats[p] = (rm, rv)
        x = F.relu(x)
    return x.permute(0, 2, 3, 1)


def forward(sd, x, pos, n_layers=3, training=True, stats=None, taps=None):
    """PointnetPP.forward (:131-196) with return_global=True -> (x, global_x, pos)."""
    b = pos.size(0)
    cache = [(x, pos)]
    for i, ns in enumerate(N_SAMPLES[:n_layers]):
        nb = len(MLPS[i])
        if ns == 1:
            g = torch.cat([pos.unsqueeze(1), x.unsqueeze(1)], dim=-1)
            g = apply_mlp(g, sd, f"mlp_layers.{i}", nb, training, stats).squeeze(1)
            x = torch.max(g, dim=1, keepdim=True)[0]
            pos = torch.zeros((b, 1, 3), dtype=torch.float32)
        else:
            g, topk_dist, topk_idx, pos = sample_and_group(x, pos, ns, k=64)
            if taps is not None:
                taps[f"topk_dist{i}"], taps[f"topk_idx{i}"], taps[f"pos{i}"] = topk_dist, topk_idx, pos
            g = apply_mlp(g, sd, f"mlp_layers.{i}", nb, training, stats)
            x = max_pooling_with_r(g, topk_dist, RADIUS[i])
        if taps is not None:
            taps[f"x{i}"] = x
        cache.append((x, pos))
    global_x = x
    n_up = n_layers
    for i in range(n_up):
        prev_x, prev_pos = cache[-i - 2]
        interp = interpolate_features(x, pos, prev_pos)
        if prev_x is None:
            prev_x = prev_pos
        elif i == n_up - 1:
            prev_x = torch.cat([prev_x, prev_pos], dim=-1)
        cur = torch.cat([interp, prev_x], dim=-1)
        x = apply_mlp(cur.unsqueeze(2), sd, f"up_mlp_layers.{i}", len(UP_MLPS[-n_layers:][i]), training, stats).squeeze(2)
        pos = prev_pos
    return x, global_x, pos
