"""torch-fp32 CPU restatement of the reference's SO(3) point-convolution math.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity is pinned: tests/golden/ holds
outputs of the reference's own Python (imported in place through oracle/ref_harness.py by
tests/golden/make_golden.py) and tests/test_oracle_golden.py checks this file against them.

Citations are relative to /root/reference.  Layouts are the reference's:
  xyz [B,3,N]   feats [B,C,N,A]   anchors [A,3,3]   kernels [K,3]   inter_idx [B,P,nn] int
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import cops

KERNEL_CONDENSE_RATIO = 0.7  # vgtk/vgtk/so3conv/modules.py:16


# ----------------------------------------------------------------------------- constants
def scaled_kernel_points(base_points, radius):
    """vgtk/vgtk/so3conv/functional.py:111-121 with radius := 0.7*radius (modules.py:132).
    `base_points` are the 24 kpsphere points (float32 [K,3]); scaled so the largest norm is
    `KERNEL_CONDENSE_RATIO*radius`."""
    pc = np.asarray(base_points, np.float32)
    radius = KERNEL_CONDENSE_RATIO * radius
    r = np.sqrt((pc ** 2).sum(1).max())
    return pc * radius / r


# ----------------------------------------------------------------------------- index ops
def ball_query(query_xyz, support_xyz, radius, n_sample):
    """vgtk/vgtk/spconv/functional.py:341-350 + pc/sample.py:54-59 -> idx [B,P,nn] (int32)
    and grouped support coordinates [B,3,P,nn] (the appended shadow point is never indexed)."""
    idx = torch.from_numpy(cops.ball_query(query_xyz.detach().float().numpy(), support_xyz.detach().float().numpy(),
                                           radius, n_sample))
    b, p, nn = idx.shape
    g = torch.gather(support_xyz, 2, idx.view(b, 1, -1).expand(-1, 3, -1).long()).view(b, 3, p, nn)
    return idx, g


def furthest_sample_index(xyz, n_sample, lazy_sample):
    """vgtk/vgtk/pc/sample.py:63-72."""
    if xyz.shape[2] == n_sample or lazy_sample:
        return torch.arange(n_sample, dtype=torch.int32).view(1, -1).expand(xyz.shape[0], -1).contiguous()
    return torch.from_numpy(cops.furthest_point_sampling(xyz.detach().float().numpy(), n_sample))


def ball_grouping(xyz, stride, radius, n_neighbor, lazy_sample=True):
    """vgtk/vgtk/spconv/functional.py:428-449 -> grouped_xyz [B,3,P,nn] (centre-relative),
    ball_idx [B,P,nn], sample_idx [B,P], sample_xyz [B,3,P]."""
    n_sample = math.ceil(xyz.shape[2] / stride)
    if stride > 1:
        sidx = furthest_sample_index(xyz, n_sample, lazy_sample)
        sample_xyz = torch.gather(xyz, 2, sidx.long().unsqueeze(1).expand(-1, 3, -1))
    else:
        sample_xyz = xyz
        sidx = torch.arange(xyz.shape[2], dtype=torch.long, device=xyz.device).unsqueeze(0).repeat(xyz.shape[0], 1)
    ball_idx, grouped = ball_query(sample_xyz, xyz, radius, n_neighbor)
    return grouped - sample_xyz.unsqueeze(3), ball_idx, sidx, sample_xyz


# ----------------------------------------------------------------------------- float ops
def anchor_weights(grouped_xyz, anchors, kernels, sigma):
    """Kernel-point correlation, vgtk/vgtk/so3conv/functional.py:2508-2549:
    w[b,p,a,k,n] = relu(1 - |g[b,:,p,n] - R_a kappa_k|^2 / sigma)."""
    rk = torch.matmul(anchors, kernels.t())              # [A,3,K] = R_a kappa_k
    rk = rk.permute(1, 0, 2)                             # [3,A,K]
    diff = grouped_xyz[:, :, :, None, None, :] - rk[None, :, None, :, :, None]
    d2 = (diff ** 2).sum(1)                              # [B,P,A,K,nn]
    return F.relu(1.0 - d2 / sigma)


def inter_group_feats(inter_idx, inter_w, feats):
    """vgtk/vgtk/spconv/functional.py:375-406: G[b,c,k,p,a] = sum_n f[b,c,idx[b,p,n],a] w[b,p,a,k,n]."""
    b, p, nn = inter_idx.shape
    _, c, n, a = feats.shape
    flat = inter_idx.long().reshape(b, 1, p * nn, 1).expand(-1, c, -1, a)
    nb = torch.gather(feats, 2, flat).view(b, c, p, nn, a)
    return torch.einsum('bcpna,bpakn->bckpa', nb, inter_w).contiguous()


def intra_group_feats(intra_idx, feats):
    """vgtk/vgtk/so3conv/functional.py:2553-2567: G[b,c,k,p,a] = f[b,c,p,intra_idx[a,k]]."""
    b, c, p, a = feats.shape
    k = intra_idx.shape[1]
    return feats.index_select(3, intra_idx.reshape(-1).long()).view(b, c, p, a, k).permute(0, 1, 4, 2, 3).contiguous()


def pose_inter_group_feats(xyz, pose, feats, idx, anchors, kernels, sigma, permute_modes=0, sample_xyz=None, sample_idx=None):
    """Pose-aware inter grouping, both branches of inter_so3poseconv_grouping_strided
    (vgtk/vgtk/so3conv/functional.py:896-1060 strided, :1061-1261 no stride).  xyz [B,3,N], pose [B,N,4,4],
    feats [B,C,N,A], idx [B,P,nn] -> (G [B,C,K,P,A], grouped_xyz [B,3,P,nn], rotated_anchor_idx [B,P,nn,A]).
    Strided: the P centres are sample_xyz [B,3,P] = xyz[sample_idx] with rotations pose[sample_idx]
    (vgtk/vgtk/spconv/functional.py:468-500); default: every point is a centre.
      R_rel[p,n] = R_p R_{j(p,n)}^T ;  g'[p,n] = R_rel (x_j - x_p) ;  w = relu(1 - |g' - R_a kappa_k|^2 / sigma)
      pi[p,n,a] = argmax_{a'} tr((R_rel^T R_a) R_{a'}^T)      (only applied when permute_modes != 0)
      G[b,c,k,p,a] = sum_n w[b,p,a,k,n] feats[b,c,j(p,n), pi[p,n,a] or a]"""
    b, _, n = xyz.shape
    p, nn_ = idx.shape[1], idx.shape[2]
    a = anchors.shape[0]
    R = pose[:, :, :3, :3]
    li = idx.long()
    if sample_idx is None:
        Rp, cxyz = R, xyz
    else:
        Rp = torch.gather(R, 1, sample_idx.long().view(b, p, 1, 1).expand(-1, -1, 3, 3))
        cxyz = sample_xyz
    Rj = torch.gather(R, 1, li.reshape(b, p * nn_, 1, 1).expand(-1, -1, 3, 3)).view(b, p, nn_, 3, 3)
    rel = torch.matmul(Rp.unsqueeze(2), Rj.transpose(3, 4))                              # [B,P,nn,3,3]
    g = torch.gather(xyz, 2, li.reshape(b, 1, -1).expand(-1, 3, -1)).view(b, 3, p, nn_) - cxyz.unsqueeze(-1)
    g = torch.matmul(rel, g.permute(0, 2, 3, 1).unsqueeze(-1)).squeeze(-1).permute(0, 3, 1, 2).contiguous()
    w = anchor_weights(g, anchors, kernels, sigma)
    rot_anchors = torch.matmul(rel.transpose(-1, -2).unsqueeze(3), anchors)              # [B,P,nn,A,3,3]
    tr = torch.einsum('bpnaij,cij->bpnac', rot_anchors, anchors)                         # tr(M R_c^T) = <M, R_c>
    pi = tr.argmax(-1)                                                                   # [B,P,nn,A]
    c = feats.shape[1]
    f = feats.permute(0, 2, 3, 1)                                                        # [B,N,A,C]
    gf = torch.gather(f, 1, li.reshape(b, p * nn_, 1, 1).expand(-1, -1, a, c)).view(b, p, nn_, a, c)
    if permute_modes != 0:
        gf = torch.gather(gf, 3, pi.unsqueeze(-1).expand(-1, -1, -1, -1, c))
    G = torch.einsum('bpnac,bpakn->bckpa', gf, w).contiguous()
    return G, g, pi


def basic_conv(W, G):
    """vgtk/vgtk/so3conv/modules.py:48-55: out[b,o,p,a] = sum_{c,k} W[o, c*K+k] G[b,c,k,p,a]."""
    b, c, k, p, a = G.shape
    return torch.matmul(W, G.reshape(b, c * k, p * a)).view(b, W.shape[0], p, a)


# ----------------------------------------------------------------------------- blocks
def _bn(x, sd, prefix, training, momentum=0.1, eps=1e-5):
    return F.batch_norm(x, sd.get(prefix + 'running_mean'), sd.get(prefix + 'running_var'),
                        sd[prefix + 'weight'], sd[prefix + 'bias'], training, momentum, eps)


def inter_block(sd, prefix, args, xyz, feats, anchors, base_kp, training=True):
    """SPConvNets/utils/base_so3conv.py:93-132 (InterSO3ConvBlock with norm=BatchNorm2d,
    activation=leaky_relu) around vgtk/vgtk/so3conv/modules.py:125-174."""
    kernels = torch.from_numpy(scaled_kernel_points(base_kp, args['radius'])).to(feats.dtype)
    gxyz, idx, sidx, new_xyz = ball_grouping(xyz, args['stride'], args['radius'], args['n_neighbor'],
                                             args.get('lazy_sample', True))
    w = anchor_weights(gxyz, anchors, kernels, args['sigma'])
    G = inter_group_feats(idx, w, feats)
    y = basic_conv(sd[prefix + 'conv.basic_conv.W'], G)
    y = F.leaky_relu(_bn(y, sd, prefix + 'norm.', training))
    return idx, w, sidx, new_xyz, y


def intra_block(sd, prefix, feats, intra_idx):
    """SPConvNets/utils/base_so3conv.py:37-67 (IntraSO3ConvBlock: InstanceNorm2d(affine=False),
    leaky_relu) around vgtk/vgtk/so3conv/modules.py:325-347."""
    y = basic_conv(sd[prefix + 'conv.basic_conv.W'], intra_group_feats(intra_idx, feats))
    return F.leaky_relu(F.instance_norm(y, eps=1e-5))


def separable_block(sd, prefix, args, xyz, feats, anchors, intra_idx, base_kp, training=True):
    """SPConvNets/utils/base_so3conv.py:174-218 (SeparableSO3ConvBlock)."""
    skip = feats
    idx, w, sidx, new_xyz, y = inter_block(sd, prefix + 'inter_conv.', args, xyz, feats, anchors, base_kp, training)
    if args['kanchor'] > 1:
        y = intra_block(sd, prefix + 'intra_conv.', y, intra_idx)
    if args['stride'] > 1:
        b, c, _, a = skip.shape
        skip = torch.gather(skip, 2, sidx.long().view(b, 1, -1, 1).expand(-1, c, -1, a))
    skip = F.conv2d(skip, sd[prefix + 'skip_conv.weight'], sd[prefix + 'skip_conv.bias'])
    skip = F.leaky_relu(_bn(skip, sd, prefix + 'norm.', training))
    return new_xyz, y + skip


def backbone_params(input_num=1024, kanchor=60, mlps=((64, 64), (128, 128), (256, 256), (256,)),
                    strides=(2, 2, 2, 2), initial_radius_ratio=0.2, sampling_ratio=0.4,
                    sampling_density=0.5, input_radius=1.0, sigma_ratio=0.5, dropout_rate=0.0):
    """The per-layer argument dicts of the classic backbone: SPConvNets/models/cls_so3net_pn.py:43-150."""
    strides = list(strides)
    if input_num > 1024:
        sampling_ratio /= (input_num / 1024)
        strides[0] = int(2 * (input_num / 1024))
    mult = [2 ** i for i in range(len(mlps) + 1)]
    num_centers = [int(input_num / m) for m in mult]
    radius_ratio = [initial_radius_ratio * m ** sampling_density for m in mult]
    radii = [r * input_radius for r in radius_ratio]
    sig = [sigma_ratio * radii[0] ** 2]
    for i in range(len(strides)):
        sig.append(sig[i] * 2)
    blocks, dim_in = [], 1
    for i, block in enumerate(mlps):
        layers = []
        for j, dim_out in enumerate(block):
            neighbor = int(sampling_ratio * num_centers[i] * radius_ratio[i] ** (1 / sampling_density))
            if j == 0:
                stride, nidx = strides[i], (i if i == 0 else i + 1)
                neighbor *= 2
            else:
                stride, nidx = 1, i + 1
            layers.append({'type': 'separable_block' if kanchor >= 60 else 'inter_block', 'args': {
                'dim_in': dim_in, 'dim_out': dim_out, 'kernel_size': 1, 'stride': stride,
                'radius': radii[nidx], 'sigma': sig[nidx], 'n_neighbor': neighbor,
                'lazy_sample': (i != 0 or j != 0), 'dropout_rate': dropout_rate, 'multiplier': 2,
                'activation': 'leaky_relu', 'pooling': None, 'kanchor': kanchor, 'norm': 'BatchNorm2d'}})
            dim_in = dim_out
        blocks.append(layers)
    return blocks


def backbone_forward(sd, params, xyz, feats, anchors, intra_idx, base_kp, training=True, prefix='backbone.'):
    """ClsSO3ConvModel.forward's backbone loop (cls_so3net_pn.py:32-33) over BasicSO3ConvBlock
    (base_so3conv.py:135-169).  Returns (xyz, feats)."""
    for bi, block in enumerate(params):
        for li, layer in enumerate(block):
            pre = f'{prefix}{bi}.blocks.{li}.'
            if layer['type'] == 'separable_block':
                xyz, feats = separable_block(sd, pre, layer['args'], xyz, feats, anchors, intra_idx, base_kp, training)
            elif layer['type'] == 'inter_block':
                _, _, _, xyz, feats = inter_block(sd, pre, layer['args'], xyz, feats, anchors, base_kp, training)
            else:
                raise ValueError(layer['type'])
    return xyz, feats


def init_backbone_state(params, seed=0, n_intra=12, n_kernel=24):
    """Random-init state dict with the reference's key names/shapes (xavier-normal W with relu
    gain: so3conv/modules.py:35-41; Conv2d / BatchNorm2d torch defaults)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def xavier(co, ci, k):
        std = math.sqrt(2.0) * math.sqrt(2.0 / ((ci + co) * k))
        return (torch.randn(co, ci, k, generator=g) * std).view(co, ci * k)

    for bi, block in enumerate(params):
        for li, layer in enumerate(block):
            a = layer['args']
            ci, co = a['dim_in'], a['dim_out']
            pre = f'backbone.{bi}.blocks.{li}.'
            sep = layer['type'] == 'separable_block'
            ip = pre + ('inter_conv.' if sep else '')
            sd[ip + 'conv.basic_conv.W'] = xavier(co, ci, n_kernel)
            sd[ip + 'norm.weight'] = 1 + 0.1 * torch.randn(co, generator=g)
            sd[ip + 'norm.bias'] = 0.1 * torch.randn(co, generator=g)
            if sep:
                sd[pre + 'intra_conv.conv.basic_conv.W'] = xavier(co, co, n_intra)
                bound = 1 / math.sqrt(ci)
                sd[pre + 'skip_conv.weight'] = (torch.rand(co, ci, 1, 1, generator=g) * 2 - 1) * bound
                sd[pre + 'skip_conv.bias'] = (torch.rand(co, generator=g) * 2 - 1) * bound
                sd[pre + 'norm.weight'] = 1 + 0.1 * torch.randn(co, generator=g)
                sd[pre + 'norm.bias'] = 0.1 * torch.randn(co, generator=g)
    return sd


def synthetic_cloud(batch, n, seed):
    """Config-2 'sphere-shell' clouds: unit directions x U(0.85, 1)."""
    g = torch.Generator().manual_seed(seed)
    d = torch.randn(batch, n, 3, generator=g)
    d = d / d.norm(dim=2, keepdim=True)
    r = 0.85 + 0.15 * torch.rand(batch, n, 1, generator=g)
    return (d * r).float()


def anchor_orbit_chamfer(canon, rot, trans, ori, glb_single_cd=0):
    """Model 38's anchor-orbit reconstruction loss, the reference way
    (SPConvNets/models/unsup_seg_so3_pose_conv_pn_38_multi_stage.py:429-450): transform the reconstruction by every
    anchor pose, replicate the input cloud A times, chamfer on [B*A, ., 3], mean over points, min over anchors.

    canon [B,3,M], rot [B,A,3,3], trans [B,A,3], ori [B,3,N]  (the reference's layouts)
    -> dict(d1 [B,A,M], d2 [B,A,N], i1, i2, cd_r2o [B,A], cd_o2r [B,A], minn [B], orbit [B])"""
    bz, na = rot.shape[0], rot.shape[1]
    m, n = canon.shape[2], ori.shape[2]
    transformed = torch.matmul(rot, canon.unsqueeze(1)).transpose(-1, -2) + trans.unsqueeze(-2)       # :429-430
    expanded = ori.transpose(-1, -2).unsqueeze(1).contiguous().repeat(1, na, 1, 1)                    # :431
    d1, d2, i1, i2 = cops.chamfer_forward(transformed.contiguous().view(bz * na, m, 3).numpy(),
                                          expanded.contiguous().view(bz * na, n, 3).numpy())          # :433-435
    d1, d2 = torch.from_numpy(d1).view(bz, na, m), torch.from_numpy(d2).view(bz, na, n)
    cd_r2o, cd_o2r = d1.mean(-1), d2.mean(-1)                                                         # :437-439
    total = cd_o2r if glb_single_cd == 1 else cd_r2o + cd_o2r                                         # :443-446
    minn, orbit = torch.min(total, dim=-1)                                                            # :449
    return dict(d1=d1, d2=d2, i1=torch.from_numpy(i1).view(bz, na, m), i2=torch.from_numpy(i2).view(bz, na, n),
                cd_r2o=cd_r2o, cd_o2r=cd_o2r, minn=minn, orbit=orbit, transformed=transformed)


def pointnet_so3conv(weight, bias, anchors, xyz, feats, return_raw=False):
    """PointnetSO3Conv.forward (vgtk/vgtk/so3conv/modules.py:392-413): centre xyz, rotate it into every anchor frame,
    concatenate to the features, 1x1 conv, max over points.
    weight [Co, C+3] (embed.weight squeezed), bias [Co], anchors [A,3,3], xyz [B,3,N], feats [B,C,N,A]."""
    na = feats.shape[3]
    xyz = xyz - xyz.mean(2, keepdim=True)                                            # :398
    if na == 1:
        cat = torch.cat([feats, xyz[..., None]], 1)                                  # :400-401
    else:
        xyzr = torch.einsum('aji,bjn->bina', anchors.to(xyz.dtype), xyz)             # :404
        cat = torch.cat([feats, xyzr], 1)                                            # :405
    out = torch.einsum('oc,bcna->bona', weight.view(weight.shape[0], -1), cat) + bias.view(1, -1, 1, 1)   # :407
    return out if return_raw else torch.max(out, 2)[0]                               # :408-412
