"""Invariant output heads of SPConvNets/utils/base_so3conv.py on the B200 kernels.

Mirror of InvOutBlockR (:481-550), InvOutBlockPointnet (:553-601), InvOutBlockMVD (:603-645), InvOutBlockOurs (:766-840),
InvOutBlockOursWithMask (:1013-1150, the head model 38 builds) and PointnetSO3ConvOurs (:1153-1214): same constructor
arguments, same module / parameter names (a reference state_dict loads), same returned tuples.  The full-size work -- the
1x1 convolutions over [B, C, N, A], their InstanceNorm / BatchNorm + relu, the PointNet embedding and its pooling over the
points -- runs on the channels-last rows through the tcgen05 contraction (ops.LinearFn), the fused norm passes
(ops.norm_act) and the pooling kernels (vgtk.so3conv.PointnetSO3Conv); what is left on torch is the tail on the pooled
[B, C, A] tensor (BatchNorm1d, the 1-channel attention conv, softmax over the 60 anchors).

The reference's own, unmodified classes also run on top of the `vgtk` drop-in (module level); this file is the block-level
fast path, like blocks.py for the backbone.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

import equi_articulated_pose_b200 as _pkg

_pkg.install()
import vgtk.spconv as zptk  # noqa: E402
import vgtk.so3conv as sptk  # noqa: E402
import vgtk.so3conv.functional as L  # noqa: E402
from . import ops as _ops  # noqa: E402
from .blocks import FusedBatchNorm2d, FusedInstanceNorm2d, _rows, _unrows  # noqa: E402


def _conv_rows(conv, rows):
    """1x1 nn.Conv2d applied to channels-last rows [M, C_in] -> [M, C_out] (tensor-core contraction)."""
    w = conv.weight.view(conv.out_channels, conv.in_channels)
    return _ops.LinearFn.apply(rows, w, conv.bias)


def _pool_anchors(self, x_out):
    """[B, C, A] -> ([B, C], attention logits or None): the pooling tail shared by the heads."""
    if self.pooling_method == 'mean':
        return x_out.mean(dim=2), None
    if self.pooling_method == 'max':
        return x_out.max(2)[0], None
    if self.pooling_method.startswith('attention'):
        att = self.attention_layer(x_out)
        confidence = F.softmax(att * self.temperature, dim=2)
        return (x_out * confidence).sum(-1), att
    raise NotImplementedError(f"Pooling mode {self.pooling_method} is not implemented!")


class PointnetSO3ConvOurs(sptk.PointnetSO3Conv):
    """base_so3conv.py:1153-1214: PointnetSO3Conv with the option of absolute (not centred) coordinates.  (The 2-D residual
    anchors `tot_anchors` of the reference only serve the `use_2d` variants, which no shipped model forwards.)"""

    def __init__(self, dim_in, dim_out, kanchor=60, return_raw=False, use_abs_pos=False):
        super().__init__(dim_in, dim_out, kanchor, return_raw)
        self.use_abs_pos = use_abs_pos
        self.center = not use_abs_pos


class InvOutBlockR(nn.Module):
    """base_so3conv.py:481-550: 1x1 convs (+ InstanceNorm + relu between them), mean over points, pooling over anchors."""

    def __init__(self, params, norm=None):
        super().__init__()
        c_in, mlp = params['dim_in'], params['mlp']
        self.pooling_method = params.get('pooling', 'max')
        self.norm = nn.ModuleList()
        if self.pooling_method == 'attention':
            self.temperature = params['temperature']
            self.attention_layer = nn.Conv1d(mlp[-1], 1, 1)
        self.linear = nn.ModuleList()
        for c in mlp:
            self.linear.append(nn.Conv2d(c_in, c, 1))
            self.norm.append(FusedInstanceNorm2d(c, affine=False))
            c_in = c

    def forward(self, feats):
        rows, (b, p, a, _) = _rows(feats)
        end = len(self.linear)
        for lid, linear in enumerate(self.linear):
            rows = _conv_rows(linear, rows)
            if lid != end - 1:
                rows = self.norm[lid].forward_rows(rows, b, 0.0)
        x_out = _unrows(rows, b, p, a)                       # logical [B, C, N, A]
        out_feat = x_out.mean(2)
        if self.pooling_method == 'mean':
            x_out = x_out.mean(dim=3).mean(dim=2)
        elif self.pooling_method == 'debug':
            x_out = x_out[..., 0].mean(2)
        elif self.pooling_method == 'max':
            x_out = out_feat.max(-1)[0]
        elif self.pooling_method == 'attention':
            x_out = out_feat
            out_feat = self.attention_layer(x_out)
            confidence = F.softmax(out_feat * self.temperature, dim=2)
            x_out = (x_out * confidence).sum(-1)
            out_feat = confidence.squeeze()
        else:
            raise NotImplementedError(f"Pooling mode {self.pooling_method} is not implemented!")
        return F.normalize(x_out, p=2, dim=1), out_feat


class InvOutBlockPointnet(nn.Module):
    """base_so3conv.py:553-601."""

    def __init__(self, params, norm=None):
        super().__init__()
        c_in, c_out, na = params['dim_in'], params['mlp'][-1], params['kanchor']
        self.pooling_method = params.get('pooling', 'max')
        self.pointnet = sptk.PointnetSO3Conv(c_in, c_out, na)
        if self.pooling_method == 'attention':
            self.temperature = params['temperature']
            self.attention_layer = nn.Conv1d(c_out, 1, 1)

    def forward(self, x):
        x_out = self.pointnet(x)                              # [B, C, A]
        out_feat = x_out
        x_out, _ = _pool_anchors(self, x_out)
        return F.normalize(x_out, p=2, dim=1), F.normalize(out_feat, p=2, dim=1)


class InvOutBlockMVD(nn.Module):
    """base_so3conv.py:603-645: per-point attention over the anchors, then the PointNet head on the single-anchor cloud."""

    def __init__(self, params, norm=None):
        super().__init__()
        c_in, c_out, na = params['dim_in'], params['mlp'][-1], params['kanchor']
        self.temperature = params['temperature']
        self.attention_layer = nn.Sequential(nn.Conv2d(c_in, c_in, 1), nn.ReLU(inplace=True), nn.Conv2d(c_in, c_in, 1))
        self.pooling_method = params.get('pooling', 'max')
        self.pointnet = sptk.PointnetSO3Conv(c_in, c_out, na)

    def forward(self, x):
        nb, nc, npt, na = x.feats.shape
        rows, _ = _rows(x.feats)
        att = _conv_rows(self.attention_layer[2], F.relu(_conv_rows(self.attention_layer[0], rows)))
        attn = F.softmax(_unrows(att, nb, npt, na), dim=3)
        x_out = (x.feats * attn).sum(-1, keepdim=True)
        x_out = self.pointnet(zptk.SphericalPointCloud(x.xyz, x_out, None)).view(nb, -1)
        return F.normalize(x_out, p=2, dim=1), attn


class _OursBase(nn.Module):
    def _mlp(self, rows, b):
        for lid, linear in enumerate(self.linear):
            rows = self.norm[lid].forward_rows(_conv_rows(linear, rows), 0.0)
        return rows


class InvOutBlockOurs(_OursBase):
    """base_so3conv.py:766-840."""

    def __init__(self, params, norm=None, pooling_method='max'):
        super().__init__()
        c_in, mlp, na = params['dim_in'], params['mlp'], params['kanchor']
        self.outDim = params['k']
        self.linear, self.norm = nn.ModuleList(), nn.ModuleList()
        for c in mlp:
            self.linear.append(nn.Conv2d(c_in, c, 1))
            self.norm.append(FusedBatchNorm2d(c))
            c_in = c
        self.pooling_method = pooling_method
        if self.pooling_method == 'attention':
            self.temperature = params['temperature']
            self.attention_layer = nn.Conv1d(c_in, 1, 1)
        self.pointnet = sptk.PointnetSO3Conv(c_in, c_in, na)
        self.norm.append(nn.BatchNorm1d(c_in))
        self.fc2 = nn.Linear(c_in, self.outDim)

    def forward(self, x, label=None):
        rows, (b, p, a, _) = _rows(x.feats)
        rows = self._mlp(rows, b)
        x_out = self.pointnet(zptk.SphericalPointCloud(x.xyz, _unrows(rows, b, p, a), x.anchors))    # [B, C, A]
        x_out = F.relu(self.norm[len(self.linear)](x_out))
        if self.pooling_method == 'debug':
            return x_out[..., 0].mean(2)
        return _pool_anchors(self, x_out)[0]


class InvOutBlockOursWithMask(_OursBase):
    """base_so3conv.py:1013-1150 (the invariant head of model 38): masked features -> 1x1 convs + BatchNorm + relu ->
    PointNet embedding (raw) -> (soft-)masked mean over the points -> BatchNorm1d + relu -> pooling over the anchors."""

    def __init__(self, params, norm=None, pooling_method='max', use_pointnet=True, sel_mode=None, use_abs_pos=False,
                 return_point_pooling_feature=False):
        super().__init__()
        c_in, mlp, na = params['dim_in'], params['mlp'], params['kanchor']
        self.outDim = params['k']
        self.linear, self.norm = nn.ModuleList(), nn.ModuleList()
        self.use_pointnet, self.sel_mode, self.use_abs_pos = use_pointnet, sel_mode, use_abs_pos
        self.return_point_pooling_feature = return_point_pooling_feature
        for c in mlp:
            self.linear.append(nn.Conv2d(c_in, c, 1))
            self.norm.append(FusedBatchNorm2d(c))
            c_in = c
        self.pooling_method = pooling_method if self.sel_mode is None else 'sel_mode'
        if self.pooling_method == 'attention':
            self.temperature = params['temperature']
            self.attention_layer = nn.Conv1d(c_in, 1, 1)
        if self.use_pointnet:
            self.pointnet = PointnetSO3ConvOurs(c_in, c_in, na, return_raw=True, use_abs_pos=self.use_abs_pos)
            self.norm.append(nn.BatchNorm1d(c_in))

    def forward(self, x, mask, label=None, soft_mask=None):
        x_out, x_xyz = x.feats, x.xyz
        out_feat = None
        if mask is not None:
            x_out = x_out * mask.unsqueeze(1).unsqueeze(-1)
            x_xyz = x_xyz * mask.unsqueeze(1)
        rows, (b, p, a, _) = _rows(x_out)
        rows = self._mlp(rows, b)
        x_out = _unrows(rows, b, p, a)
        if mask is not None:
            x_out = x_out * mask.unsqueeze(1).unsqueeze(-1)
        if not self.use_pointnet:
            if soft_mask is not None:
                x_out = torch.sum(x_out, dim=2) / torch.clamp(torch.sum(soft_mask.unsqueeze(1).unsqueeze(-1), dim=2), min=1e-8)
            else:
                x_out = torch.mean(x_out, dim=2)
        else:
            out_feat = x_out                                                              # what the reference returns for non-attention pooling
            x_out = self.pointnet(zptk.SphericalPointCloud(x_xyz, x_out, x.anchors))      # raw: [B, C, N, A]
            if soft_mask is not None:
                sm = soft_mask.unsqueeze(1).unsqueeze(-1)
                x_out = torch.sum(x_out * sm, dim=2) / torch.clamp(torch.sum(sm, dim=2), min=1e-8)
            else:
                x_out = torch.mean(x_out, dim=2)
            x_out = F.relu(self.norm[len(self.linear)](x_out))
        x_out_points_pooling = x_out.clone()
        if self.pooling_method == 'debug':
            x_out = x_out[..., 0].mean(2)
        elif self.pooling_method == 'sel_mode':
            x_out = x_out[..., self.sel_mode]
        else:
            x_out, att = _pool_anchors(self, x_out)
            if att is not None:
                out_feat = att
        if self.return_point_pooling_feature:
            return x_out_points_pooling, x_out, out_feat.squeeze(1)
        return x_out, out_feat.squeeze(1)
