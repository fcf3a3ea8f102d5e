"""vgtk.spconv.functional -- ball-query grouping and index helpers
(reference: vgtk/vgtk/spconv/functional.py).  Index work runs in libvgtkb200; the helper
functions that the reference evaluates with torch indexing keep their torch form."""
import math

import numpy as np
import torch
import torch.nn.functional as F

import vgtk.pc as pctk
from equi_articulated_pose_b200 import ops as _ops


# ---- shadow point / feature (reference :83-96).  Index N is never produced by ball_query, so the
# fused path does not need them; they are kept for callers of the literal API.
def add_shadow_point(x):
    b, c, _ = x.shape
    return torch.cat((x, torch.full((b, c, 1), 1e4, dtype=torch.float32, device=x.device)), dim=2).contiguous()


def add_shadow_feature(x):
    b, c, _, a = x.shape
    return torch.cat((x, torch.zeros(b, c, 1, a, dtype=torch.float32, device=x.device)), dim=2).contiguous()


class Gathering(torch.autograd.Function):
    """[nb,c,np] x [nb,m] -> [nb,c,m] with a scatter-add backward (reference :102-129)."""

    @staticmethod
    def forward(ctx, points, idx):
        ctx.save_for_backward(idx)
        ctx.n = points.size(2)
        return _ops.gather_points_forward(points, idx)

    @staticmethod
    def backward(ctx, grad):
        (idx,) = ctx.saved_tensors
        return _ops.gather_points_backward(grad.contiguous(), idx, ctx.n), None


def ball_query(query_points, support_points, radius, n_sample, support_feats=None):
    """[b,3,m] x [b,3,n] -> idx [b,m,k], grouped xyz [b,3,m,k] (reference :341-350)."""
    idx = pctk.ball_query_index(query_points, support_points, radius, n_sample)
    if support_feats is None:
        return idx, pctk.group_nd(support_points, idx)
    return idx, pctk.group_nd(support_points, idx), pctk.group_nd(support_feats, idx)


def batched_index_select(input, dim, index):
    """reference :364-372"""
    for ii in range(1, len(input.shape)):
        if ii != dim:
            index = index.unsqueeze(ii)
    expanse = list(input.shape)
    expanse[0] = -1
    expanse[dim] = -1
    return torch.gather(input, dim, index.expand(expanse))


def batched_index_select_other(values, indices, dim=1):
    """values [b, n, ...], indices [b, m1(, m2..)] -> [b, m1(, m2..), ...] (reference :452-466)"""
    if dim != 1:
        raise NotImplementedError("batched_index_select_other: only dim=1 is used by the reference")
    b = values.shape[0]
    trailing = values.shape[2:]
    flat = indices.reshape(b, -1)
    view = flat.view(b, -1, *([1] * len(trailing))).expand(b, flat.shape[1], *trailing)
    return torch.gather(values, 1, view).view(*indices.shape, *trailing)


def inter_zpconv_grouping_naive(inter_idx, inter_w, feats):
    """G[b,c,k,p,a] = sum_n feats[b,c,idx[b,p,n],a] w[b,p,a,k,n] with EXPLICIT weights
    (reference :375-406).  Accepts feats with or without the shadow row."""
    if hasattr(inter_w, 'materialize'):
        inter_w = inter_w.materialize()
    b, p, nn = inter_idx.shape
    c, a = feats.shape[1], feats.shape[3]
    flat = inter_idx.long().reshape(b, 1, p * nn, 1).expand(-1, c, -1, a)
    nb = torch.gather(feats, 2, flat).view(b, c, p, nn, a)
    return torch.einsum('bcpna,bpakn->bckpa', nb, inter_w).contiguous()


def inter_pooling_naive(inter_idx, sample_idx, feats, alpha=0.5):
    b, p, pnn = inter_idx.shape
    a = feats.shape[3]
    new_feats = batched_index_select(feats, 2, sample_idx.long())
    grouped = batched_index_select(feats, 2, inter_idx.long().view(b, -1)).view(b, -1, p, pnn, a)
    return alpha * new_feats + (1 - alpha) * grouped.mean(3)


def inter_blurring_naive(inter_idx, feats, alpha=0.5):
    b, p, pnn = inter_idx.shape
    a = feats.shape[3]
    grouped = batched_index_select(feats, 2, inter_idx.long().view(b, -1)).view(b, -1, p, pnn, a)
    return alpha * feats + (1 - alpha) * grouped.mean(3)


def inter_zpconv_grouping_ball(xyz, stride, radius, n_neighbor, lazy_sample=True):
    """[b,3,n] -> grouped_xyz [b,3,p,nn] (centre relative), ball_idx [b,p,nn] int32,
    sample_idx [b,p], sample_xyz [b,3,p]   (reference :428-449)."""
    n_sample = math.ceil(xyz.shape[2] / stride)
    if stride > 1:
        idx, sample_xyz = pctk.furthest_sample(xyz, n_sample, lazy_sample)
    else:
        sample_xyz = xyz
        idx = torch.arange(xyz.shape[2], dtype=torch.long, device=xyz.device).unsqueeze(0).repeat(xyz.shape[0], 1)
    ball_idx, grouped_xyz = ball_query(sample_xyz, xyz, radius, n_neighbor)
    grouped_xyz = grouped_xyz - sample_xyz.unsqueeze(3)
    return grouped_xyz, ball_idx, idx, sample_xyz


def ball_indices(xyz, stride, radius, n_neighbor, lazy_sample=True):
    """Index-only form of inter_zpconv_grouping_ball used by the fused path: no grouped_xyz
    tensor is built (the kernels recompute the offsets from xyz).  -> ball_idx, sample_idx, sample_xyz"""
    n_sample = math.ceil(xyz.shape[2] / stride)
    if stride > 1:
        idx, sample_xyz = pctk.furthest_sample(xyz, n_sample, lazy_sample)
    else:
        sample_xyz = xyz
        idx = torch.arange(xyz.shape[2], dtype=torch.long, device=xyz.device).unsqueeze(0).repeat(xyz.shape[0], 1)
    ball_idx = pctk.ball_query_index(sample_xyz, xyz, radius, n_neighbor)
    return ball_idx, idx, sample_xyz
