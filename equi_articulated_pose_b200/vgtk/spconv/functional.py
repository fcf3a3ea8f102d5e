"""vgtk.spconv.functional -- ball-query grouping and index helpers
(reference: vgtk/vgtk/spconv/functional.py).  Index work runs in libvgtkb200; the helper
functions that the reference evaluates with torch indexing keep their torch form."""
import math

import numpy as np
import torch
import torch.nn.functional as F

import vgtk.pc as pctk
from equi_articulated_pose_b200 import ops as _ops


# ---- legacy S^2 anchors / intra kernel weights (reference :20-39, :142-207); constants of the legacy ZPConv modules
def get_anchors(anchor):
    """12 S^2 direction anchors = the unit icosahedron vertices of data/anchors/sphere12.ply (norm filter > 0.5)."""
    if isinstance(anchor, torch.Tensor):
        return anchor.detach().cpu()
    if isinstance(anchor, int) and anchor == 12:
        from equi_articulated_pose_b200 import so3_constants as _C
        pts = _C._load()["ico_vertices"].astype("float32")
    elif isinstance(anchor, str):
        pts = pctk.load_ply(anchor).astype("float32")
    else:
        raise ValueError(f"Not recognized anchor type {type(anchor)} (only the 12-vertex sphere ships with this package)")
    norms = np.sqrt(np.sum(pts ** 2, axis=1))
    keep = norms > 0.5
    return torch.from_numpy(pts[keep] / norms[keep, None])


def get_intra_kernels(aperature, kernel_size):
    return torch.from_numpy(np.linspace(0, 0.5 * aperature, kernel_size, dtype=np.float32))


def acos_safe(x, eps=1e-4):
    """acos continued linearly outside [-1+eps, 1-eps] (reference :148-154)."""
    sign = torch.sign(x)
    slope = np.arccos(1 - eps) / eps
    return torch.where(abs(x) <= 1 - eps, torch.acos(x), torch.acos(sign * (1 - eps)) - slope * sign * (abs(x) - 1 + eps))


def anchor_knn(a_src, a_tgt, k=3, metric="spherical"):
    """for every anchor of a_tgt the k nearest anchors of a_src -> (values [a,k], indices [a,k]) (reference :156-175)."""
    a_src, a_tgt = a_src.unsqueeze(0), a_tgt.unsqueeze(1)
    if metric == "spherical":
        return (torch.sum(a_src * a_tgt, dim=2) - 1.0).topk(k=k, dim=1, largest=True)
    if metric == "angular":
        return acos_safe(torch.sum(a_src * a_tgt, dim=2)).topk(k=k, dim=1, largest=False)
    return torch.sum((a_src - a_tgt) ** 2, dim=2).topk(k=k, dim=1, largest=False)


def get_intra_kernel_weights(anchor_in, anchor_out, kernels, ann, aperature, sigma=1e-1, use_suppression=False):
    """idx [a_out, ann] int32 (nearest input anchors by angle) and the linear angular kernel weights
    influence [a_out, ks, ann] = relu(1 - |angle - kernel| / pi / (3 sqrt(sigma/2)))  (reference :179-218)."""
    anchor_out = anchor_in if anchor_out is None else anchor_out
    angles, idx = anchor_knn(anchor_in, anchor_out, k=ann, metric="angular")
    influence = (angles.unsqueeze(1) - kernels.unsqueeze(0).unsqueeze(-1)).abs() / np.pi
    influence = F.relu(1.0 - influence / (3 * (sigma / 2.0) ** 0.5))
    if use_suppression:
        influence = influence * angles.le(0.5 * aperature).unsqueeze(1).expand(-1, kernels.size(0), -1).float()
    return idx.int().contiguous(), influence.contiguous()


class IntraZPConvGrouping(torch.autograd.Function):
    """[nb,c,np,na_in] -> [nb,c,ks,np,na_out] with explicit (idx, w) (reference :222-248), on the CUDA slot kernels."""

    @staticmethod
    def forward(ctx, intra_idx, intra_w, feats):
        import vgtk.cuda.zpconv as cuda_zpconv
        ctx.save_for_backward(intra_idx, intra_w)
        ctx.ain = feats.shape[3]
        return cuda_zpconv.intra_zpconv_forward(intra_idx, intra_w, feats)

    @staticmethod
    def backward(ctx, grad):
        import vgtk.cuda.zpconv as cuda_zpconv
        intra_idx, intra_w = ctx.saved_tensors
        return None, None, cuda_zpconv.intra_zpconv_backward(intra_idx, intra_w, grad.contiguous(), ctx.ain)


def intra_zpconv_grouping(intra_idx, intra_w, feats):
    return IntraZPConvGrouping.apply(intra_idx, intra_w, feats)


def intra_zpconv_grouping_naive(intra_idx, intra_w, feats):
    """G[b,c,k,p,a] = sum_n feats[b,c,p,idx[a,n]] w[a,k,n] (reference :252-272 evaluates index_select + einsum)."""
    return IntraZPConvGrouping.apply(intra_idx, intra_w, feats)


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
        idx = pctk.identity_index(xyz.shape[0], xyz.shape[2], xyz.device, torch.long)
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
        idx = pctk.identity_index(xyz.shape[0], xyz.shape[2], xyz.device, torch.long)
    ball_idx = pctk.ball_query_index(sample_xyz, xyz, radius, n_neighbor)
    return ball_idx, idx, sample_xyz
