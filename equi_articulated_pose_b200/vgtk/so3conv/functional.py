"""vgtk.so3conv.functional -- SO(3) anchors, kernel points and the inter-/intra-anchor grouping
(reference: vgtk/vgtk/so3conv/functional.py).

The grouping functions keep the reference's names, argument order and returned tuples.  What
changes is what the returned objects ARE: `new_feats` is a permuted view of a channels-last
buffer produced by the fused sm_100a kernel, and `inter_w` is a `LazyInterWeights` handle -- the
[B,P,A,K,nn] tensor the reference materialises (755 MB per layer at config 2) exists only if a
caller asks for it with `.materialize()`."""
import numpy as np
import torch
import torch.nn.functional as F

import vgtk.pc as pctk
import vgtk.spconv as zpconv
from equi_articulated_pose_b200 import ops as _ops
from equi_articulated_pose_b200 import so3_constants as _C

inter_so3conv_feat_grouping = zpconv.inter_zpconv_grouping_naive
batched_index_select = zpconv.batched_index_select
batched_index_select_other = zpconv.batched_index_select_other

GAMMA_SIZE = 3
Rs, R_idx, canonical_relative = _C.anchors_all(), _C.intra_idx(), None


def select_anchor(anchors, k):
    return _C.select_anchor(anchors, k)


def get_anchors(k=60):
    return _C.get_anchors(k)


def get_intra_idx():
    return R_idx


def get_canonical_relative():
    return canonical_relative


def get_occupancy_features(pc, n_anchor, use_center=False):
    """pc [nb,np,3] -> ones [nb,1,np,na] (reference :50-69; the normals branch of the reference
    references an undefined name and cannot run, so only xyz input is supported)."""
    nb, npts, nd = pc.shape
    if nd == 6:
        raise NotImplementedError("occupancy features from normals: unreachable in the reference (NameError at :61)")
    features = torch.ones(nb, 1, npts, n_anchor, dtype=torch.float32, device=pc.device)
    if use_center:
        features[:, :, 0, :] = 0.0
    return features


def get_sphereical_kernel_points_from_ply(radius, kernel_size):
    """24/30/66 kernel points scaled to `radius` (reference :111-121); kernel_size 1 only."""
    assert kernel_size <= 3 and kernel_size > 0
    return _C.scaled_kernel_points(radius, kernel_size)


def rotated_kernels(anchors, kernels):
    """[A,3,3] x [K,3] -> R_a kappa_k as [A,K,3] (reference :2519 builds the [3,A,K] permutation)."""
    return torch.matmul(anchors, kernels.transpose(0, 1)).permute(0, 2, 1).contiguous()


class LazyInterWeights:
    """Stand-in for inter_w [B,P,A,K,nn]: everything needed to recompute it on the fly."""

    def __init__(self, xyz, sample_xyz, idx, rot_kernels, sigma):
        self.xyz, self.sample_xyz, self.idx, self.rot_kernels, self.sigma = xyz, sample_xyz, idx, rot_kernels, sigma

    @property
    def shape(self):
        b, p, nn = self.idx.shape
        a, k = self.rot_kernels.shape[:2]
        return torch.Size((b, p, a, k, nn))

    def size(self, *d):
        return self.shape if not d else self.shape[d[0]]

    def materialize(self):
        return _ops.inter_weights(self.xyz, self.sample_xyz, self.idx, self.rot_kernels, self.sigma)


def inter_so3conv_grouping_anchor(grouped_xyz, anchors, kernels, sigma, interpolate='linear'):
    """Materialised kernel-point correlation from grouped offsets [b,3,p,nn] -> [b,p,a,k,nn]
    (reference :2508-2549).  API parity; the conv path uses the fused kernel instead."""
    if interpolate != 'linear':
        raise NotImplementedError("kernel function %s is not implemented!" % interpolate)
    b, _, p, nn = grouped_xyz.shape
    rk = rotated_kernels(anchors, kernels)
    # offsets as a 'cloud' of p*nn points around the origin: idx = identity, sample_xyz = 0
    xyz = grouped_xyz.reshape(b, 3, p * nn).contiguous()
    idx = torch.arange(p * nn, dtype=torch.int32, device=xyz.device).view(1, p, nn).expand(b, -1, -1).contiguous()
    zero = torch.zeros(b, 3, p, dtype=torch.float32, device=xyz.device)
    return _ops.inter_weights(xyz, zero, idx, rk, sigma)


def inter_so3conv_blurring(xyz, feats, n_neighbor, radius, stride, inter_idx=None, lazy_sample=True,
                           radius_expansion=1.0):
    if inter_idx is None:
        inter_idx, sample_idx, sample_xyz = zpconv.functional.ball_indices(xyz, stride, radius * radius_expansion,
                                                                           n_neighbor, lazy_sample)
    if stride == 1:
        return zpconv.inter_blurring_naive(inter_idx, feats), xyz
    return zpconv.inter_pooling_naive(inter_idx, sample_idx, feats), sample_xyz


def _channels_last(feats):
    """logical [B,C,N,A] -> contiguous [B,N,A,C] (no copy when already channels-last)."""
    return feats.permute(0, 2, 3, 1).contiguous()


def _inter_so3conv_indices(xyz, feats, stride, n_neighbor, anchors, kernels, radius, sigma, inter_idx, inter_w, lazy_sample,
                           radius_expansion, pooling, rot_kernels):
    """Index part of inter_so3conv_grouping (reference :144-189): optional pooling, ball query / cached indices.
    -> (xyz, feats, inter_idx, inter_w, new_xyz, sample_idx)."""
    if pooling is not None and stride > 1 and feats.shape[1] > 1:
        if pooling == 'stride':
            pool_stride, stride_nn, stride = stride, int(n_neighbor * stride ** 0.5), 1
        elif pooling == 'no-stride':
            pool_stride, stride_nn = 1, n_neighbor
        else:
            raise NotImplementedError(f"Pooling mode {pooling} is not implemented!")
        feats, xyz = inter_so3conv_blurring(xyz, feats, stride_nn, radius, pool_stride, inter_idx, lazy_sample)
        inter_idx = None

    xyz = xyz.contiguous()
    if inter_idx is None:
        inter_idx, sample_idx, new_xyz = zpconv.functional.ball_indices(xyz, stride, radius * radius_expansion,
                                                                        n_neighbor, lazy_sample)
        if rot_kernels is None:
            rot_kernels = rotated_kernels(anchors, kernels)
        inter_w = LazyInterWeights(xyz, new_xyz.contiguous(), inter_idx, rot_kernels, sigma)
    else:
        sample_idx, new_xyz = None, xyz
    return xyz, feats, inter_idx, inter_w, new_xyz, sample_idx


def inter_so3conv_grouping(xyz, feats, stride, n_neighbor, anchors, kernels, radius, sigma, inter_idx=None,
                           inter_w=None, lazy_sample=True, radius_expansion=1.0, pooling=None, rot_kernels=None):
    """Reference :144-203.  xyz [b,3,p1], feats [b,c,p1,a] ->
    (inter_idx [b,p2,nn], inter_w (lazy), new_xyz [b,3,p2], new_feats [b,c,k,p2,a] (view), sample_idx)."""
    xyz, feats, inter_idx, inter_w, new_xyz, sample_idx = _inter_so3conv_indices(
        xyz, feats, stride, n_neighbor, anchors, kernels, radius, sigma, inter_idx, inter_w, lazy_sample, radius_expansion,
        pooling, rot_kernels)
    if not isinstance(inter_w, LazyInterWeights):
        # explicit weights handed in by the caller: literal (unfused) evaluation
        new_feats = inter_so3conv_feat_grouping(inter_idx, inter_w, feats)
        return inter_idx, inter_w, new_xyz, new_feats, sample_idx

    w = inter_w
    g = _ops.InterGroupFn.apply(_channels_last(feats), w.xyz, w.sample_xyz, w.idx, w.rot_kernels, w.sigma)
    b, p, a, kc = g.shape
    k = w.rot_kernels.shape[1]
    new_feats = g.view(b, p, a, k, kc // k).permute(0, 4, 3, 1, 2)          # logical [b,c,k,p,a]
    return inter_idx, inter_w, new_xyz, new_feats, sample_idx


def inter_so3conv(xyz, feats, w_kc, stride, n_neighbor, anchors, kernels, radius, sigma, inter_idx=None, inter_w=None,
                  lazy_sample=True, radius_expansion=1.0, pooling=None, rot_kernels=None, grad_slot=None, wp=None):
    """inter_so3conv_grouping + BasicSO3Conv in one call (vgtkb_inter_conv_forward) when the shape is taken: the grouped
    tensor of the reference (:192-203) only exists as the contraction's bf16 operand planes.
    -> (inter_idx, inter_w, new_xyz, out_feats logical [b,co,p2,a], sample_idx), or None when the caller has to take the
    two-step path (explicit weights, shapes outside the fused kernels, contraction modes other than bf16x3)."""
    if not feats.is_cuda or (inter_idx is not None and not isinstance(inter_w, LazyInterWeights)):
        return None
    xyz, feats, inter_idx, inter_w, new_xyz, sample_idx = _inter_so3conv_indices(
        xyz, feats, stride, n_neighbor, anchors, kernels, radius, sigma, inter_idx, inter_w, lazy_sample, radius_expansion,
        pooling, rot_kernels)
    w = inter_w
    b, ci, n, a = feats.shape
    p, nn = w.idx.shape[1], w.idx.shape[2]
    # wp: ops.PreparedWeight of the conv weight (operand planes produced once per step); w_kc may then be None
    k, co = w.rot_kernels.shape[1], (wp.co if wp is not None else w_kc.shape[0])
    if not _ops.inter_conv_supported(b, n, p, nn, a, k, ci, co):
        if w_kc is None:
            w_kc = wp.kc()
        g = _ops.InterGroupFn.apply(_channels_last(feats), w.xyz, w.sample_xyz, w.idx, w.rot_kernels, w.sigma)
        rows = _ops.LinearFn.apply(g.view(b * p * a, k * ci), w_kc, None)
    elif wp is not None and (wp.ci, wp.k) == (ci, k):
        rows = _ops.InterConvFn.apply(_channels_last(feats), wp.matrix(), w.xyz, w.sample_xyz, w.idx, w.rot_kernels, w.sigma,
                                      grad_slot if pooling is None else None, wp)
    else:
        if w_kc is None:
            w_kc = wp.kc()
        rows = _ops.InterConvFn.apply(_channels_last(feats), w_kc, w.xyz, w.sample_xyz, w.idx, w.rot_kernels, w.sigma,
                                      grad_slot if pooling is None else None)
    out = rows.view(b, p, a, co).permute(0, 3, 1, 2)
    return inter_idx, inter_w, new_xyz, out, sample_idx


def intra_so3conv_grouping(intra_idx, feature):
    """feature [nb,c,np,na] -> [nb,c,pnn,np,na] with out[b,c,k,p,a] = feature[b,c,p,intra_idx[a,k]]
    (reference :2553-2567); returned as a view of the channels-last gather."""
    nb, c, nq, na = feature.shape
    kk = intra_idx.shape[1]
    y = _channels_last(feature).view(nb * nq, na, c)
    g = _ops.IntraGroupFn.apply(y, intra_idx.to(torch.int32).contiguous())
    return g.view(nb, nq, na, kk, c).permute(0, 4, 3, 1, 2)
