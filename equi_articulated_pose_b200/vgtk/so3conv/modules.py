"""vgtk.so3conv.modules -- the SO(3) convolution modules (reference: vgtk/vgtk/so3conv/modules.py).

Same constructor signatures, parameter / buffer names (`basic_conv.W`, `anchors`, `kernels`,
`intra_idx`) and forward return tuples as the reference, so reference checkpoints load and
SPConvNets' block builders run unchanged."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from vgtk.spconv import SphericalPointCloud, SphericalPointCloudPose
import vgtk.pc as pctk
from . import functional as L
import vgtk.spconv.functional as zptk_F
from equi_articulated_pose_b200 import ops as _ops

KERNEL_CONDENSE_RATIO = 0.7


class BasicSO3Conv(nn.Module):
    """[b,c1,k,p,a] -> [b,c2,p,a]: out[b,o,p,a] = sum_{c,k} W[o, c*K+k] x[b,c,k,p,a]
    (reference :21-55).  W keeps the reference layout [dim_out, dim_in*kernel_size] (c major)."""

    def __init__(self, dim_in, dim_out, kernel_size, debug=False):
        super().__init__()
        self.dim_in, self.dim_out, self.kernel_size = dim_in, dim_out, kernel_size
        if debug:
            self.register_buffer('W', torch.ones(dim_out, dim_in * kernel_size))
        else:
            W = torch.empty(dim_out, dim_in, kernel_size)
            nn.init.xavier_normal_(W, gain=nn.init.calculate_gain('relu'))
            self.register_parameter('W', nn.Parameter(W.view(dim_out, dim_in * kernel_size)))

    def weight_kc(self):
        """W re-indexed to the kernels' column order k*Ci + c (autograd-tracked view+copy, tiny)."""
        return self.W.view(self.dim_out, self.dim_in, self.kernel_size).transpose(1, 2).reshape(self.dim_out, -1)

    def prepared(self):
        """The PreparedWeight of W when the owning model produced its operand planes for the current parameter value
        (ops.WeightPlanes, one launch per step), else None: callers then re-index / split W per call."""
        wp = getattr(self, '_wp', None)
        return wp if (wp is not None and wp.param is self.W and wp.valid()) else None

    def forward_rows(self, rows):
        """rows [M, K*Ci] (column k*Ci + c) -> [M, dim_out] on the tensor-core GEMM."""
        return _ops.LinearFn.apply(rows, self.weight_kc(), None)

    def forward(self, x):
        bs, c, k, npt, na = x.shape
        rows = x.permute(0, 3, 4, 2, 1).reshape(bs * npt * na, k * c)   # no copy for channels-last views
        out = self.forward_rows(rows)
        return out.view(bs, npt, na, self.dim_out).permute(0, 3, 1, 2)


class KernelPropagation(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError("KernelPropagation is built by no shipped model (SURVEY.md 2.2 #7); out of scope")


class InterSO3Conv(nn.Module):
    """[b,c1,p1,a] -> [b,c2,p2,a]: ball-query neighbourhood, kernel-point correlation under the
    rotation anchors, contraction with W (reference :125-174)."""

    def __init__(self, dim_in, dim_out, kernel_size, stride, radius, sigma, n_neighbor, lazy_sample=True,
                 pooling=None, kanchor=60):
        super().__init__()
        kernels = L.get_sphereical_kernel_points_from_ply(KERNEL_CONDENSE_RATIO * radius, kernel_size)
        anchors = L.get_anchors(kanchor)
        self.dim_in, self.dim_out = dim_in, dim_out
        self.kernel_size = kernels.shape[0]
        self.stride, self.radius, self.sigma, self.n_neighbor = stride, radius, sigma, n_neighbor
        self.lazy_sample, self.pooling = lazy_sample, pooling
        self.basic_conv = BasicSO3Conv(dim_in, dim_out, self.kernel_size)
        self.register_buffer('anchors', torch.from_numpy(np.ascontiguousarray(anchors)))
        self.register_buffer('kernels', torch.from_numpy(np.ascontiguousarray(kernels)))
        self._rk = None

    def rot_kernels(self):
        if self._rk is None or self._rk.device != self.anchors.device:
            self._rk = L.rotated_kernels(self.anchors, self.kernels)
        return self._rk

    def _load_from_state_dict(self, *args, **kwargs):
        self._rk = None                  # `anchors` / `kernels` may be overwritten in place: recompute the (tiny) cache
        return super()._load_from_state_dict(*args, **kwargs)

    def forward(self, x, inter_idx=None, inter_w=None):
        slot, self._grad_slot = getattr(self, '_grad_slot', None), None     # one-shot hand-over set by the owning block
        wp = self.basic_conv.prepared() if x.feats.is_cuda else None
        fused = L.inter_so3conv(x.xyz, x.feats, None if wp is not None else self.basic_conv.weight_kc(), self.stride,
                                self.n_neighbor, self.anchors, self.kernels, self.radius, self.sigma, inter_idx, inter_w,
                                self.lazy_sample, pooling=self.pooling, rot_kernels=self.rot_kernels(), grad_slot=slot, wp=wp)
        if fused is not None:
            inter_idx, inter_w, xyz, feats, sample_idx = fused
            return inter_idx, inter_w, sample_idx, SphericalPointCloud(xyz, feats, self.anchors)
        inter_idx, inter_w, xyz, feats, sample_idx = L.inter_so3conv_grouping(
            x.xyz, x.feats, self.stride, self.n_neighbor, self.anchors, self.kernels, self.radius, self.sigma,
            inter_idx, inter_w, self.lazy_sample, pooling=self.pooling, rot_kernels=self.rot_kernels())
        feats = self.basic_conv(feats)
        return inter_idx, inter_w, sample_idx, SphericalPointCloud(xyz, feats, self.anchors)


class InterSO3PoseConv(InterSO3Conv):
    """Pose-aware variant (reference :177-322).  With an identity per-point pose -- the only case the shipped
    configurations produce -- it is bit-identical to InterSO3Conv and runs the same tuned
    kernels.  With arbitrary per-point rotations (stride 1: the no-stride branch of
    inter_so3poseconv_grouping_strided, functional.py:1061-1261) the neighbour offsets are rotated by R_p R_j^T
    and, for permute_modes != 0, every neighbour's anchors are permuted by the nearest-anchor table; both are
    produced by one kernel instead of the reference's [B,N,nn,A,A,3,3] temporary."""

    def __init__(self, dim_in, dim_out, kernel_size, stride, radius, sigma, n_neighbor, lazy_sample=True,
                 pooling=None, kanchor=60, permute_modes=0, use_2d=False, use_art_mode=False):
        super().__init__(dim_in, dim_out, kernel_size, stride, radius, sigma, n_neighbor, lazy_sample, pooling, kanchor)
        self.permute_modes, self.use_2d, self.use_art_mode = permute_modes, use_2d, use_art_mode

    def forward(self, x, inter_idx=None, inter_w=None, seg=None):
        pose = x.pose
        eye = torch.eye(4, dtype=pose.dtype, device=pose.device)
        if torch.equal(pose, eye.expand_as(pose)):
            # the reference recomputes the ball query every call and ignores a cached inter_idx (:931,1025)
            idx, w, sample_idx, out = super().forward(x, None, None)
            if sample_idx is not None and self.stride > 1:
                sampled_pose = torch.gather(pose, 1, sample_idx.long().view(*sample_idx.shape, 1, 1).expand(-1, -1, 4, 4))
            else:
                sampled_pose = pose
            return idx, w, sample_idx, SphericalPointCloudPose(out.xyz, out.feats, self.anchors, sampled_pose)
        if self.use_2d or self.use_art_mode:
            raise NotImplementedError("InterSO3PoseConv: the 2D / articulation-mode groupings are not reached by the shipped flags")
        xyz = x.xyz.contiguous()
        if self.stride > 1:
            # strided branch (functional.py:931-1029): FPS / lazy sampling of the centres, ball query of the centres in the
            # full cloud, centre rotations pose[sample_idx] (spconv/functional.py:468-500)
            idx, sample_idx, sample_xyz = zptk_F.ball_indices(xyz, self.stride, self.radius, self.n_neighbor, self.lazy_sample)
            sample_xyz = sample_xyz.contiguous()
            sampled_pose = torch.gather(pose, 1, sample_idx.long().view(*sample_idx.shape, 1, 1).expand(-1, -1, 4, 4))
            rel_xyz, perm = _ops.pose_neighbourhood(xyz, pose, idx, self.anchors, with_perm=self.permute_modes != 0,
                                                    sample_xyz=sample_xyz, sample_idx=sample_idx)
        else:
            idx = pctk.ball_query_index(xyz, xyz, self.radius, self.n_neighbor)
            sample_idx, sample_xyz, sampled_pose = None, xyz, pose
            rel_xyz, perm = _ops.pose_neighbourhood(xyz, pose, idx, self.anchors, with_perm=self.permute_modes != 0)
        feats_cl = x.feats.permute(0, 2, 3, 1).contiguous()
        g = _ops.PoseGroupFn.apply(feats_cl, idx, rel_xyz, perm, self.rot_kernels(), self.sigma)
        b, p, a, kc = g.shape
        k = self.kernel_size
        feats = self.basic_conv(g.view(b, p, a, k, kc // k).permute(0, 4, 3, 1, 2))
        if self.stride > 1:
            # the reference clears inter_idx after a strided layer (:1029)
            return None, None, sample_idx, SphericalPointCloudPose(sample_xyz, feats, self.anchors, sampled_pose)
        w = L.LazyInterWeights(xyz, xyz, idx, self.rot_kernels(), self.sigma)
        return idx, w, None, SphericalPointCloudPose(xyz, feats, self.anchors, pose)


class IntraSO3Conv(nn.Module):
    """Group convolution over the 12 nearest rotation anchors (reference :325-347)."""

    def __init__(self, dim_in, dim_out):
        super().__init__()
        anchors = L.get_anchors()
        intra_idx = L.get_intra_idx()
        self.dim_in, self.dim_out = dim_in, dim_out
        self.kernel_size = intra_idx.shape[1]
        self.basic_conv = BasicSO3Conv(dim_in, dim_out, self.kernel_size)
        self.register_buffer('anchors', torch.from_numpy(np.ascontiguousarray(anchors)))
        self.register_buffer('intra_idx', torch.from_numpy(np.ascontiguousarray(intra_idx)).long())
        self._tables = None

    def tables(self):
        """int32 neighbour table [A,12] and its per-column inverse (every column of the icosahedral table is a
        permutation of the anchors) for the gather-GEMM path."""
        if self._tables is None or self._tables[0].device != self.intra_idx.device:
            t = self.intra_idx.to(torch.int32).contiguous()
            inv = torch.empty_like(t)
            na, kk = t.shape
            cols = torch.arange(kk, device=t.device).view(1, kk).expand(na, kk)
            inv[t.long(), cols] = torch.arange(na, device=t.device, dtype=torch.int32).view(na, 1).expand(na, kk)
            is_perm = bool((torch.sort(t, dim=0)[0] == torch.arange(na, device=t.device, dtype=torch.int32).view(na, 1)).all())
            self._tables = (t, inv.contiguous(), is_perm)
        return self._tables

    def _load_from_state_dict(self, *args, **kwargs):
        self._tables = None              # `intra_idx` may be overwritten in place
        return super()._load_from_state_dict(*args, **kwargs)

    def forward(self, x):
        nb, c, npt, na = x.feats.shape
        t, inv, is_perm = self.tables()
        if x.feats.is_cuda and is_perm and _ops.gather_gemm_supported(c, self.dim_out, nb * npt):
            # fused: the [b,c,12,p,a] gather of the reference (functional.py:2565-2567) is never materialised
            wp = self.basic_conv.prepared()
            rows = _ops.IntraConvFn.apply(x.feats.permute(0, 2, 3, 1).contiguous().view(nb * npt, na, c),
                                          self.basic_conv.W if wp is not None else self.basic_conv.weight_kc(), t, inv, wp)
            feats = rows.view(nb, npt, na, self.dim_out).permute(0, 3, 1, 2)
        else:
            feats = L.intra_so3conv_grouping(self.intra_idx, x.feats)
            feats = self.basic_conv(feats)
        return SphericalPointCloud(x.xyz, feats, self.anchors)


class IntraSO3Conv2D(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError("IntraSO3Conv2D: `use_2d` is never forwarded by the shipped models; out of scope")


class PointnetSO3Conv(nn.Module):
    """Equivariant PointNet pooling head (reference :376-413): anchor-rotated centred xyz concatenated to the features,
    1x1 conv, max over points.  Same parameters (`embed.weight [Co, C+3, 1, 1]`, `embed.bias`, buffer `anchors`); the
    [B, C+3, N, A] concatenation is never built: the feature part of the conv is the tcgen05 contraction on the
    channels-last rows, the coordinate part (three FMAs per output) is folded into the pooling pass."""

    def __init__(self, dim_in, dim_out, kanchor=60, return_raw=False):
        super().__init__()
        anchors = L.get_anchors(kanchor)
        self.dim_in, self.dim_out, self.return_raw = dim_in + 3, dim_out, return_raw
        self.embed = nn.Conv2d(self.dim_in, self.dim_out, 1)
        self.register_buffer('anchors', torch.from_numpy(np.ascontiguousarray(anchors)))

    def forward(self, x):
        xyz, feats = x.xyz, x.feats
        nb, nc, npt, na = feats.shape
        xc = (xyz - xyz.mean(2, keepdim=True)).contiguous() if getattr(self, 'center', True) else xyz.contiguous()
        w = self.embed.weight.view(self.dim_out, self.dim_in)
        w_f, w_x = w[:, :nc].contiguous(), w[:, nc:]
        if na == 1:
            v = w_x.unsqueeze(0)                                        # xyz itself is concatenated (:400-401)
        else:
            v = torch.einsum('oi,aji->aoj', w_x, self.anchors)          # W_x applied to R_a^T xyz (:404)
        rows = L._channels_last(feats).view(nb * npt * na, nc)
        e = _ops.LinearFn.apply(rows, w_f, self.embed.bias).view(nb, npt, na, self.dim_out)
        if self.return_raw:
            if torch.is_grad_enabled() and (e.requires_grad or v.requires_grad):
                term = torch.einsum('aoj,bjn->bnao', v, xc)             # differentiable form of the same three FMAs
                return (e + term).permute(0, 3, 1, 2)
            return _ops.pointnet_embed_xyz_(e, v.contiguous(), xc).permute(0, 3, 1, 2)
        return _ops.PointnetPoolFn.apply(e, v.contiguous(), xc)


class PointnetSO3PoseConv(PointnetSO3Conv):
    """reference :416-...: identical pooling on a pose-carrying cloud."""
    pass
