"""vgtk.spconv.modules -- legacy S^2-anchor ZPConv modules (reference: vgtk/vgtk/spconv/modules.py).

`BasicZPConv` and `IntraZPConv` (BASELINE config 1b: 12 direction anchors) are built: the grouping runs in the
`vgtk.cuda.zpconv` slot kernels of libvgtkb200, the contraction on the same GEMM as the SO(3) path.  The shipped
models only use the SO(3) modules; `InterZPConv` / `AnchorProp` (which need the dead `anchor_query` kernel,
SURVEY.md 2.2 #5) are outside the hot-path scope and raise on construction."""
import torch
import torch.nn as nn

from vgtk.spconv import SphericalPointCloud
from . import functional as L
from equi_articulated_pose_b200 import ops as _ops


class BasicZPConv(nn.Module):
    """[b,c1,k,p,a] -> [b,c2,p,a]: W [c2, c1*k] (c major) and a bias initialised to 1e-3 (reference :17-56)."""

    def __init__(self, dim_in, dim_out, kernel_size, debug=False):
        super().__init__()
        self.dim_in, self.dim_out, self.kernel_size = dim_in, dim_out, kernel_size
        if debug:
            self.register_buffer('W', torch.ones(dim_out, dim_in * kernel_size))
            self.bias = None
        else:
            W = torch.empty(dim_out, dim_in, kernel_size)
            nn.init.xavier_normal_(W, gain=nn.init.calculate_gain('relu'))
            self.register_parameter('W', nn.Parameter(W.view(dim_out, dim_in * kernel_size)))
            self.register_parameter('bias', nn.Parameter((torch.zeros(dim_out) + 1e-3).view(1, dim_out, 1)))

    def forward(self, x):
        bs, c, k, npt, na = x.shape
        rows = x.permute(0, 3, 4, 1, 2).reshape(bs * npt * na, c * k)          # column = c*K + k, like W
        bias = self.bias.view(-1) if self.bias is not None else None
        out = _ops.LinearFn.apply(rows, self.W, bias)
        return out.view(bs, npt, na, self.dim_out).permute(0, 3, 1, 2)


class IntraZPConv(nn.Module):
    """[b,c1,p,a_in] -> [b,c1,k,p,a_out] -> [b,c2,p,a_out]: angular-kNN anchor neighbourhood with linear angular
    kernel weights, then BasicZPConv (reference :61-98)."""

    def __init__(self, dim_in, dim_out, kernel_size, aperture, sigma, anchor_nn, anchor_in, anchor_out=None):
        super().__init__()
        if anchor_out is None:
            anchor_out = anchor_in
        anchor_in = L.get_anchors(anchor_in)
        anchor_out = L.get_anchors(anchor_out)
        kernels = L.get_intra_kernels(aperture, kernel_size)
        self.dim_in, self.dim_out = dim_in, dim_out
        self.kernel_size = kernels.shape[0]
        self.basic_conv = BasicZPConv(dim_in, dim_out, self.kernel_size)
        self.aperture, self.sigma, self.anchor_nn = aperture, sigma, anchor_nn
        intra_idx, intra_w = L.get_intra_kernel_weights(anchor_in, anchor_out, kernels, anchor_nn, aperture, sigma)
        self.register_buffer('anchor_out', anchor_out)
        self.register_buffer('kernels', kernels)
        self.register_buffer('intra_idx', intra_idx)
        self.register_buffer('intra_w', intra_w)

    def forward(self, x):
        feats = L.intra_zpconv_grouping_naive(self.intra_idx, self.intra_w, x.feats)
        feats = self.basic_conv(feats)
        return SphericalPointCloud(x.xyz, feats, self.anchor_out)


class _OutOfScope(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError(f"{type(self).__name__}: this legacy S^2 ZPConv module needs the reference's dead "
                                  "anchor_query path and is outside the B200 hot-path scope (no shipped model builds it)")


class InterZPConv(_OutOfScope):
    pass


class AnchorProp(_OutOfScope):
    pass
