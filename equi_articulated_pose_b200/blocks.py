"""Block builders of the equivariant backbone, fused for the B200 kernels.

Mirror of SPConvNets/utils/base_so3conv.py:21-221 (preprocess_input, IntraSO3ConvBlock,
InterSO3ConvBlock, BasicSO3ConvBlock, SeparableSO3ConvBlock) and of the backbone part of
SPConvNets/models/cls_so3net_pn.py:43-168 (build_model): same constructor arguments, same
parameter dictionaries, same module / parameter names (so a reference state_dict loads), same
returned tuples.  The difference is the execution: features stay channels-last, the grouping
runs in the fused sm_100a kernels, the contractions on tcgen05, and BatchNorm / InstanceNorm +
leaky_relu (+ the residual add) are single fused passes instead of separate torch ops.

The reference's own, unmodified base_so3conv.py also runs on top of the `vgtk` modules of this
package (module-level drop-in); this file is the faster block-level path the benchmark uses.
"""
import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

import equi_articulated_pose_b200 as _pkg

_pkg.install()
import vgtk.spconv as zptk  # noqa: E402
import vgtk.so3conv as sptk  # noqa: E402
from . import ops as _ops  # noqa: E402


def preprocess_input(x, na, add_center=True):
    """[nb,np,3] -> SphericalPointCloud(xyz [nb,3,np], ones [nb,1,np,na])  (base_so3conv.py:21-28)."""
    has_normals = x.shape[2] == 6
    if add_center and not has_normals:
        center = x.mean(1, keepdim=True)
        x = torch.cat((center, x), dim=1)[:, :-1]
    xyz = x[:, :, :3]
    return zptk.SphericalPointCloud(xyz.permute(0, 2, 1).contiguous(), sptk.get_occupancy_features(x, na, add_center), None)


def _slope(activation):
    if activation in ('leaky_relu',):
        return 0.01
    if activation in ('relu',):
        return 0.0
    if activation is None:
        return 1.0
    raise NotImplementedError(f"activation {activation}")


def _rows(feats):
    """logical [B,C,P,A] -> ([B*P*A, C] channels-last rows, (B,P,A,C))."""
    b, c, p, a = feats.shape
    return feats.permute(0, 2, 3, 1).reshape(b * p * a, c), (b, p, a, c)


def _unrows(rows, b, p, a):
    return rows.view(b, p, a, rows.shape[-1]).permute(0, 3, 1, 2)


class FusedBatchNorm2d(nn.BatchNorm2d):
    """nn.BatchNorm2d (identical parameters / buffers) whose forward on channels-last rows is one
    statistics pass + one normalise-activate pass (vgtkb_norm_stats / vgtkb_norm_act_forward).
    `sync_group` (set by convert_sync_batchnorm): statistics over all data-parallel ranks, nn.SyncBatchNorm semantics."""

    sync_group = None
    planes_fwd = False      # the result feeds a tensor-core contraction: also write its bf16 operand planes
    planes_bwd = False      # the input comes from a contraction: write the planes of its gradient in backward
    counter_managed = False # num_batches_tracked is a view of the owning model's flat counter, bumped once per step there

    def _step_momentum(self, training):
        """Bump the batch counter (unless the owning model does it) and return the running-average factor of this step."""
        if training and self.track_running_stats and self.num_batches_tracked is not None and not self.counter_managed:
            self.num_batches_tracked += 1
        if self.momentum is None:       # torch: cumulative moving average when momentum is None
            return 1.0 / max(float(self.num_batches_tracked), 1.0) if training and self.num_batches_tracked is not None else 0.0
        return self.momentum

    @staticmethod
    def pair_ok(n1, n2):
        """Both are training-mode SyncBatchNorm layers over the same ranks and channel count: their statistics can share one
        exchange per direction (ops.SyncNormPairFn).  VGTKB_SYNCBN_PAIR=0 keeps one exchange per norm (A/B runs)."""
        return (isinstance(n1, FusedBatchNorm2d) and isinstance(n2, FusedBatchNorm2d) and n1.training and n2.training
                and n1.sync_group is not None and n1.sync_group is n2.sync_group and _ops._sync_world(n1.sync_group) > 1
                and n1.num_features == n2.num_features and n1.eps == n2.eps and n1.affine == n2.affine
                and os.environ.get("VGTKB_SYNCBN_PAIR", "1") != "0")

    @staticmethod
    def forward_rows_pair(n1, rows1, slope1, n2, rows2, slope2):
        """(leaky_relu(n1(rows1)), leaky_relu(n2(rows2))) with one statistics exchange for both (pair_ok must hold)."""
        m1, m2 = n1._step_momentum(True), n2._step_momentum(True)
        y1, y2 = _ops.SyncNormPairFn.apply(
            rows1.unsqueeze(0), n1.weight, n1.bias, n1.running_mean, n1.running_var, m1, slope1, n1.planes_fwd, n1.planes_bwd,
            rows2.unsqueeze(0), n2.weight, n2.bias, n2.running_mean, n2.running_var, m2, slope2, n2.planes_fwd, n2.planes_bwd,
            n1.eps, n1.sync_group)
        return y1.squeeze(0), y2.squeeze(0)

    def forward_rows(self, rows, slope, residual=None):
        training = self.training or not self.track_running_stats
        mom = self._step_momentum(training)
        return _ops.norm_act(rows.unsqueeze(0), self.weight, self.bias, None if residual is None else residual.unsqueeze(0),
                             self.running_mean if training else self.running_mean,
                             self.running_var if training else self.running_var,
                             mom, self.eps, slope, use_running=not training, sync_group=self.sync_group,
                             planes_fwd=self.planes_fwd, planes_bwd=self.planes_bwd and training).squeeze(0)


class FusedInstanceNorm2d(nn.InstanceNorm2d):
    """nn.InstanceNorm2d(affine=False): per-(sample, channel) statistics over (points, anchors)."""

    planes_fwd = False
    planes_bwd = False

    def forward_rows(self, rows, batch, slope, residual=None):
        m, c = rows.shape
        res = None if residual is None else residual.view(batch, m // batch, c)
        return _ops.norm_act(rows.view(batch, m // batch, c), None, None, res, None, None, 0.1, self.eps, slope,
                             planes_fwd=self.planes_fwd, planes_bwd=self.planes_bwd).view(m, c)


def convert_sync_batchnorm(module, process_group=None, peer_memory=True):
    """Counterpart of nn.SyncBatchNorm.convert_sync_batchnorm (the reference trainer applies it to the whole model,
    SPConvNets/trainer_unsup_arti_align.py:430): every fused BatchNorm of `module` takes its statistics over all ranks of
    `process_group` (None = the default group).  InstanceNorm layers are per-sample and stay local.
    peer_memory: exchange the sums through NVLink peer mailboxes (one node; csrc/peer.cu) when they can be set up on
    every rank, NCCL all-reduces otherwise.  Collective: call it on all ranks."""
    import torch.distributed as dist
    for m in module.modules():
        if isinstance(m, FusedBatchNorm2d):
            m.sync_group = True if process_group is None else process_group
    if peer_memory and dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1 \
            and dist.get_backend(process_group) == "nccl" and _ops.peer_mailbox_for(process_group) is None \
            and os.environ.get("VGTKB_PEER_SYNCBN", "1") != "0":        # VGTKB_PEER_SYNCBN=0: NCCL all-reduces instead
        from . import dataparallel as _dp
        dev = next(module.parameters()).device
        try:
            _ops.set_peer_mailbox(_dp.PeerMailbox(dev, process_group), process_group)   # one mailbox per process group
        except RuntimeError:
            _ops.set_peer_mailbox(None, process_group)   # every rank raised (agreed outcome): NCCL path
    return module


def _make_norm(norm, dim):
    if norm is None:
        return FusedInstanceNorm2d(dim, affine=False)
    if norm == 'BatchNorm2d':
        return FusedBatchNorm2d(dim)
    raise NotImplementedError(f"norm {norm}")


def _apply_norm(norm, rows, batch, slope, residual=None):
    if isinstance(norm, FusedBatchNorm2d):
        return norm.forward_rows(rows, slope, residual)
    return norm.forward_rows(rows, batch, slope, residual)


class IntraSO3ConvBlock(nn.Module):
    """base_so3conv.py:37-67"""

    def __init__(self, dim_in, dim_out, norm=None, activation='relu', dropout_rate=0):
        super().__init__()
        self.conv = sptk.IntraSO3Conv(dim_in, dim_out)
        self.norm = _make_norm(norm, dim_out)
        self.norm.planes_bwd = dim_in % 64 == 0 and dim_out % 64 == 0    # the gradient of the conv output feeds two contractions
        self.slope = _slope(activation)
        self.dropout = nn.Dropout(dropout_rate) if dropout_rate > 0 else None

    def forward(self, x, residual_rows=None):
        y = self.conv(x)
        rows, (b, p, a, _) = _rows(y.feats)
        if self.training and self.dropout is not None:
            # the reference drops out relu(norm(intra)) and only THEN adds the skip branch (base_so3conv.py:59-64,210-217):
            # the residual must not ride through the fused pass, or dropout would zero / rescale the skip path too
            rows = self.dropout(_apply_norm(self.norm, rows, b, self.slope))
            if residual_rows is not None:
                rows = rows + residual_rows
        else:
            rows = _apply_norm(self.norm, rows, b, self.slope, residual_rows)
        return zptk.SphericalPointCloud(y.xyz, _unrows(rows, b, p, a), y.anchors)


class InterSO3ConvBlock(nn.Module):
    """base_so3conv.py:93-132"""

    def __init__(self, dim_in, dim_out, kernel_size, stride, radius, sigma, n_neighbor, multiplier, kanchor=60,
                 lazy_sample=None, norm=None, activation='relu', pooling='none', dropout_rate=0):
        super().__init__()
        if lazy_sample is None:
            lazy_sample = True
        pooling_method = None if pooling == 'none' else pooling
        self.conv = sptk.InterSO3Conv(dim_in, dim_out, kernel_size, stride, radius, sigma, n_neighbor,
                                      kanchor=kanchor, lazy_sample=lazy_sample, pooling=pooling_method)
        self.norm = _make_norm(norm, dim_out)
        self.norm.planes_bwd = dim_in % 32 == 0 and dim_out % 64 == 0     # consumed by vgtkb_inter_conv_backward
        self.slope = _slope(activation)
        self.dropout = nn.Dropout(dropout_rate) if dropout_rate > 0 else None

    def forward(self, x, inter_idx=None, inter_w=None):
        inter_idx, inter_w, sample_idx, y = self.conv(x, inter_idx, inter_w)
        rows, (b, p, a, _) = _rows(y.feats)
        rows = _apply_norm(self.norm, rows, b, self.slope)
        if self.training and self.dropout is not None:
            rows = self.dropout(rows)
        return inter_idx, inter_w, sample_idx, zptk.SphericalPointCloud(y.xyz, _unrows(rows, b, p, a), y.anchors)


class SeparableSO3ConvBlock(nn.Module):
    """inter conv -> intra conv, plus a 1x1-conv skip branch (base_so3conv.py:174-221)."""

    def __init__(self, params):
        super().__init__()
        dim_in, dim_out = params['dim_in'], params['dim_out']
        self.use_intra = params['kanchor'] > 1
        self.inter_conv = InterSO3ConvBlock(**params)
        if self.use_intra:
            self.intra_conv = IntraSO3ConvBlock(dim_in=dim_out, dim_out=dim_out, dropout_rate=params['dropout_rate'],
                                                activation=params['activation'])
            # operand planes: the inter norm's result feeds the intra gather-GEMM
            self.inter_conv.norm.planes_fwd = dim_out % 64 == 0
        self.stride = params['stride']
        self.skip_conv = nn.Conv2d(dim_in, dim_out, 1)
        self.norm = _make_norm(params.get('norm'), dim_out)
        self.slope = _slope(params['activation'])

    def forward(self, x, inter_idx, inter_w):
        skip = x.feats
        b, ci, n, a = skip.shape
        # the block input feeds the inter conv and the skip branch: in backward the skip branch (recorded later, run first)
        # deposits its input gradient in the slot and the inter conv's scatter adds onto it (ops.GradSlot)
        slot = None
        if self.training and skip.requires_grad and skip.is_cuda and self.inter_conv.conv.pooling is None and os.environ.get("VGTKB_GRAD_SLOT", "1") != "0":
            slot = _ops.GradSlot((b, n, a, ci))
            self.inter_conv.conv._grad_slot = slot
        # SyncBatchNorm: the inter conv's norm and the skip branch's norm share ONE statistics exchange per direction
        pair = (self.use_intra and self.inter_conv.dropout is None and skip.is_cuda
                and FusedBatchNorm2d.pair_ok(self.inter_conv.norm, self.norm))
        if pair:
            inter_idx, inter_w, sample_idx, y = self.inter_conv.conv(x, inter_idx, inter_w)      # norm applied below
        else:
            inter_idx, inter_w, sample_idx, y = self.inter_conv(x, inter_idx, inter_w)
        self.inter_conv.conv._grad_slot = None
        if slot is not None and not slot.armed:
            slot = None                             # the conv did not take the fused path: ordinary autograd accumulation
        # skip branch first, so that its result rides along as the residual of the last fused pass
        srows = skip.permute(0, 2, 3, 1)
        if self.stride > 1:
            srows = _ops.RowGatherFn.apply(srows.reshape(b, n, a * ci), sample_idx.to(torch.int32).contiguous(), slot)
            lin_slot = None
        else:
            lin_slot = slot
        p = srows.shape[1]
        srows = srows.reshape(b * p * a, ci)
        w = self.skip_conv.weight.view(self.skip_conv.out_channels, ci)
        wp = getattr(self.skip_conv, '_wp', None)       # operand planes produced once per step (SO3Backbone)
        if wp is not None and not (srows.is_cuda and wp.param is self.skip_conv.weight and wp.valid()):
            wp = None
        srows = _ops.LinearFn.apply(srows, w, self.skip_conv.bias, None, lin_slot, wp)
        if pair:
            yrows, (yb, yp, ya, _) = _rows(y.feats)
            yrows, srows = FusedBatchNorm2d.forward_rows_pair(self.inter_conv.norm, yrows, self.inter_conv.slope,
                                                             self.norm, srows, self.slope)
            y = zptk.SphericalPointCloud(y.xyz, _unrows(yrows, yb, yp, ya), y.anchors)
        else:
            srows = _apply_norm(self.norm, srows, b, self.slope)
        if self.use_intra:
            y = self.intra_conv(y, residual_rows=srows)
            out = y.feats
        else:
            out = y.feats + _unrows(srows, b, p, a)
        return inter_idx, inter_w, sample_idx, zptk.SphericalPointCloud(y.xyz, out, y.anchors)

    def get_anchor(self):
        return torch.from_numpy(sptk.get_anchors())


class BasicSO3ConvBlock(nn.Module):
    """A list of inter / intra / separable layers threading (inter_idx, inter_w)  (base_so3conv.py:135-172)."""

    def __init__(self, params):
        super().__init__()
        self.blocks = nn.ModuleList()
        self.layer_types = []
        for param in params:
            if param['type'] == 'intra_block':
                conv = IntraSO3ConvBlock(**param['args'])
            elif param['type'] == 'inter_block':
                conv = InterSO3ConvBlock(**param['args'])
            elif param['type'] == 'separable_block':
                conv = SeparableSO3ConvBlock(param['args'])
            else:
                raise ValueError(f'No such type of SO3Conv {param["type"]}')
            self.layer_types.append(param['type'])
            self.blocks.append(conv)
        self.params = params

    def forward(self, x):
        inter_idx, inter_w = None, None
        for conv, param in zip(self.blocks, self.params):
            if param['type'] in ['inter', 'inter_block', 'separable_block']:
                inter_idx, inter_w, _, x = conv(x, inter_idx, inter_w)
                if param['args']['stride'] > 1:
                    inter_idx, inter_w = None, None
            elif param['type'] in ['intra_block']:
                x = conv(x)
            else:
                raise ValueError(f'No such type of SO3Conv {param["type"]}')
        return x

    def get_anchor(self):
        return torch.from_numpy(sptk.get_anchors())


def backbone_params(input_num=1024, kanchor=60, mlps=((64, 64), (128, 128), (256, 256), (256,)),
                    strides=(2, 2, 2, 2), initial_radius_ratio=0.2, sampling_ratio=0.4, sampling_density=0.5,
                    kernel_multiplier=2, input_radius=1.0, sigma_ratio=0.5, xyz_pooling=None, dropout_rate=0.0):
    """Per-layer argument dictionaries of the classic equivariant backbone
    (SPConvNets/models/cls_so3net_pn.py:43-150 with its default arguments)."""
    strides = list(strides)
    if input_num > 1024:
        sampling_ratio /= (input_num / 1024)
        strides[0] = int(2 * (input_num / 1024))
    n_layer = len(mlps)
    multipliers = [2 ** i for i in range(n_layer + 1)]
    num_centers = [int(input_num / m) for m in multipliers]
    radius_ratio = [initial_radius_ratio * m ** sampling_density for m in multipliers]
    radii = [r * input_radius for r in radius_ratio]
    weighted_sigma = [sigma_ratio * radii[0] ** 2]
    for i in range(len(strides)):
        weighted_sigma.append(weighted_sigma[i] * 2)
    out, dim_in = [], 1
    for i, block in enumerate(mlps):
        layers = []
        for j, dim_out in enumerate(block):
            lazy_sample = i != 0 or j != 0
            stride_conv = i == 0 or xyz_pooling != 'stride'
            neighbor = int(sampling_ratio * num_centers[i] * radius_ratio[i] ** (1 / sampling_density))
            if j == 0:
                inter_stride = strides[i]
                nidx = i if i == 0 else i + 1
                if stride_conv:
                    neighbor *= 2
            else:
                inter_stride, nidx = 1, i + 1
            layers.append({'type': 'inter_block' if kanchor < 60 else 'separable_block', 'args': {
                'dim_in': dim_in, 'dim_out': dim_out, 'kernel_size': 1, 'stride': inter_stride,
                'radius': radii[nidx], 'sigma': weighted_sigma[nidx], 'n_neighbor': neighbor,
                'lazy_sample': lazy_sample, 'dropout_rate': dropout_rate, 'multiplier': kernel_multiplier,
                'activation': 'leaky_relu', 'pooling': xyz_pooling, 'kanchor': kanchor, 'norm': 'BatchNorm2d'}})
            dim_in = dim_out
        out.append(layers)
    return out


def model38_backbone_params(input_num=512, kanchor=60, init_radius=0.2, mlps=((64,), (128,), (512,)), dropout_rate=0.0):
    """Per-layer argument dictionaries of the equivariant backbone model 38 (`--use-equi=38`) builds for `kanchor=60`
    (SPConvNets/models/unsup_seg_so3_pose_conv_pn_38_multi_stage.py:2084-2225): three stride-1 separable blocks
    64 / 128 / 512, 64 neighbours each, input radius 0.4, radii and sigmas taken from the stride ladder [1, 2, 4, 8]."""
    input_radius, sampling_density, sigma_ratio = 0.4, 0.5, 0.5
    strides = [2, 2, 2, 2]
    n_layer = len(mlps)
    mult = [1]
    for i in range(n_layer):
        mult.append(mult[-1] * strides[i])
    radius_ratio = [init_radius * m ** sampling_density for m in mult]
    radii = [r * input_radius for r in radius_ratio]
    weighted_sigma = [sigma_ratio * radii[0] ** 2]
    for i, st in enumerate(strides):
        weighted_sigma.append(weighted_sigma[i] * st)
    out, dim_in = [], 1
    for i, block in enumerate(mlps):
        layers = []
        for j, dim_out in enumerate(block):
            neighbor = 32
            nidx = i + 1
            if j == 0:
                nidx = i if i == 0 else i + 1
                neighbor *= 2
            layers.append({'type': 'inter_block' if kanchor < 60 else 'separable_block', 'args': {
                'dim_in': dim_in, 'dim_out': dim_out, 'kernel_size': 1, 'stride': 1,
                'radius': radii[nidx], 'sigma': weighted_sigma[nidx], 'n_neighbor': neighbor,
                'lazy_sample': i != 0 or j != 0, 'dropout_rate': dropout_rate, 'multiplier': 2,
                'activation': 'leaky_relu', 'pooling': None, 'kanchor': kanchor, 'norm': 'BatchNorm2d'}})
            dim_in = dim_out
        out.append(layers)
    return out


class SO3Backbone(nn.Module):
    """`ClsSO3ConvModel.backbone` + its forward loop (cls_so3net_pn.py:15-33), without the head."""

    def __init__(self, params, na=60):
        super().__init__()
        self.backbone = nn.ModuleList([BasicSO3ConvBlock(p) for p in params])
        self.na_in = na

    def _bump_counters(self):
        """All `num_batches_tracked` buffers of the fused BatchNorm layers live in ONE int64 tensor (views; state-dict keys and
        values unchanged) and are incremented by one launch per step instead of one per layer (14 launches in the classic
        backbone).  Rebuilt when the module moved to another device."""
        bns = [m for m in self.modules() if isinstance(m, FusedBatchNorm2d) and m.track_running_stats
               and m.num_batches_tracked is not None]
        if not bns:
            return
        flat = getattr(self, '_nbt_flat', None)
        dev = bns[0].num_batches_tracked.device
        ok = flat is not None and flat.device == dev and flat.numel() == len(bns) and all(
            m.counter_managed and m.num_batches_tracked.data_ptr() == flat[i].data_ptr() for i, m in enumerate(bns))
        if not ok:
            flat = torch.stack([m.num_batches_tracked.detach().reshape(()) for m in bns]).to(dev)
            for i, m in enumerate(bns):
                m._buffers['num_batches_tracked'] = flat[i]
                m.counter_managed = True
            self._nbt_flat = flat
        flat += 1

    def _prepare_weight_planes(self):
        """bf16 operand planes of every conv weight of the backbone (forward and data-gradient orders) in ONE launch per
        step: ops.WeightPlanes.  The inter / intra / skip contractions then take their weight operand from the planes
        instead of a permute (or transpose) copy and a split launch each -- ~70 small launches per step of the classic
        backbone.  Only in the bf16 tensor-core modes; VGTKB_WEIGHT_PLANES=0 switches it off (A/B runs)."""
        wps = getattr(self, '_wps', None)
        if _ops.get_gemm_mode() not in (3, 4) or os.environ.get("VGTKB_WEIGHT_PLANES", "1") == "0":
            if wps is not None:
                for pw in wps[0].weights:
                    pw.version = -1                     # planes of an earlier call must not be picked up
            return
        params = [p for p in self.parameters()]
        if wps is None or wps[1] != [id(p) for p in params] or (params and params[0].device != wps[2]):
            planes = _ops.WeightPlanes()
            for m in self.modules():
                if isinstance(m, sptk.InterSO3Conv):
                    bc, role = m.basic_conv, "inter"
                elif isinstance(m, sptk.IntraSO3Conv):
                    bc, role = m.basic_conv, "intra"
                else:
                    bc = None
                if bc is not None and isinstance(getattr(bc, 'W', None), nn.Parameter):
                    ok = (bc.dim_in % 32 == 0 and bc.dim_out % 8 == 0) if role == "inter" else \
                         (bc.dim_in % 64 == 0 and bc.dim_out % 64 == 0)
                    bc._wp = planes.add(bc.W, bc.dim_out, bc.dim_in, bc.kernel_size, role) if ok else None
                if isinstance(m, SeparableSO3ConvBlock):
                    sc = m.skip_conv
                    ok = sc.in_channels % 8 == 0 and sc.out_channels % 8 == 0
                    sc._wp = planes.add(sc.weight, sc.out_channels, sc.in_channels, 1, "linear") if ok else None
            wps = (planes, [id(p) for p in params], params[0].device if params else None)
            self._wps = wps
        if params and params[0].is_cuda:
            wps[0].prepare()

    def forward(self, points):
        """points [B,N,3] -> SphericalPointCloud (xyz [B,3,P], feats logical [B,C,P,A])."""
        if self.training:
            self._bump_counters()
        if points.is_cuda:
            self._prepare_weight_planes()
        x = preprocess_input(points, self.na_in, False)
        for block in self.backbone:
            x = block(x)
        return x
