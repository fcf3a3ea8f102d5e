"""extensions.chamfer_dist (reference: extensions/chamfer_dist/__init__.py:13-45)."""
import torch

import chamfer


class ChamferFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        dist1, dist2, idx1, idx2 = chamfer.forward(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        return dist1, dist2

    @staticmethod
    def backward(ctx, grad_dist1, grad_dist2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        return chamfer.backward(xyz1, xyz2, idx1, idx2, grad_dist1.contiguous(), grad_dist2.contiguous())


class ChamferDistance(torch.nn.Module):
    def __init__(self, ignore_zeros=False):
        super().__init__()
        self.ignore_zeros = ignore_zeros

    def forward(self, xyz1, xyz2, return_raw=False):
        if xyz1.size(0) == 1 and self.ignore_zeros:
            xyz1 = xyz1[torch.sum(xyz1, dim=2).ne(0)].unsqueeze(dim=0)
            xyz2 = xyz2[torch.sum(xyz2, dim=2).ne(0)].unsqueeze(dim=0)
        dist1, dist2 = ChamferFunction.apply(xyz1, xyz2)
        if return_raw:
            return dist1, dist2
        return torch.mean(dist1) + torch.mean(dist2)


class AnchorChamferDistance(torch.nn.Module):
    """Fused form of the anchor-orbit reconstruction loss of model 38
    (SPConvNets/models/unsup_seg_so3_pose_conv_pn_38_multi_stage.py:429-450): the reference transforms the reconstruction
    by every anchor pose ([B,A,M,3]), repeats the input cloud A times and calls ChamferDistance on [B*A,...]; this module
    takes the same four tensors and returns the same per-anchor chamfer means, the minimum over the orbit and its index.

        canon [B,3,M] or [B,M,3], rot [B,A,3,3], trans [B,A,3], ori [B,3,N] or [B,N,3]
        -> (minn_chamfer [B], orbit [B] int64, chamfer_recon_to_ori [B,A], chamfer_ori_to_recon [B,A])
    """

    def __init__(self, glb_single_cd=0):
        super().__init__()
        self.glb_single_cd = glb_single_cd

    def forward(self, canon, rot, trans, ori):
        from equi_articulated_pose_b200 import ops
        if canon.shape[-1] != 3:
            canon = canon.transpose(1, 2)
        if ori.shape[-1] != 3:
            ori = ori.transpose(1, 2)
        d1, d2, _, _ = ops.anchor_chamfer(canon.contiguous(), rot.contiguous(), trans.contiguous(), ori.contiguous())
        cd_r2o, cd_o2r = d1.mean(-1), d2.mean(-1)
        total = cd_o2r if self.glb_single_cd == 1 else cd_r2o + cd_o2r
        minn, orbit = torch.min(total, dim=-1)
        return minn, orbit, cd_r2o, cd_o2r
