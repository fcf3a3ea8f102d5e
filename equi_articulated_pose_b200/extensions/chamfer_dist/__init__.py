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
