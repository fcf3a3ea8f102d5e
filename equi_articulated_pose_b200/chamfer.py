"""Top-level `chamfer` module of the reference (extensions/chamfer_dist/chamfer_cuda.cpp:36-39)."""
from equi_articulated_pose_b200 import ops as _ops


def forward(xyz1, xyz2):
    return _ops.chamfer_forward(xyz1, xyz2)


def backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2):
    return _ops.chamfer_backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2)
