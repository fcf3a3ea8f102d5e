"""vgtk.cuda.gathering (reference: vgtk/vgtk/cuda/gathering_cuda.cpp:62-65)."""
from equi_articulated_pose_b200 import ops as _ops


def _check(x, name):
    if not x.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not x.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def gather_points_forward(support_points, grouped_indices):
    _check(support_points, "support_points"); _check(grouped_indices, "grouped_indices")
    return _ops.gather_points_forward(support_points, grouped_indices)


def gather_points_backward(grad_out, grouped_indices, npoint):
    _check(grad_out, "grad_out"); _check(grouped_indices, "grouped_indices")
    return _ops.gather_points_backward(grad_out, grouped_indices, npoint)
