"""vgtk.cuda.grouping (reference: vgtk/vgtk/cuda/grouping_cuda.cpp:176-181)."""
import torch

from equi_articulated_pose_b200 import ops as _ops


def _check(x, name):
    # the reference's CHECK_INPUT: CUDA + contiguous, else RuntimeError (grouping_cuda.cpp:66-68)
    if not x.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not x.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def ball_query(new_xyz, xyz, radius, nsample):
    _check(new_xyz, "new_xyz"); _check(xyz, "xyz")
    return _ops.ball_query(new_xyz, xyz, radius, nsample)


def furthest_point_sampling(source_xyz, m):
    _check(source_xyz, "source_xyz")
    return _ops.furthest_point_sampling(source_xyz, m)


def anchor_query(*args, **kwargs):
    raise NotImplementedError("anchor_query is dead code in the reference (call commented out at "
                              "vgtk/vgtk/spconv/functional.py:556-570); not part of the hot path")


def initial_anchor_query(*args, **kwargs):
    raise NotImplementedError("initial_anchor_query is only used by KernelPropagation, which no shipped model builds")
