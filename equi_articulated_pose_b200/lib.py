"""ctypes binding of libvgtkb200.so (the C ABI declared in include/vgtkb.h).

There is NO CPU fallback: if the library is missing, or a tensor is not a contiguous CUDA
tensor, the call raises.  PyTorch is used for device memory and streams only.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvgtkb200.so")
ABI_VERSION = 13

_lib = None
_device_ok = set()

c_int, c_i64, c_f32, c_vp = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p

# name -> argtypes; every entry point returns int (0 = ok).  Must list every symbol of vgtkb.h.
SIGNATURES = {
    "vgtkb_ball_query": [c_int, c_int, c_int, c_f32, c_int, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_furthest_point_sampling": [c_int, c_int, c_int, c_vp, c_vp, c_vp],
    "vgtkb_fps_plain": [c_int, c_int, c_int, c_vp, c_vp, c_vp],
    "vgtkb_gather_points_forward": [c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_gather_points_backward": [c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_chamfer_forward": [c_int, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_chamfer_backward": [c_int, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_anchor_chamfer_forward": [c_int, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_anchor_chamfer_backward": [c_int, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_inter_weights": [c_int] * 6 + [c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp],
    "vgtkb_inter_group_forward": [c_int] * 7 + [c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_int, c_vp],
    "vgtkb_inter_group_backward": [c_int] * 7 + [c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_int, c_vp],
    "vgtkb_intra_group_forward": [c_i64, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_intra_group_backward": [c_i64, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_pose_neighbourhood": [c_int] * 4 + [c_vp] * 7,
    "vgtkb_inter_pose_group_forward": [c_int] * 6 + [c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp],
    "vgtkb_inter_pose_group_backward": [c_int] * 6 + [c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp],
    "vgtkb_pose_neighbourhood_strided": [c_int] * 5 + [c_vp] * 9,
    "vgtkb_inter_pose_group_forward_strided": [c_int] * 7 + [c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp],
    "vgtkb_inter_pose_group_backward_strided": [c_int] * 7 + [c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp],
    "vgtkb_inter_zpconv_forward": [c_int] * 7 + [c_vp, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_inter_zpconv_backward": [c_int] * 7 + [c_vp, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_intra_zpconv_forward": [c_int] * 7 + [c_vp, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_intra_zpconv_backward": [c_int] * 7 + [c_vp, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_row_gather_forward": [c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_row_gather_backward": [c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_gemm_nt": [c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp],
    "vgtkb_gemm_tn": [c_int, c_int, c_i64, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp],
    "vgtkb_gather_gemm_nt": [c_i64, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp],
    "vgtkb_gather_gemm_tn": [c_i64, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp],
    "vgtkb_norm_stats": [c_int, c_i64, c_int, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp, c_f32, c_vp],
    "vgtkb_norm_sums": [c_int, c_i64, c_int, c_vp, c_vp, c_vp],
    "vgtkb_norm_finalize": [c_int, c_i64, c_int, c_f32, c_vp, c_vp, c_vp, c_vp, c_f32, c_vp],
    "vgtkb_norm_bwd_sums": [c_int, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_norm_bwd_apply": [c_int, c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_norm_act_forward": [c_int, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp],
    "vgtkb_norm_act_backward": [c_int, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_col_sum": [c_i64, c_int, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_pointnet_pool_forward": [c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_pointnet_embed_xyz": [c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_pointnet_pool_backward": [c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_split_bf16": [c_i64, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_gemm_nt_presplit": [c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_gemm_tn_presplit": [c_int, c_int, c_i64, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp],
    "vgtkb_inter_conv_forward": [c_int] * 8 + [c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp],
    "vgtkb_inter_conv_backward": [c_int] * 8 + [c_vp, c_vp, c_vp, c_vp, c_f32] + [c_vp] * 10 + [c_int, c_vp],
    "vgtkb_gemm_tn_planes": [c_int, c_int, c_i64] + [c_vp] * 7 + [c_int, c_vp, c_vp],
    "vgtkb_gather_gemm_nt_planes": [c_i64, c_int, c_int, c_int, c_int] + [c_vp] * 8,
    "vgtkb_gather_gemm_tn_planes": [c_i64, c_int, c_int, c_int, c_int] + [c_vp] * 7 + [c_int, c_vp, c_vp],
    "vgtkb_norm_act_forward_planes": [c_int, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_norm_bwd_apply_planes": [c_int, c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_norm_act_backward_planes": [c_int, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_knn_query": [c_int] * 4 + [c_vp] * 5,
    "vgtkb_sa_group_forward": [c_int] * 6 + [c_vp] * 6,
    "vgtkb_sa_group_backward": [c_int] * 6 + [c_vp] * 4,
    "vgtkb_sa_maxpool_forward": [c_i64, c_int, c_int, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp],
    "vgtkb_sa_maxpool_backward": [c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_vp],
    "vgtkb_three_nn": [c_int] * 3 + [c_vp] * 5,
    "vgtkb_three_interpolate_forward": [c_int] * 4 + [c_vp] * 5,
    "vgtkb_three_interpolate_backward": [c_int] * 4 + [c_vp] * 5,
    "vgtkb_peer_allreduce_f64": [c_int, c_vp, c_int, c_int, c_vp, ctypes.c_uint64, c_vp],
    "vgtkb_norm_finalize_peer": [c_int, c_f32, c_vp, c_vp, c_vp, c_vp, c_f32, c_int, c_int, c_vp, ctypes.c_uint64, c_vp],
    "vgtkb_peer_status": [c_vp, c_int, ctypes.POINTER(c_i64), c_vp],
    "vgtkb_weight_planes": [c_int, c_vp, c_int, c_vp],
}
# entry points without a stream argument (setup of the peer mailboxes); status int like the others
SETUP_SIGNATURES = {
    "vgtkb_peer_mailbox_bytes": [c_int, ctypes.POINTER(c_i64)],
    "vgtkb_peer_alloc": [c_i64, ctypes.POINTER(c_vp), c_vp],
    "vgtkb_peer_open": [c_vp, ctypes.POINTER(c_vp)],
    "vgtkb_peer_close": [c_vp],
    "vgtkb_peer_free": [c_vp],
}
# (vgtkb_peer_status takes a stream but returns through a host pointer: bound in SIGNATURES below)
NO_STATUS = {"vgtkb_last_error": (ctypes.c_char_p, []), "vgtkb_version": (c_int, []),
             "vgtkb_device_check": (c_int, []), "vgtkb_inter_conv_supported": (c_int, [c_int] * 8)}


class VgtkbError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VgtkbError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(nvcc, sm_100a). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in NO_STATUS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    for name, args in list(SIGNATURES.items()) + list(SETUP_SIGNATURES.items()):
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = c_int, args
    if lib.vgtkb_version() != ABI_VERSION:
        raise VgtkbError(f"ABI mismatch: library {lib.vgtkb_version()} vs binding {ABI_VERSION}")
    _lib = lib
    return lib


def _ensure_device(dev_index):
    if dev_index in _device_ok:
        return
    lib = load()
    with torch.cuda.device(dev_index):
        if lib.vgtkb_device_check() != 0:
            raise VgtkbError(lib.vgtkb_last_error().decode())
    _device_ok.add(dev_index)


def setup_call(name, device, *args):
    """Invoke a stream-less setup entry point with `device` current; raise on a non-zero status."""
    lib = load()
    idx = device.index if device.index is not None else torch.cuda.current_device()
    _ensure_device(idx)
    with torch.cuda.device(idx):
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise VgtkbError(f"{name} failed ({rc}): {lib.vgtkb_last_error().decode()}")


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise VgtkbError("expected a CUDA tensor (there is no CPU path)")
    if not t.is_contiguous():
        raise VgtkbError("expected a contiguous tensor")
    return t.data_ptr()


def call(name, device, *args):
    """Invoke an entry point on `device`'s current stream; raise on a non-zero status."""
    lib = load()
    if len(args) + 1 != len(SIGNATURES[name]):       # ctypes would pass surplus arguments through silently
        raise VgtkbError(f"{name}: {len(args)} arguments given, {len(SIGNATURES[name]) - 1} expected (+ stream)")
    _ensure_device(device.index if device.index is not None else torch.cuda.current_device())
    with torch.cuda.device(device):
        cur = torch.cuda.current_stream()
        if PROFILE is not None:
            # CUDA events on the launching stream around this entry point (bench.py roofline leg)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(cur)
            rc = getattr(lib, name)(*args, cur.cuda_stream)
            e1.record(cur)
            PROFILE.append((name, args, e0, e1))
        else:
            rc = getattr(lib, name)(*args, cur.cuda_stream)
    if rc != 0:
        raise VgtkbError(f"{name} failed ({rc}): {lib.vgtkb_last_error().decode()}")
    COUNTERS["launch_calls"] += 1
    kernels = KERNELS_PER_CALL.get(name, 1)
    w = PREPARED_WEIGHT_ARG.get(name)
    if w is not None and args[w[0]] is None:         # prepared weight planes: no transpose / split launches in this call
        kernels -= w[1]
    COUNTERS["kernels"] += kernels


COUNTERS = {"launch_calls": 0, "kernels": 0}
# entry point -> (index of its fp32 weight argument, kernels saved when it is NULL = planes prepared by vgtkb_weight_planes)
PREPARED_WEIGHT_ARG = {"vgtkb_gemm_nt": (4, 1), "vgtkb_gemm_nt_presplit": (5, 1), "vgtkb_gather_gemm_nt_planes": (8, 1),
                       "vgtkb_inter_conv_forward": (14, 1), "vgtkb_inter_conv_backward": (13, 2)}
PROFILE = None  # set to a list to record (entry point, args, start event, end event) per call
# device kernels launched per entry point (memsets not counted); used for bench.py's gpu_launches
KERNELS_PER_CALL = {"vgtkb_gemm_nt": 2, "vgtkb_gemm_tn": 2, "vgtkb_gather_gemm_nt": 2, "vgtkb_gather_gemm_tn": 2,
                    "vgtkb_chamfer_forward": 2, "vgtkb_chamfer_backward": 4, "vgtkb_anchor_chamfer_forward": 2, "vgtkb_anchor_chamfer_backward": 2, "vgtkb_norm_stats": 2,
                    "vgtkb_norm_act_backward": 2, "vgtkb_norm_bwd_sums": 2, "vgtkb_col_sum": 2,
                    "vgtkb_inter_conv_forward": 3, "vgtkb_inter_conv_backward": 6, "vgtkb_gemm_nt_presplit": 2,
                    "vgtkb_gemm_tn_presplit": 2, "vgtkb_gemm_tn_planes": 1, "vgtkb_gather_gemm_nt_planes": 2,
                    "vgtkb_gather_gemm_tn_planes": 1, "vgtkb_norm_act_backward_planes": 2}
