"""`torch_cluster` shim: only `fps`, backed by the sm_100a FPS kernel (vgtkb_fps_plain).

The reference calls torch_cluster.fps (torch-cluster==1.5.9, env.yaml:13) from
SPConvNets/models/model_util.py:183-200 (`farthest_point_sampling`), models/utils.py:80-97,
unsup_seg_so3_pose_conv_pn_38_multi_stage.py:1739 and datasets/MotionDataset.py:630-631, always with
`random_start=False` on a batch vector made of equal-length, contiguous segments.  The package is not installed in
this image and is not part of /root/reference (its tie-break is unpinned); `equi_articulated_pose_b200.install()`
appends this directory to the END of sys.path, so a real torch_cluster installation wins when there is one.
"""
import math

import torch


def fps(src, batch=None, ratio=0.5, random_start=True):
    """src [sum N_b, D>=3 (first three = xyz)], batch [sum N_b] sorted -> flat global indices (int64) of the
    ceil(ratio * N_b) samples of every segment, segment after segment, in sampling order."""
    from equi_articulated_pose_b200 import ops
    if random_start:
        raise NotImplementedError("torch_cluster shim: only random_start=False (what the reference passes)")
    if isinstance(ratio, torch.Tensor):
        ratio = float(ratio.flatten()[0])
    n_total = src.shape[0]
    if batch is None:
        nb, n = 1, n_total
    else:
        nb = int(batch[-1]) + 1 if n_total > 0 else 0
        n = n_total // max(nb, 1)
        if nb * n != n_total or not bool((batch.view(nb, n) == torch.arange(nb, device=batch.device).view(nb, 1)).all()):
            raise NotImplementedError("torch_cluster shim: segments must be contiguous and of equal length")
    m = int(math.ceil(ratio * n))
    xyz = src[:, :3].float().reshape(nb, n, 3).permute(0, 2, 1).contiguous()
    idx = ops.fps_plain(xyz, m).long()
    return (idx + torch.arange(nb, device=idx.device).view(nb, 1) * n).reshape(-1)
