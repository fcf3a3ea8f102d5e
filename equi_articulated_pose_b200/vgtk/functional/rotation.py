"""vgtk.functional.rotation -- rotation helpers used by the model heads
(reference: vgtk/vgtk/functional/rotation.py:379-519).  The anchor generator of the reference
(`icosahedron_so3_trimesh`, :236-343) is replaced by the committed constant table, see
equi_articulated_pose_b200/so3_constants.py."""
import numpy as np
import torch
import torch.nn.functional as F

from equi_articulated_pose_b200 import so3_constants as _C


def icosahedron_so3_trimesh(mesh_path=None, gsize=3, use_quats=False):
    if use_quats or gsize != 3:
        raise NotImplementedError("only the 60-element rotation-matrix table is provided")
    return _C.anchors_all(), _C.intra_idx(), None


def rotation_distance_np(r0, r1):
    """Closeness of r0 to every anchor in r1 [n,3,3] measured by trace(r1_i^T r0)."""
    if r0.ndim == 3:
        tr = np.einsum('nij,bij->bn', r1, r0)
        return tr.astype(np.int32), tr.astype(np.int32).argmax(1).astype(np.int32)
    rel = np.einsum('nji,jk->nik', r1, r0)
    tr = np.trace(rel, axis1=1, axis2=2)
    return tr, int(tr.argmax()), rel


def compute_rotation_matrix_from_quaternion(quaternion):
    """[b,4] (w,x,y,z) -> [b,3,3]"""
    q = F.normalize(quaternion, dim=1)
    qw, qx, qy, qz = q[:, 0:1], q[:, 1:2], q[:, 2:3], q[:, 3:4]
    xx, yy, zz = qx * qx, qy * qy, qz * qz
    xy, xz, yz = qx * qy, qx * qz, qy * qz
    xw, yw, zw = qx * qw, qy * qw, qz * qw
    row0 = torch.cat((1 - 2 * yy - 2 * zz, 2 * xy - 2 * zw, 2 * xz + 2 * yw), 1)
    row1 = torch.cat((2 * xy + 2 * zw, 1 - 2 * xx - 2 * zz, 2 * yz - 2 * xw), 1)
    row2 = torch.cat((2 * xz - 2 * yw, 2 * yz + 2 * xw, 1 - 2 * xx - 2 * yy), 1)
    return torch.stack((row0, row1, row2), 1)


def compute_rotation_matrix_from_ortho6d(ortho6d):
    """[b,6] -> [b,3,3] by Gram-Schmidt on the two 3-vectors (columns x, y, z)."""
    x = F.normalize(ortho6d[:, 0:3], dim=1)
    z = F.normalize(torch.cross(x, ortho6d[:, 3:6], dim=1), dim=1)
    y = torch.cross(z, x, dim=1)
    return torch.stack((x, y, z), 2)


def so3_mean(Rs, weights=None):
    """Chordal L2 mean of rotations Rs [b,n,3,3] (optional weights [b,n]): project the (weighted)
    sum onto SO(3) with an SVD, flipping the last singular direction when det < 0."""
    acc = Rs.sum(1) if weights is None else (weights[:, :, None, None] * Rs).sum(1)
    u, _, v = torch.svd(acc)
    vt = v.transpose(1, 2)
    sign = torch.det(u @ vt)
    d = torch.diag_embed(torch.stack((torch.ones_like(sign), torch.ones_like(sign), sign), 1))
    return u @ d @ vt
