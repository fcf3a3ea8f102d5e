"""Constants of the SO(3) convolution: 60 icosahedral rotation anchors, the 60x12 intra-anchor
neighbour table and the 24 kernel points.

The arrays in data/so3_constants.npz are produced by running the reference's own generator
(vgtk/vgtk/functional/rotation.py:236-343 on data/anchors/sphere12.ply, and
data/anchors/kpsphere24.ply) through tests/golden/make_golden.py, so they are bit-identical to
what `import vgtk.so3conv` computes at import time in the reference
(vgtk/vgtk/so3conv/functional.py:2630-2638).  `derive_anchor_group()` re-derives the same group
from the icosahedron geometry in float64 and is checked against the table in the tests.
"""
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "so3_constants.npz")
_cache = {}


def _load():
    if not _cache:
        with np.load(_DATA) as z:
            for k in z.files:
                _cache[k] = z[k]
    return _cache


def anchors_all():
    return _load()["anchors"]          # float32 [60,3,3], anchors[29] == I


def intra_idx():
    return _load()["intra_idx"]        # int64 [60,12], column 9 is the identity


def kernel_points_base(kernel_size=1):
    if kernel_size != 1:
        raise NotImplementedError("only kernel_size=1 (kpsphere24) is used by the shipped configurations")
    return _load()["kpsphere24"]       # float32 [24,3]


def select_anchor(anchors, k):
    """vgtk/vgtk/so3conv/functional.py:2641-2649."""
    if k == 1:
        return anchors[29][None]
    if k == 20:
        return anchors[::3]
    if k == 40:
        return anchors.reshape(20, 3, 3, 3)[:, :2].reshape(-1, 3, 3)
    return anchors


def get_anchors(k=60):
    return select_anchor(anchors_all(), k)


def scaled_kernel_points(radius, kernel_size=1):
    """vgtk/vgtk/so3conv/functional.py:111-121: points scaled so the largest norm equals `radius`."""
    pc = kernel_points_base(kernel_size).astype(np.float32)
    r = np.sqrt((pc ** 2).sum(1).max())
    return pc * radius / r


def derive_anchor_group():
    """Re-derive the rotation group from the icosahedron (float64), independent of the table.

    For every face f with outward unit normal n_f the reference builds R = Rx(g) Ry(beta) Rz(alpha)
    with (alpha, beta) the azimuth / elevation of n_f and three in-plane angles g = -2*pi*j/3
    (shifted by 60 degrees on the two 'odd' latitude rings), then re-bases the set so that
    element 29 is the identity (rotation.py:141-219,257).  Returns float64 [60,3,3]."""
    c = _load()
    v, f = c["ico_vertices"].astype(np.float64), c["ico_faces"]
    tri = v[f]
    nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    out = []
    for n in nrm:
        sb = n[2]
        cb = np.sqrt(1 - sb * sb)
        ca, sa = n[0] / cb, n[1] / cb
        Rz = np.array([[ca, sa, 0], [-sa, ca, 0], [0, 0, 1]])
        Ry = np.array([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]])
        shift = np.pi / 3 if (abs(sb + 0.19) < 0.01 or abs(sb - 0.79) < 0.01) else 0.0
        for j in range(3):
            g = -2 * np.pi * j / 3 + shift
            Rx = np.array([[1, 0, 0], [0, np.cos(g), np.sin(g)], [0, -np.sin(g), np.cos(g)]])
            out.append(Rx @ Ry @ Rz)
    Rs = np.stack(out)
    return np.einsum('bij,kj->bik', Rs, Rs[29])


def derive_intra_idx(Rs, row0):
    """The neighbour table is the left translate of its first row:
        intra_idx[a,k] = index of  R_a R_0^-1 R_{row0[k]}
    (the algebra behind rotation.py:263-300, where row0 comes from the face adjacency of the
    icosahedron around anchor 0 plus its own three in-plane rotations)."""
    step = np.einsum('ji,kjm->kim', Rs[0], Rs[row0])                  # R_0^T R_row0[k]
    cand = np.einsum('aij,kjm->akim', Rs, step)                       # R_a R_0^T R_row0[k]
    score = np.einsum('akim,cim->akc', cand, Rs)                      # <cand, R_c>
    return score.argmax(-1)
