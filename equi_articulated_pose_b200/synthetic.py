"""Synthetic inputs and random-init weights for benchmarks and smoke runs (there is no dataset / checkpoint access).

Product-side twins of the generators the checker uses (oracle/so3.py keeps its own copies; tests/test_cabi.py asserts both
produce identical tensors, so the CUDA arm and the CPU baseline of bench.py see the same clouds and weights)."""
import math

import torch


def synthetic_cloud(batch, n, seed):
    """BASELINE config 2 'sphere-shell' clouds: unit directions x U(0.85, 1) -> [batch, n, 3]."""
    g = torch.Generator().manual_seed(seed)
    d = torch.randn(batch, n, 3, generator=g)
    d = d / d.norm(dim=2, keepdim=True)
    r = 0.85 + 0.15 * torch.rand(batch, n, 1, generator=g)
    return (d * r).float()


def init_backbone_state(params, seed=0, n_intra=12, n_kernel=24):
    """Random-init state dict with the reference's key names / shapes: xavier-normal W with relu gain
    (vgtk/vgtk/so3conv/modules.py:35-41), torch defaults for the 1x1 skip Conv2d, BatchNorm affine near (1, 0)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def xavier(co, ci, k):
        std = math.sqrt(2.0) * math.sqrt(2.0 / ((ci + co) * k))
        return (torch.randn(co, ci, k, generator=g) * std).view(co, ci * k)

    for bi, block in enumerate(params):
        for li, layer in enumerate(block):
            a = layer['args']
            ci, co = a['dim_in'], a['dim_out']
            pre = f'backbone.{bi}.blocks.{li}.'
            sep = layer['type'] == 'separable_block'
            ip = pre + ('inter_conv.' if sep else '')
            sd[ip + 'conv.basic_conv.W'] = xavier(co, ci, n_kernel)
            sd[ip + 'norm.weight'] = 1 + 0.1 * torch.randn(co, generator=g)
            sd[ip + 'norm.bias'] = 0.1 * torch.randn(co, generator=g)
            if sep:
                sd[pre + 'intra_conv.conv.basic_conv.W'] = xavier(co, co, n_intra)
                bound = 1 / math.sqrt(ci)
                sd[pre + 'skip_conv.weight'] = (torch.rand(co, ci, 1, 1, generator=g) * 2 - 1) * bound
                sd[pre + 'skip_conv.bias'] = (torch.rand(co, generator=g) * 2 - 1) * bound
                sd[pre + 'norm.weight'] = 1 + 0.1 * torch.randn(co, generator=g)
                sd[pre + 'norm.bias'] = 0.1 * torch.randn(co, generator=g)
    return sd


# ----------------------------------------------------------------------------- articulated objects (BASELINE configs 3 / 5)
def _box_surface(rs, n, size, thickness=None):
    """n points on the surface of an axis-aligned cuboid centred at the origin (area-weighted faces)."""
    import numpy as np
    sx, sy, sz = size
    areas = np.array([sy * sz, sy * sz, sx * sz, sx * sz, sx * sy, sx * sy])
    face = rs.choice(6, size=n, p=areas / areas.sum())
    u, v = rs.uniform(-0.5, 0.5, n), rs.uniform(-0.5, 0.5, n)
    pts = np.zeros((n, 3))
    for f in range(6):
        m = face == f
        axis, sign = f // 2, (1.0 if f % 2 == 0 else -1.0)
        other = [a for a in range(3) if a != axis]
        pts[m, axis] = sign * 0.5 * size[axis]
        pts[m, other[0]] = u[m] * size[other[0]]
        pts[m, other[1]] = v[m] * size[other[1]]
    return pts


def _fps_numpy(pts, m):
    import numpy as np
    idx = np.zeros(m, dtype=np.int64)
    d = np.full(len(pts), np.inf)
    for i in range(1, m):
        d = np.minimum(d, ((pts - pts[idx[i - 1]]) ** 2).sum(1))
        idx[i] = int(d.argmax())
    return pts[idx]


def articulated_cloud(kind, batch, n, seed):
    """Synthetic stand-ins for the reference's 'oven' / 'laptop' categories: two rigid
    parts joined by a revolute hinge, opened by theta drawn from the range the dataset uses
    (SPConvNets/datasets/MotionDataset.py:410-418: oven 45..125 degrees about a vertical edge, laptop -85..9 degrees
    about the rear edge), surface-sampled 4n then farthest-point-sampled to n (in FPS order, like the dataset, :630-631),
    centred and scaled to a unit bounding-box diagonal (:331-337), random global rotation.  -> float32 [batch, n, 3]"""
    import numpy as np
    from scipy.spatial.transform import Rotation
    rs = np.random.RandomState(seed)
    out = np.zeros((batch, n, 3), np.float32)
    for b in range(batch):
        if kind == "oven":
            body = _box_surface(rs, 3 * n, (0.6, 0.5, 0.5))
            door = _box_surface(rs, n, (0.6, 0.5, 0.03))
            theta = np.deg2rad(rs.uniform(45.0, 125.0))
            hinge, axis = np.array([-0.3, 0.0, 0.25]), np.array([0.0, 1.0, 0.0])         # vertical edge of the front face
            door = door + np.array([0.0, 0.0, 0.265]) - hinge
        elif kind == "laptop":
            body = _box_surface(rs, 2 * n, (0.6, 0.03, 0.4))
            door = _box_surface(rs, 2 * n, (0.6, 0.4, 0.02))
            theta = rs.uniform(-0.45 * np.pi, 0.05 * np.pi) - 0.5 * np.pi * 0.0
            hinge, axis = np.array([0.0, 0.015, -0.2]), np.array([1.0, 0.0, 0.0])        # rear edge of the base
            door = door + np.array([0.0, 0.215, -0.2]) - hinge
        else:
            raise ValueError(f"unknown category {kind!r}")
        door = Rotation.from_rotvec(axis * theta).apply(door) + hinge
        pts = np.concatenate([body, door], 0)
        pts = _fps_numpy(pts, n)
        lo, hi = pts.min(0), pts.max(0)
        pts = (pts - (lo + hi) / 2) / np.linalg.norm(hi - lo)
        pts = Rotation.random(random_state=rs.randint(1 << 30)).apply(pts)
        out[b] = pts.astype(np.float32)
    return torch.from_numpy(out)
