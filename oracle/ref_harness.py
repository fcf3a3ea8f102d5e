"""Import the reference's own Python stack IN PLACE from /root/reference on a CPU-only host.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py (fixture generation) and by
the `-m "not gpu"` tests that cross-check oracle/so3.py against the reference when the
reference tree is mounted.  It never travels to the GPU box (the tree is absent there) and
the product never imports it.

What is stubbed, nothing is copied:
  * ``plyfile``   -- two tiny PLY readers (ASCII vertex list; binary-LE icosahedron)
  * ``trimesh``   -- ``load()`` -> faces / face_normals / face_adjacency / fix_normals()
  * ``vgtk.cuda.{gathering,grouping,zpconv}`` and top-level ``chamfer`` -- CPU ops backed by
    oracle/oracle_ops.c (restatements of the reference kernels)
  * ``np.float``, ``torch.cuda.synchronize``, ``Tensor.cuda`` and ``LOCAL_RANK``
"""
import os
import struct
import sys
import types

import numpy as np

REF = os.environ.get("VGTK_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "vgtk", "vgtk"))


# ----------------------------------------------------------------------------- PLY readers
def read_ply(path):
    """Return (vertices float32 [V,3], faces int32 [F,3] or None)."""
    with open(path, "rb") as fh:
        raw = fh.read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    header = raw[:end].decode("ascii", "replace").splitlines()
    body = raw[end:]
    fmt = [l.split()[1] for l in header if l.startswith("format")][0]
    elems, cur = [], None
    for l in header:
        t = l.split()
        if not t:
            continue
        if t[0] == "element":
            cur = {"name": t[1], "count": int(t[2]), "props": []}
            elems.append(cur)
        elif t[0] == "property" and cur is not None:
            cur["props"].append(t[1:])
    size = {"float": 4, "float32": 4, "uchar": 1, "uint8": 1, "int": 4, "int32": 4, "uint": 4, "double": 8}
    code = {"float": "f", "float32": "f", "uchar": "B", "uint8": "B", "int": "i", "int32": "i", "uint": "I", "double": "d"}
    verts, faces = None, None
    if fmt == "ascii":
        lines = body.decode("ascii").strip().splitlines()
        pos = 0
        for e in elems:
            rows = [lines[pos + i].split() for i in range(e["count"])]
            pos += e["count"]
            if e["name"] == "vertex":
                verts = np.array([[float(r[0]), float(r[1]), float(r[2])] for r in rows], np.float32)
            elif e["name"] == "face":
                faces = np.array([[int(v) for v in r[1:1 + int(r[0])]] for r in rows], np.int32)
        return verts, faces
    assert fmt == "binary_little_endian"
    off = 0
    for e in elems:
        rows = []
        for _ in range(e["count"]):
            row = []
            for p in e["props"]:
                if p[0] == "list":
                    cnt = struct.unpack_from("<" + code[p[1]], body, off)[0]
                    off += size[p[1]]
                    vals = struct.unpack_from("<" + code[p[2]] * cnt, body, off)
                    off += size[p[2]] * cnt
                    row.append(list(vals))
                else:
                    row.append(struct.unpack_from("<" + code[p[0]], body, off)[0])
                    off += size[p[0]]
            rows.append(row)
        if e["name"] == "vertex":
            verts = np.array([r[:3] for r in rows], np.float32)
        elif e["name"] == "face":
            faces = np.array([r[0] for r in rows], np.int32)
    return verts, faces


def _plyfile_stub():
    m = types.ModuleType("plyfile")

    class PlyData(dict):
        @staticmethod
        def read(path):
            v, f = read_ply(path)
            d = PlyData()
            d["vertex"] = {"x": v[:, 0], "y": v[:, 1], "z": v[:, 2]}
            if f is not None:
                d["face"] = {"vertex_indices": list(f)}
            return d

    class PlyElement:  # only referenced by save paths we never call
        pass

    m.PlyData, m.PlyElement = PlyData, PlyElement
    return m


class _Mesh:
    """What functional/rotation.py:117-139,236-243 reads from a trimesh.Trimesh."""

    def __init__(self, verts, faces):
        self.vertices = verts.astype(np.float64)
        self.faces = faces.astype(np.int64)
        tri = self.vertices[self.faces]
        nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
        self.face_normals = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
        # trimesh 3.2: shared edges in sorted-edge order, each pair sorted by face index
        edges = {}
        for fi, f in enumerate(self.faces):
            for a, b in ((f[0], f[1]), (f[1], f[2]), (f[2], f[0])):
                edges.setdefault((min(a, b), max(a, b)), []).append(fi)
        # trimesh hashes an edge row as (v0 | v1 << 32)-like packed integer and sorts on it
        keys = sorted(edges, key=lambda e: (e[1], e[0]))
        self.face_adjacency = np.array([sorted(edges[k]) for k in keys if len(edges[k]) == 2], np.int64)

    def fix_normals(self):
        # sphere12.ply is already outward wound (checked: normals . centroids > 0)
        c = self.vertices[self.faces].mean(1)
        assert ((self.face_normals * c).sum(1) > 0).all()


def _trimesh_stub():
    m = types.ModuleType("trimesh")
    m.load = lambda path, *a, **k: _Mesh(*read_ply(path))
    return m


# ----------------------------------------------------------------------------- CUDA op shims
def _cuda_stubs():
    import torch
    from . import cops

    grouping = types.ModuleType("vgtk.cuda.grouping")
    gathering = types.ModuleType("vgtk.cuda.gathering")
    zpconv = types.ModuleType("vgtk.cuda.zpconv")
    chamfer = types.ModuleType("chamfer")

    def t(a):
        return torch.from_numpy(np.ascontiguousarray(a))

    grouping.ball_query = lambda new_xyz, xyz, r, ns: t(cops.ball_query(new_xyz.detach().numpy(), xyz.detach().numpy(), r, ns))
    grouping.furthest_point_sampling = lambda xyz, m: t(cops.furthest_point_sampling(xyz.detach().numpy(), m))
    gathering.gather_points_forward = lambda pts, idx: t(cops.gather_points_forward(pts.detach().numpy(), idx.numpy()))
    gathering.gather_points_backward = lambda g, idx, n: t(cops.gather_points_backward(g.detach().numpy(), idx.numpy(), n))

    def ch_fwd(a, b):
        return tuple(t(x) for x in cops.chamfer_forward(a.detach().numpy(), b.detach().numpy()))

    def ch_bwd(a, b, i1, i2, g1, g2):
        return tuple(t(x) for x in cops.chamfer_backward(a.detach().numpy(), b.detach().numpy(), i1.numpy(), i2.numpy(),
                                                         g1.detach().numpy(), g2.detach().numpy()))

    chamfer.forward, chamfer.backward = ch_fwd, ch_bwd
    pkg = types.ModuleType("vgtk.cuda")
    pkg.__path__ = []
    pkg.grouping, pkg.gathering, pkg.zpconv = grouping, gathering, zpconv
    return {"vgtk.cuda": pkg, "vgtk.cuda.grouping": grouping, "vgtk.cuda.gathering": gathering,
            "vgtk.cuda.zpconv": zpconv, "chamfer": chamfer}


_installed = False


def install():
    """Make ``import vgtk`` / ``import SPConvNets.utils.base_so3conv`` resolve to the reference."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not mounted at " + REF)
    import torch

    for name in list(sys.modules):
        if name == "vgtk" or name.startswith("vgtk.") or name == "chamfer":
            raise RuntimeError("a vgtk implementation is already imported in this process; "
                               "run the reference harness in its own interpreter")
    if not hasattr(np, "float"):
        np.float = float
    os.environ.setdefault("LOCAL_RANK", "0")
    if not torch.cuda.is_available():
        torch.cuda.synchronize = lambda *a, **k: None
        torch.Tensor.cuda = lambda self, *a, **k: self
    sys.modules["plyfile"] = _plyfile_stub()
    sys.modules["trimesh"] = _trimesh_stub()
    sys.modules.update(_cuda_stubs())
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "vgtk"))
    _installed = True


def import_blocks():
    """The reference block builders, imported directly (the package __init__ star-imports the
    pose variant over the classic one: SPConvNets/utils/__init__.py:6-7)."""
    install()
    import importlib
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        import vgtk  # noqa: F401
        mod = importlib.import_module("SPConvNets.utils.base_so3conv")
    return mod
