"""PLY reading/writing without the `plyfile` dependency (reference: vgtk/vgtk/pc/io.py:6-31)."""
import struct

import numpy as np

_SIZE = {"float": 4, "float32": 4, "uchar": 1, "uint8": 1, "char": 1, "int": 4, "int32": 4, "uint": 4,
         "short": 2, "ushort": 2, "double": 8, "float64": 8}
_CODE = {"float": "f", "float32": "f", "uchar": "B", "uint8": "B", "char": "b", "int": "i", "int32": "i",
         "uint": "I", "short": "h", "ushort": "H", "double": "d", "float64": "d"}


def _parse(path):
    with open(path, "rb") as fh:
        raw = fh.read()
    end = raw.index(b"end_header") + len(b"end_header")
    end = raw.index(b"\n", end) + 1
    lines = raw[:end].decode("ascii", "replace").splitlines()
    fmt, elems = "ascii", []
    for ln in lines:
        t = ln.split()
        if not t:
            continue
        if t[0] == "format":
            fmt = t[1]
        elif t[0] == "element":
            elems.append({"name": t[1], "count": int(t[2]), "props": []})
        elif t[0] == "property" and elems:
            elems[-1]["props"].append(t[1:])
    body = raw[end:]
    data = {}
    if fmt == "ascii":
        rows = body.decode("ascii").split("\n")
        pos = 0
        for e in elems:
            recs = []
            for _ in range(e["count"]):
                tok = rows[pos].split()
                pos += 1
                rec, ti = {}, 0
                for pr in e["props"]:
                    if pr[0] == "list":
                        cnt = int(tok[ti]); ti += 1
                        rec[pr[3]] = [float(v) for v in tok[ti:ti + cnt]]; ti += cnt
                    else:
                        rec[pr[1]] = float(tok[ti]); ti += 1
                recs.append(rec)
            data[e["name"]] = recs
        return data
    end_c = "<" if fmt == "binary_little_endian" else ">"
    off = 0
    for e in elems:
        recs = []
        for _ in range(e["count"]):
            rec = {}
            for pr in e["props"]:
                if pr[0] == "list":
                    cnt = struct.unpack_from(end_c + _CODE[pr[1]], body, off)[0]; off += _SIZE[pr[1]]
                    rec[pr[3]] = list(struct.unpack_from(end_c + _CODE[pr[2]] * cnt, body, off)); off += _SIZE[pr[2]] * cnt
                else:
                    rec[pr[1]] = struct.unpack_from(end_c + _CODE[pr[0]], body, off)[0]; off += _SIZE[pr[0]]
            recs.append(rec)
        data[e["name"]] = recs
    return data


def load_ply(file_name, with_faces=False, with_color=False, with_normal=False):
    d = _parse(file_name)
    v = d["vertex"]
    points = np.array([[r["x"], r["y"], r["z"]] for r in v], dtype=np.float32)
    ret = [points]
    if with_faces:
        ret.append(np.vstack([np.asarray(r["vertex_indices"], dtype=np.int32) for r in d["face"]]))
    if with_color:
        ret.append(np.array([[r["red"], r["green"], r["blue"]] for r in v]))
    pc = ret[0] if len(ret) == 1 else ret
    if with_normal:
        return pc, np.array([[r["nx"], r["ny"], r["nz"]] for r in v], dtype=np.float32)
    return pc


def save_ply(filepath, color_pc, c=None, use_color=False, use_normal=False, verbose=False):
    pc = np.asarray(color_pc)
    palette = {'r': (255, 0, 0), 'g': (0, 255, 0), 'b': (0, 0, 255)}
    with open(filepath, 'w') as f:
        f.write("ply\nformat ascii 1.0\n")
        f.write(f"element vertex {int(pc.shape[0])}\n")
        f.write("property float x\nproperty float y\nproperty float z\n")
        if use_normal:
            f.write("property float nx\nproperty float ny\nproperty float nz\n")
        f.write("property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n")
        for row in pc:
            xyz = " ".join(f"{v:.6f}" for v in row[:3])
            nrm = (" " + " ".join(f"{v:.6f}" for v in row[3:6])) if use_normal else ""
            if use_color:
                col = tuple(int(v) for v in row[-3:])
            else:
                col = palette.get(c, (255, 255, 255))
            f.write(f"{xyz}{nrm} {col[0]} {col[1]} {col[2]}\n")
