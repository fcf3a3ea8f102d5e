"""CPU suite: the C-ABI library builds, loads and exports every symbol include/vgtkb.h declares,
and the Python binding declares a prototype for each of them.  No compute call is made here."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vgtkb.h")
CSRC = os.path.join(ROOT, "equi_articulated_pose_b200", "csrc")


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vgtkb_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def libpath():
    subprocess.check_call(["make", "-C", CSRC, "-j", "8"], stdout=subprocess.DEVNULL)
    from equi_articulated_pose_b200 import lib
    assert os.path.exists(lib.LIB_PATH)
    return lib.LIB_PATH


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("vgtkb_ball_query", "vgtkb_furthest_point_sampling", "vgtkb_gather_points_forward",
                 "vgtkb_gather_points_backward", "vgtkb_chamfer_forward", "vgtkb_chamfer_backward",
                 "vgtkb_inter_group_forward", "vgtkb_intra_group_forward", "vgtkb_gemm_nt"):
        assert must in syms


def test_library_exports_every_declared_symbol(libpath):
    so = ctypes.CDLL(libpath)
    for s in declared_symbols():
        assert hasattr(so, s), f"{s} declared in vgtkb.h but not exported"


def test_binding_covers_every_declared_symbol(libpath):
    from equi_articulated_pose_b200 import lib
    bound = set(lib.SIGNATURES) | set(lib.NO_STATUS) | set(lib.SETUP_SIGNATURES)
    assert set(declared_symbols()) == bound
    so = lib.load()
    assert so.vgtkb_version() == lib.ABI_VERSION
    assert isinstance(so.vgtkb_last_error(), bytes)


def test_ops_refuse_cpu_tensors(libpath):
    """No CPU fallback: a CPU tensor (or a missing device) raises instead of computing."""
    import torch
    from equi_articulated_pose_b200 import ops, lib
    xyz = torch.zeros(1, 3, 8)
    with pytest.raises((lib.VgtkbError, RuntimeError, AssertionError)):
        ops.ball_query(xyz, xyz, 0.1, 4)


def test_sass_has_blackwell_paths(libpath):
    """The shipped binary is sm_100a code with TMA bulk copies (UBLKCP) in it."""
    out = subprocess.run(["cuobjdump", "-sass", libpath], capture_output=True, text=True).stdout
    assert "sm_100a" in out or "SM100" in out.upper()
    assert "UBLKCP" in out


def test_product_synthetic_generators_match_the_checkers():
    """bench.py's CUDA arm builds its clouds / weights from the product package (nothing under oracle/ on that arm); the CPU
    baseline arm uses the oracle's generators: both must see identical tensors."""
    import torch
    from equi_articulated_pose_b200 import blocks, synthetic
    from oracle import so3 as O
    for n in (1024, 4096):
        assert blocks.backbone_params(input_num=n) == O.backbone_params(input_num=n) or \
            all(a['args'][k] == b['args'][k] for pa, pb in zip(blocks.backbone_params(input_num=n), O.backbone_params(input_num=n))
                for a, b in zip(pa, pb) for k in b['args'])
    p = blocks.backbone_params(input_num=1024)
    s1, s2 = synthetic.init_backbone_state(p, seed=0), O.init_backbone_state(O.backbone_params(input_num=1024), seed=0)
    assert s1.keys() == s2.keys() and all(torch.equal(s1[k], s2[k]) for k in s1)
    assert torch.equal(synthetic.synthetic_cloud(3, 1024, 2000), O.synthetic_cloud(3, 1024, 2000))


def test_pointnetpp_constructs_with_reference_keys_and_validates_sizes():
    """CPU-side checks of the PointnetPP drop-in (no kernel launches): state-dict keys of the reference class
    (SPConvNets/models/PointNet2.py:22-64) and a clear error for clouds smaller than a level."""
    import pytest
    import torch
    from equi_articulated_pose_b200.pointnet2 import PointnetPP
    from oracle import pointnet2 as OP
    net = PointnetPP(6)
    assert set(net.state_dict().keys()) == set(OP.make_state(6).keys())
    net.load_state_dict(OP.make_state(6, seed=1))
    two = PointnetPP(6, type("A", (), {"pnpp_n_layers": 2})())
    assert set(two.state_dict().keys()) == set(OP.make_state(6, n_layers=2).keys())
    with pytest.raises(ValueError):
        net(torch.zeros(1, 100, 3), torch.zeros(1, 100, 3))           # 100 points < 512 centres: rejected before any launch


def test_articulated_cloud_generators():
    """synthetic 'oven' / 'laptop' clouds (BASELINE configs 3 / 5): deterministic per seed, unit bounding-box diagonal before
    the global rotation (so every point within 0.5 of the centre), two parts, distinct poses per sample."""
    import numpy as np
    import torch
    from equi_articulated_pose_b200 import synthetic
    for kind in ("oven", "laptop"):
        a = synthetic.articulated_cloud(kind, 3, 192, 11)
        b = synthetic.articulated_cloud(kind, 3, 192, 11)
        assert a.dtype == torch.float32 and tuple(a.shape) == (3, 192, 3) and torch.equal(a, b)
        assert not torch.equal(a, synthetic.articulated_cloud(kind, 3, 192, 12))
        assert float(a.norm(dim=2).max()) <= 0.5 + 1e-6           # inside the sphere circumscribing the unit-diagonal box
        assert float(torch.cdist(a[0], a[0]).max()) > 0.6          # and spanning most of it
        assert not torch.allclose(a[0], a[1])
        d = torch.cdist(a[0], a[0]) + torch.eye(192) * 10
        assert float(d.min()) > 1e-4                                # farthest-point sampled: no duplicate points
    import pytest
    with pytest.raises(ValueError):
        synthetic.articulated_cloud("chair", 1, 16, 0)
