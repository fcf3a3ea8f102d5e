"""CPU suite (needs the reference tree mounted; skipped on the GPU box): the reference's OWN block builders and
model registry load and construct unchanged on top of the drop-in `vgtk` / `extensions.chamfer_dist` packages --
the "SPConvNets ... loads unchanged" part of the boundary.  Runs in a subprocess so that the drop-in modules do
not leak into the other tests' interpreter."""
import os
import subprocess
import sys
import textwrap

import pytest

REF = os.environ.get("VGTK_REFERENCE_ROOT", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "SPConvNets")), reason="reference tree not mounted")


def _run(code):
    env = dict(os.environ, LOCAL_RANK="0", PYTHONWARNINGS="ignore")
    r = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], capture_output=True, text=True, cwd=ROOT, env=env,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return r.stdout


PRELUDE = f"""
import sys, os
sys.path.insert(0, {ROOT!r})
import numpy as np, torch
if not hasattr(np, 'float'): np.float = float
import equi_articulated_pose_b200 as eap
eap.install()
import vgtk
assert vgtk.__file__.startswith({ROOT!r})
sys.path.insert(0, {REF!r})
"""


def test_reference_block_builders_construct_on_dropin():
    out = _run(PRELUDE + """
import importlib, io, contextlib
sys.path.insert(0, 'tests/golden')
from make_golden import small_params
with contextlib.redirect_stdout(io.StringIO()):
    M = importlib.import_module('SPConvNets.utils.base_so3conv')
    MP = importlib.import_module('SPConvNets.utils.base_so3poseconv')
    blk = torch.nn.ModuleList([M.BasicSO3ConvBlock(p) for p in small_params()])
keys = set(blk.state_dict().keys())
g = np.load('tests/golden/ref_blocks_small.npz')
want = {k[len('state/backbone.'):] for k in g.files if k.startswith('state/')}
assert want <= keys, sorted(want - keys)[:5]
import vgtk.so3conv as sptk
assert any(isinstance(m, sptk.InterSO3Conv) for m in blk.modules()) and any(isinstance(m, sptk.IntraSO3Conv) for m in blk.modules())
print('OK', len(keys))
""")
    assert "OK" in out


def test_model38_registry_builds_on_dropin():
    """run_unsup_arti_align.py maps --use-equi=38 to this registry entry (run_unsup_arti_align.py:8-17)."""
    out = _run(PRELUDE + """
import io, contextlib
from equi_articulated_pose_b200 import blocks          # before the reference's own `extensions` package gets imported
torch.Tensor.cuda = lambda self, *a, **k: self          # the model constructor calls .cuda() (no GPU here)
sys.argv = ['run_unsup_arti_align.py', '-d', '/tmp/none', '--use-equi=38', '--kanchor=60', '--kpconv-kanchor=60',
            '--input-num=512', '--bsz=1', '--nmasks=2', '--cur-stage=0']
with contextlib.redirect_stdout(io.StringIO()):
    from SPConvNets.options import opt
    import SPConvNets.models as models
    name = 'unsup_seg_so3_pose_conv_pn_38_multi_stage'
    opt.model.model = name
    opt.device = torch.device('cpu')
    net = getattr(models, name).build_model_from(opt, None)
import vgtk.so3conv as sptk
n_pose = sum(isinstance(m, sptk.InterSO3PoseConv) for m in net.modules())
n_intra = sum(isinstance(m, sptk.IntraSO3Conv) for m in net.modules())
assert n_pose >= 3 and n_intra >= 3, (n_pose, n_intra)
from extensions.chamfer_dist import ChamferDistance
assert ChamferDistance.__module__.startswith('extensions.chamfer_dist')
# the per-layer geometry the reference's builder hands to its convolutions == blocks.model38_backbone_params()
mine = [l['args'] for blk in blocks.model38_backbone_params(input_num=512) for l in blk]
theirs = [m for m in net.modules() if isinstance(m, sptk.InterSO3Conv)]
groups = [theirs[i:i + len(mine)] for i in range(0, len(theirs) - len(mine) + 1, len(mine))]
assert groups, len(theirs)
for grp in groups:
    for a, m in zip(mine, grp):
        assert (a['dim_in'], a['dim_out'], a['stride'], a['n_neighbor']) == (m.dim_in, m.dim_out, m.stride, m.n_neighbor), (a, m.dim_in, m.dim_out)
        assert abs(a['radius'] - m.radius) < 1e-12 and abs(a['sigma'] - m.sigma) < 1e-12, (a, m.radius, m.sigma)
print('OK', sum(p.numel() for p in net.parameters()), len(theirs))
""")
    assert "OK" in out
