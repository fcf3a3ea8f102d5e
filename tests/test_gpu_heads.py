"""Invariant output heads (equi_articulated_pose_b200/heads.py) against a fixture of the reference's own modules
(SPConvNets/utils/base_so3conv.py:481-645, 766-840, 1013-1150; tests/golden/make_golden.py heads): same state dict, train
mode, forward + backward."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = [
    ("r_att", "InvOutBlockR", {}, {"pooling": "attention"}),
    ("r_max", "InvOutBlockR", {}, {"pooling": "max"}),
    ("pn_max", "InvOutBlockPointnet", {}, {"pooling": "max"}),
    ("mvd", "InvOutBlockMVD", {}, {}),
    ("ours_max", "InvOutBlockOurs", {"pooling_method": "max"}, {}),
    ("mask_att", "InvOutBlockOursWithMask", {"norm": 1, "pooling_method": "attention", "use_pointnet": True}, {}),
]


def rel_err(a, b):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


@pytest.mark.parametrize("name,cls,kw,over", CASES)
def test_head_matches_reference_fixture(name, cls, kw, over):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from equi_articulated_pose_b200 import heads
    import vgtk.so3conv as sptk
    dev = torch.device("cuda:0")
    g = np.load(os.path.join(GOLD, "ref_heads_small.npz"))
    params = {"dim_in": 16, "mlp": [32, 64], "fc": [64], "k": 8, "kanchor": 60, "temperature": 3.0}
    params.update(over)
    head = getattr(heads, cls)(params, **kw)
    sd = {k[len(name) + 4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(name + "_sd_")}
    missing, unexpected = head.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(m.endswith("anchors") for m in missing), missing       # buffers the reference keeps elsewhere
    head = head.to(dev).train()
    feats = torch.from_numpy(g["feats"]).to(dev).requires_grad_(True)
    xyz = torch.from_numpy(g["xyz"]).to(dev)
    anchors = torch.from_numpy(sptk.get_anchors(60)).to(dev)
    if cls == "InvOutBlockR":
        res = head(feats)
    elif cls == "InvOutBlockOursWithMask":
        res = head(sptk.SphericalPointCloud(xyz, feats, anchors), torch.from_numpy(g["mask"]).to(dev),
                   soft_mask=torch.from_numpy(g["soft_mask"]).to(dev))
    else:
        res = head(sptk.SphericalPointCloud(xyz, feats, anchors))
    res = res if isinstance(res, tuple) else (res,)
    gg = torch.Generator().manual_seed(7)
    loss = 0
    for i, r in enumerate(res):
        want = g[f"{name}_out{i}"]
        assert tuple(r.shape) == tuple(want.shape), (i, r.shape, want.shape)
        assert rel_err(r, want) < 2e-4, (name, i, rel_err(r, want))
        loss = loss + (r * torch.randn(r.shape, generator=gg).to(dev)).sum()
    loss.backward()
    assert rel_err(feats.grad, g[f"{name}_grad_feats"]) < 1e-3, rel_err(feats.grad, g[f"{name}_grad_feats"])
    # parameter gradients: relative to the tensor's own maximum, with a floor at 1e-5 of the largest parameter gradient of the
    # head (the bias of an attention layer feeds a softmax: its true gradient is zero and the reference value is rounding noise)
    gmax = max(float(np.abs(g[f"{name}_pg_{k}"]).max()) for k, _ in head.named_parameters() if g[f"{name}_pg_{k}"].size)
    for k, p in head.named_parameters():
        want = g[f"{name}_pg_{k}"]
        if want.size == 0:
            continue
        err = float((p.grad.detach().double().cpu() - torch.from_numpy(want).double()).abs().max())
        assert err < 2e-3 * max(float(np.abs(want).max()), 1e-2 * gmax), (name, k, err, float(np.abs(want).max()), gmax)
