"""BASELINE config 3 / 5 shape on the GPU: the equivariant backbone model 38 (`--use-equi=38`, kanchor 60) builds -- three
stride-1 separable blocks 64 / 128 / 512 with 64-neighbour balls (blocks.model38_backbone_params, checked against the
reference's own builder in tests/test_reference_compat.py) -- on synthetic 'oven' / 'laptop' clouds, fwd + bwd against the
CPU oracle.  n_neighbor = 64 is beyond the 32-neighbour tensor-core grouping kernels, so this also covers the exact fp32
grouping kernels inside a full block stack."""
import pytest
import torch

from tests.helpers import rel_err, build_backbone

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-4


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from equi_articulated_pose_b200 import lib
    lib.load()
    return torch.device("cuda:0")


def _oracle(params, sd, pts, dtype):
    from oracle import so3 as O
    from equi_articulated_pose_b200 import so3_constants as C
    b, n, _ = pts.shape
    sdo = {k: v.clone().to(dtype).requires_grad_(True) for k, v in sd.items()}
    for bi, blk in enumerate(params):
        for li, layer in enumerate(blk):
            co = layer['args']['dim_out']
            for pre in (f'backbone.{bi}.blocks.{li}.inter_conv.norm.', f'backbone.{bi}.blocks.{li}.norm.'):
                sdo[pre + 'running_mean'], sdo[pre + 'running_var'] = torch.zeros(co, dtype=dtype), torch.ones(co, dtype=dtype)
    xyz = pts.permute(0, 2, 1).contiguous().to(dtype)
    oxyz, of = O.backbone_forward(sdo, params, xyz, torch.ones(b, 1, n, 60, dtype=dtype),
                                  torch.from_numpy(C.anchors_all()).to(dtype), torch.from_numpy(C.intra_idx()),
                                  C.kernel_points_base(), training=True)
    w = torch.randn(of.shape, generator=torch.Generator().manual_seed(5)).to(dtype)
    loss = (of * w).mean()
    loss.backward()
    return sdo, of.detach(), oxyz, loss.detach(), w


@pytest.mark.parametrize("kind,seed", [("oven", 3000), ("laptop", 5000)])
def test_model38_backbone_on_articulated_clouds_vs_oracle(dev, kind, seed):
    from equi_articulated_pose_b200 import blocks, synthetic
    n, b = 256, 2
    params = blocks.model38_backbone_params(input_num=n)
    assert [l['args']['n_neighbor'] for blk in params for l in blk] == [64, 64, 64]
    sd = synthetic.init_backbone_state(params, seed=21)
    pts = synthetic.articulated_cloud(kind, b, n, seed)
    sdo, of, oxyz, loss, w = _oracle(params, sd, pts, torch.float32)
    net = build_backbone(params, sd, dev).train()
    out = net(pts.to(dev))
    assert torch.equal(out.xyz.cpu(), oxyz)                         # stride 1: every point kept, in order
    assert tuple(out.feats.shape) == (b, 512, n, 60)
    assert rel_err(out.feats, of) < FP32_TOL
    l2 = (out.feats * w.to(dev)).mean()
    assert abs(float(l2) - float(loss)) < 2e-4 * max(abs(float(loss)), 1e-3)
    l2.backward()
    sd64, _, _, _, _ = _oracle(params, sd, pts, torch.float64)
    rows = []
    for name, p in net.named_parameters():
        truth = sd64[name].grad
        scale = float(truth.abs().max())
        rows.append((name, scale, float((p.grad.double().cpu() - truth).abs().max()),
                     float((sdo[name].grad.double() - truth).abs().max())))
    gmax = max(r[1] for r in rows)
    for name, scale, e_gpu, e_ref in rows:
        if scale < 1e-6 * gmax:          # structurally zero gradients (bias in front of BatchNorm, constant first skip branch)
            assert e_gpu < 1e-4 * gmax, (name, e_gpu)
        else:                            # same bar as the classic backbone test: as close to fp64 as the fp32 reference
            assert e_gpu <= max(5 * e_ref, 5e-2 * scale), (name, e_gpu / scale, e_ref / scale)
    print(kind, "fwd rel err", rel_err(out.feats, of), "worst grad err / scale", max(r[2] / r[1] for r in rows if r[1] >= 1e-6 * gmax))
