"""CPU: the index tables of ops.WeightPlanes (one launch per step produces the bf16 operand planes of every conv weight)
name the same elements as the per-call permute / transpose copies they replace."""
import numpy as np
import torch

from equi_articulated_pose_b200 import ops


def emulate(src, sizes, strides):
    """What vgtkb_weight_planes' kernel reads: dst[i0, i1, i2] = src_flat[i0*s0 + i1*s1 + i2*s2] (before the hi / lo split)."""
    i0, i1, i2 = np.meshgrid(*[np.arange(n) for n in sizes], indexing="ij")
    return src.reshape(-1)[i0 * strides[0] + i1 * strides[1] + i2 * strides[2]].reshape(-1)


def test_plane_orders_match_the_per_call_copies():
    g = torch.Generator().manual_seed(0)
    for role, co, ci, k in (("inter", 16, 32, 24), ("intra", 64, 64, 12), ("linear", 24, 8, 1)):
        w = torch.randn(co, ci * k, generator=g)
        pw = ops.PreparedWeight(w, co, ci, k, role)
        fwd, bwd = ops.WeightPlanes._items(pw)
        w_kc = w.view(co, ci, k).transpose(1, 2).reshape(co, -1)                 # BasicSO3Conv.weight_kc()
        assert torch.equal(pw.kc(), w_kc)
        np.testing.assert_array_equal(emulate(w.numpy(), *fwd), w_kc.contiguous().numpy().reshape(-1))
        if role == "intra":                                                     # IntraConvFn.backward's operand
            ref = w_kc.view(co, k, ci).permute(2, 1, 0).reshape(ci, k * co)
        else:                                                                   # W^T of the kernel-order matrix / of the 1x1 conv
            ref = w_kc.t()
        np.testing.assert_array_equal(emulate(w.numpy(), *bwd), ref.contiguous().numpy().reshape(-1))
        # gradient w.r.t. the kernel-order matrix -> the parameter's own column order
        gkc = torch.randn(co, k * ci, generator=g)
        wl = w.clone().requires_grad_(True)
        (wl.view(co, ci, k).transpose(1, 2).reshape(co, -1) * gkc).sum().backward()
        assert torch.equal(pw.grad_from_kc(gkc), wl.grad)


def test_prepared_weight_validity_follows_the_version_counter():
    w = torch.nn.Parameter(torch.randn(8, 64))
    pw = ops.PreparedWeight(w, 8, 64, 1, "linear")
    assert not pw.valid()                                   # nothing prepared yet
    pw.fwd = pw.bwd = torch.zeros(2 * 8 * 64, dtype=torch.bfloat16)
    pw.version = w._version
    assert pw.valid()
    with torch.no_grad():
        w.add_(1.0)                                         # an optimizer step
    assert not pw.valid()


# ---------------------------------------------------------------------------------- GPU
import pytest  # noqa: E402


@pytest.mark.gpu
def test_weight_planes_kernel_matches_per_call_split():
    """vgtkb_weight_planes (one launch for all items) == permute copy + vgtkb_split_bf16 per weight, bit for bit."""
    from equi_articulated_pose_b200 import lib, ops
    lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    planes, refs = ops.WeightPlanes(), []
    for role, co, ci, k in (("inter", 64, 32, 24), ("intra", 128, 64, 12), ("linear", 64, 8, 1), ("inter", 256, 256, 24),
                            ("linear", 256, 128, 1), ("intra", 64, 64, 12)):
        w = torch.nn.Parameter(torch.randn(co, ci * k, generator=g).to(dev))
        pw = planes.add(w, co, ci, k, role)
        w_kc = pw.kc().detach().contiguous()
        bwd = w_kc.view(co, k, ci).permute(2, 1, 0).reshape(ci, k * co) if role == "intra" else w_kc.t()
        refs.append((pw, w_kc, bwd.contiguous()))
    planes.prepare()
    for pw, fwd, bwd in refs:
        assert pw.valid()
        for got, ref in ((pw.fwd, fwd), (pw.bwd, bwd)):
            hi, lo = ops.split_bf16(ref)
            n = ref.numel()
            assert torch.equal(got[:n].view(torch.int16), hi.reshape(-1).view(torch.int16))
            assert torch.equal(got[n:].view(torch.int16), lo.reshape(-1).view(torch.int16))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [3, 4])
def test_backbone_with_prepared_weights_is_bit_identical(mode, monkeypatch):
    """Classic backbone fwd + bwd with the weight planes produced once per step (default) against the per-call permute /
    transpose / split launches (VGTKB_WEIGHT_PLANES=0): same operand bits into the same kernels, so the output is
    bit-identical; the gradients agree to the order of the weight-gradient kernels' red.global.add partial sums."""
    from equi_articulated_pose_b200 import blocks, lib, ops, synthetic
    lib.load()
    dev = torch.device("cuda:0")
    ops.set_gemm_mode(mode)
    try:
        params = blocks.backbone_params(input_num=256)
        clouds = synthetic.synthetic_cloud(2, 256, 77).to(dev)
        res = {}
        for flag in ("1", "0", "0b"):                       # "0b": second run of the same path = run-to-run spread of the atomics
            monkeypatch.setenv("VGTKB_WEIGHT_PLANES", flag[0])
            net = blocks.SO3Backbone(params)
            net.load_state_dict(synthetic.init_backbone_state(params, seed=0), strict=False)
            net = net.to(dev).train()
            kernels0 = lib.COUNTERS["kernels"]
            out = net(clouds)
            out.feats.square().mean().backward()
            res[flag] = (out.feats.detach().clone(), {n: p.grad.detach().clone() for n, p in net.named_parameters()},
                         lib.COUNTERS["kernels"] - kernels0)
            n_prepared = sum(1 for m in net.modules() if getattr(m, '_wp', None) is not None and m._wp.valid())
            assert (n_prepared > 15) == (flag == "1"), n_prepared
        assert torch.equal(res["1"][0], res["0"][0])
        # the launches it is there to remove (mode 4 has no activation planes: the intra convs and the inter conv backward
        # keep their per-call weight handling there)
        assert res["1"][2] < res["0"][2] - (40 if mode == 3 else 10), (res["1"][2], res["0"][2])
        for n, g1 in res["1"][1].items():
            g0, g0b = res["0"][1][n], res["0b"][1][n]
            assert g1.shape == g0.shape
            scale = float(g0.abs().max()) + 1e-30
            # the scatter and the weight-gradient reductions are atomics: two runs of ONE path already differ.  `spread` is a
            # single sample of that difference, so the bar leaves room (10x); in mode 4 an upstream difference of one ulp can
            # flip a bf16 rounding of an operand (2^-9 relative), which moves a gradient by ~1e-4 .. 1e-3 of its maximum
            spread = float((g0b - g0).abs().max())
            floor = (1e-5 if mode == 3 else 2e-3) * scale
            assert float((g1 - g0).abs().max()) <= max(10.0 * spread, floor), (n, spread / scale)
    finally:
        ops.set_gemm_mode(3)
