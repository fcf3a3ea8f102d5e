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
