"""Generate tests/golden/bench_config2_b8.npz: the BENCHMARKED workload (bench.py, BASELINE.json configs[1]: classic backbone,
8 sphere-shell clouds x 1024 points, seed-0 random init, train mode) evaluated fwd + bwd by the oracle (oracle/so3.py, the
CPU restatement pinned on the reference's own outputs) in fp32 -- the reference's arithmetic -- and in fp64 -- ground truth.

Stored (small fixture, ~2 MB): the two losses of bench.py's step 0 ('square' = feats.square().mean(), the benchmark loss,
and 'proj' = projection on a fixed random tensor, well-conditioned), a seeded subsample of the output features, and for
every parameter a seeded subsample (<= 4096 entries, all entries of small tensors) of its gradient under the 'proj' loss:
the fp64 value and the fp32 oracle's value, so a test can report e_gpu = |g_gpu - g64| and e_ref = |g32 - g64| per tensor
against the same ground truth.  Activation checkpointing per layer keeps the fp64 run inside the container's 62 GB.

    python tests/golden/make_bench_golden.py        (CPU, ~10 min on 8 cores)
"""
import os
import sys

import numpy as np
import torch
from torch.utils.checkpoint import checkpoint

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import so3 as O                                   # noqa: E402
from equi_articulated_pose_b200 import so3_constants as C    # noqa: E402

B, N, SEED_W, SEED_X, NSUB = 8, 1024, 0, 2000, 4096


def sub_index(name, numel):
    if numel <= NSUB:
        return np.arange(numel)
    rs = np.random.RandomState(abs(hash(name)) % (1 << 31) if False else sum(ord(ch) for ch in name) * 7919 % (1 << 31))
    return np.sort(rs.choice(numel, NSUB, replace=False))


def run(dtype, loss_kind):
    params = O.backbone_params(input_num=N)
    sd = {k: v.to(dtype).requires_grad_(True) for k, v in O.init_backbone_state(params, seed=SEED_W).items()}
    for bi, blk in enumerate(params):
        for li, layer in enumerate(blk):
            co = layer['args']['dim_out']
            for pre in (f'backbone.{bi}.blocks.{li}.inter_conv.norm.', f'backbone.{bi}.blocks.{li}.norm.'):
                sd[pre + 'running_mean'], sd[pre + 'running_var'] = torch.zeros(co, dtype=dtype), torch.ones(co, dtype=dtype)
    pts = O.synthetic_cloud(B, N, SEED_X)
    xyz = pts.permute(0, 2, 1).contiguous().to(dtype)
    feats = torch.ones(B, 1, N, 60, dtype=dtype)
    anchors, intra = torch.from_numpy(C.anchors_all()).to(dtype), torch.from_numpy(C.intra_idx())
    base_kp = C.kernel_points_base()
    for bi, block in enumerate(params):
        for li, layer in enumerate(block):
            pre = f'backbone.{bi}.blocks.{li}.'

            def f(xyz_, feats_, pre=pre, args=layer['args']):
                return O.separable_block(sd, pre, args, xyz_, feats_, anchors, intra, base_kp, True)
            # first layer: feats do not require grad, checkpoint needs at least one differentiable input -> run it plainly
            if bi == 0 and li == 0:
                xyz, feats = f(xyz, feats)
            else:
                xyz, feats = checkpoint(f, xyz, feats, use_reentrant=False)
    if loss_kind == 'square':
        loss = feats.square().mean()
    else:
        w = torch.randn(feats.shape, generator=torch.Generator().manual_seed(77)).to(dtype)
        loss = (feats * w).mean()
    loss.backward()
    return sd, feats.detach(), xyz.detach(), float(loss)


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    out = {}
    for kind in ('proj', 'square'):
        sd32, f32, x32, l32 = run(torch.float32, kind)
        print(kind, 'fp32 loss', l32, flush=True)
        sd64, f64, x64, l64 = run(torch.float64, kind)
        print(kind, 'fp64 loss', l64, flush=True)
        out[f'loss32_{kind}'], out[f'loss64_{kind}'] = np.float64(l32), np.float64(l64)
        if kind == 'proj':
            fi = sub_index('out_feats', f64.numel())
            out['feats_idx'], out['feats64'], out['feats32'] = fi, f64.reshape(-1).numpy()[fi], f32.reshape(-1).numpy()[fi]
            out['feats_absmax64'] = np.float64(f64.abs().max())
            out['out_xyz'] = x32.numpy()
            for name, p in sd64.items():
                if p.grad is None:
                    continue
                gi = sub_index(name, p.numel())
                g64, g32 = p.grad.reshape(-1).numpy(), sd32[name].grad.reshape(-1).numpy()
                out['gidx/' + name], out['g64/' + name], out['g32/' + name] = gi, g64[gi], g32[gi]
                out['gmax64/' + name] = np.float64(np.abs(g64).max())
                out['eref/' + name] = np.float64(np.abs(g32.astype(np.float64) - g64).max())    # over ALL entries
        else:
            for name, p in sd64.items():
                if p.grad is None:
                    continue
                gi = sub_index(name, p.numel())
                g64, g32 = p.grad.reshape(-1).numpy(), sd32[name].grad.reshape(-1).numpy()
                out['sq_g64/' + name], out['sq_gmax64/' + name] = g64[gi], np.float64(np.abs(g64).max())
                out['sq_eref/' + name] = np.float64(np.abs(g32.astype(np.float64) - g64).max())
    path = os.path.join(ROOT, 'tests', 'golden', 'bench_config2_b8.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
