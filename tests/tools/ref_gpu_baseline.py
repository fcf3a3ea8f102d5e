"""Reference-style GPU baseline (test tooling, not product): the reference's algorithm the way the reference runs
it on a GPU -- its own CUDA kernels (oracle/_ref: ball query, FPS, gather recompiled for sm_100a) plus eager PyTorch
for everything else (materialised kernel weights, gathered neighbour features, einsum, cuBLAS fp32 matmul,
BatchNorm / InstanceNorm) -- timed on the same B200 on BASELINE config 2 (classic backbone fwd+bwd+Adam, N=1024,
A=60, 8 clouds).  The eager math is oracle/so3.py evaluated on CUDA tensors (the reference Python itself cannot
travel to the GPU box); BASELINE.md section 3.2.

    python tests/tools/ref_gpu_baseline.py [--clouds 8] [--steps 5] [--tf32 0]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import so3 as O, build_ref  # noqa: E402
from equi_articulated_pose_b200 import so3_constants as C  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clouds", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--tf32", type=int, default=0)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = bool(args.tf32)
    torch.backends.cudnn.allow_tf32 = bool(args.tf32)
    grp = build_ref.load_ref("vgtk_ref_grouping")
    gat = build_ref.load_ref("vgtk_ref_gathering")
    assert grp is not None and gat is not None, "oracle/_ref not built"

    def ball_query(q, s, radius, n_sample):
        idx = grp.ball_query(q.contiguous(), s.contiguous(), radius, n_sample)
        b, p, nn = idx.shape
        g = gat.gather_points_forward(s.contiguous(), idx.view(b, -1).contiguous()).view(b, 3, p, nn)
        return idx, g

    def furthest_sample_index(xyz, n_sample, lazy):
        if xyz.shape[2] == n_sample or lazy:
            return torch.arange(n_sample, dtype=torch.int32, device=xyz.device).view(1, -1).expand(xyz.shape[0], -1).contiguous()
        return grp.furthest_point_sampling(xyz.contiguous(), n_sample)

    O.ball_query = ball_query
    O.furthest_sample_index = furthest_sample_index
    base_kp = C.kernel_points_base()
    orig_skp = O.scaled_kernel_points

    params = O.backbone_params(input_num=1024)
    sd = {k: v.to(dev).requires_grad_(True) for k, v in O.init_backbone_state(params, seed=0).items()}
    for bi, blk in enumerate(params):
        for li, layer in enumerate(blk):
            co = layer['args']['dim_out']
            for pre in (f'backbone.{bi}.blocks.{li}.inter_conv.norm.', f'backbone.{bi}.blocks.{li}.norm.'):
                sd[pre + 'running_mean'], sd[pre + 'running_var'] = torch.zeros(co, device=dev), torch.ones(co, device=dev)
    leaves = [v for v in sd.values() if v.requires_grad]
    opt = torch.optim.Adam(leaves, lr=1e-3, fused=True)
    anchors = torch.from_numpy(C.anchors_all()).to(dev)
    intra = torch.from_numpy(C.intra_idx()).to(dev)
    xyz = O.synthetic_cloud(args.clouds, 1024, 2000).permute(0, 2, 1).contiguous().to(dev)
    feats = torch.ones(args.clouds, 1, 1024, 60, device=dev)

    def step():
        opt.zero_grad(set_to_none=True)
        _, of = O.backbone_forward(sd, params, xyz, feats, anchors, intra, base_kp, training=True)
        loss = of.square().mean()
        loss.backward()
        opt.step()
        return loss

    # same as oracle.so3.inter_block, with the kernel points created on the device
    def inter_block(sd_, prefix, a, xyz_, feats_, anchors_, base_kp_, training=True):
        kernels = torch.from_numpy(orig_skp(base_kp_, a['radius'])).to(feats_.device, feats_.dtype)
        gxyz, idx, sidx, new_xyz = O.ball_grouping(xyz_, a['stride'], a['radius'], a['n_neighbor'], a.get('lazy_sample', True))
        w = O.anchor_weights(gxyz, anchors_, kernels, a['sigma'])
        G = O.inter_group_feats(idx, w, feats_)
        y = O.basic_conv(sd_[prefix + 'conv.basic_conv.W'], G)
        y = torch.nn.functional.leaky_relu(O._bn(y, sd_, prefix + 'norm.', training))
        return idx, w, sidx, new_xyz, y
    O.inter_block = inter_block

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"what": "reference-style GPU path: reference CUDA kernels (recompiled, sm_100a) + eager torch fp32",
                      "points_per_s": args.clouds * 1024 / (ms * 1e-3), "ms_per_step": ms, "clouds": args.clouds,
                      "allow_tf32": bool(args.tf32), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                      "loss": float(loss)}))


if __name__ == "__main__":
    main()
