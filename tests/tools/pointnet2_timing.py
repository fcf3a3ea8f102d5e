"""dev/measurement: PointnetPP (SPConvNets/models/PointNet2.py) fwd+bwd on the sm_100a kernels, B clouds of N points.
Prints one JSON object: ms per step (CUDA events, L2 flushed between steps) and the per-entry-point CUDA-event table.
    python tests/tools/pointnet2_timing.py [B] [N] [steps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from equi_articulated_pose_b200 import lib
from equi_articulated_pose_b200.pointnet2 import PointnetPP


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    dev = torch.device("cuda:0")
    lib.load()
    torch.manual_seed(0)
    net = PointnetPP(6).to(dev).train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, fused=True)
    g = torch.Generator().manual_seed(1)
    pos = (torch.rand(b, n, 3, generator=g) - 0.5).to(dev)
    x = torch.randn(b, n, 3, generator=g).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        opt.zero_grad(set_to_none=True)
        out, glb, _ = net(x, pos, return_global=True)
        loss = out.square().mean() + glb.square().mean()
        loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        step()
        flush.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lib.PROFILE = []
    e0.record()
    for _ in range(steps):
        step()
        flush.zero_()
    e1.record()
    torch.cuda.synchronize()
    rec, lib.PROFILE = lib.PROFILE, None
    ms = e0.elapsed_time(e1) / steps
    table = {}
    for name, a, s0, s1 in rec:
        d = table.setdefault(name, [0.0, 0])
        d[0] += s0.elapsed_time(s1)
        d[1] += 1
    table = {k: {"ms_per_step": v[0] / steps, "calls_per_step": v[1] / steps} for k, v in sorted(table.items(), key=lambda kv: -kv[1][0])}
    # knn_query roofline: algorithmic bytes = cloud + centres read, (idx, dist) written
    print(json.dumps({"workload": f"PointnetPP(in_feat_dim=6) fwd+bwd+Adam, {b} clouds x {n} points, fp32 (3xTF32 contractions)",
                      "ms_per_step": ms, "points_per_s": b * n / (ms * 1e-3), "kernel_table": table}))


if __name__ == "__main__":
    main()
