"""dev: which evaluation order does torch.sum / torch.norm use on THIS host, and does the knn kernel match it."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from equi_articulated_pose_b200 import ops
torch.manual_seed(0)
print("cpu capability:", torch.backends.cpu.get_cpu_capability(), "threads", torch.get_num_threads())
b, n, s, k = 3, 700, 50, 64
g = torch.Generator().manual_seed(800)
pos = torch.rand(b, n, 3, generator=g) - 0.5
cen = pos[:, torch.randperm(n, generator=g)[:s]].contiguous()
d = (cen.unsqueeze(2) - pos.unsqueeze(1)) ** 2
ssum = torch.sum(d, dim=-1).numpy()
dn = d.numpy()
for name, o in (("(0+1)+2", (dn[..., 0] + dn[..., 1]) + dn[..., 2]), ("0+(1+2)", dn[..., 0] + (dn[..., 1] + dn[..., 2])),
                ("(0+2)+1", (dn[..., 0] + dn[..., 2]) + dn[..., 1])):
    print("torch.sum order", name, "mismatches", int((ssum != o).sum()))
full = torch.sqrt(torch.sum(d, dim=-1))
rd, ri = torch.topk(full, k=k, dim=2, largest=False)
idx, dist = ops.knn_query(pos.cuda(), cen.cuda(), k)
idx, dist = idx.cpu().long(), dist.cpu()
print("dist mismatches", int((dist != rd).sum()), "of", rd.numel(), "max abs", float((dist - rd).abs().max()))
print("sorted ascending:", bool((dist[..., 1:] >= dist[..., :-1]).all()))
print("idx mismatches", int((idx != ri).sum()))
o1 = np.sqrt((dn[..., 0] + dn[..., 1]) + dn[..., 2])
print("gathered o1 at idx == dist:", bool((np.take_along_axis(o1, idx.numpy(), 2) == dist.numpy()).all()))
print("first row dist", dist[0, 0, :6].tolist(), "ref", rd[0, 0, :6].tolist())
print("first row idx", idx[0, 0, :6].tolist(), "ref", ri[0, 0, :6].tolist())
