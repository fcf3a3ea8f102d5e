"""dev: how much of a bench step is host-side enqueue time?  (decides whether a CUDA graph is worth it)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from equi_articulated_pose_b200 import lib, blocks, synthetic, dataparallel as dp
lib.load()
dev = torch.device("cuda:0")
params = blocks.backbone_params(input_num=1024)
net = blocks.SO3Backbone(params)
net.load_state_dict(synthetic.init_backbone_state(params, seed=0), strict=False)
net = net.to(dev).train()
bucket = dp.FlatGradBucket(net.parameters())
opt = torch.optim.Adam(bucket.params, lr=1e-3, fused=True, capturable=True)
pts = synthetic.synthetic_cloud(8, 1024, 2000).to(dev)
def step():
    bucket.zero_()
    out = net(pts)
    loss = out.feats.square().mean()
    loss.backward()
    bucket.all_reduce_mean()
    opt.step()
    return loss.detach()
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"enqueue {1e3*(t1-t0)/10:.2f} ms/step, total {1e3*(t2-t0)/10:.2f} ms/step")
# forward only / backward only host time
torch.cuda.synchronize(); t0 = time.perf_counter(); out = net(pts); loss = out.feats.square().mean(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"fwd enqueue {1e3*(t1-t0):.2f} ms, fwd total {1e3*(t2-t0):.2f}")
t0 = time.perf_counter(); loss.backward(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"bwd enqueue {1e3*(t1-t0):.2f} ms, bwd total {1e3*(t2-t0):.2f}")
del out, loss      # a live loss keeps the autograd graph (and its AccumulateGrad nodes, bound to the default stream) alive

# ---- the same step as a CUDA graph
from equi_articulated_pose_b200 import graph as G
l_eager = float(step())
cs = G.CapturedStep(lambda p: step(), [pts])
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(10): out = cs(pts)
ev1.record(); torch.cuda.synchronize()
print(f"graph replay {ev0.elapsed_time(ev1)/10:.2f} ms/step, loss {float(out):.6f} (eager before capture {l_eager:.6f})")
