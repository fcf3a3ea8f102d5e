"""dev: which part of the step breaks CUDA-graph capture?"""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from equi_articulated_pose_b200 import lib, blocks, synthetic, dataparallel as dp, ops
lib.load()
dev = torch.device("cuda:0")
params = blocks.backbone_params(input_num=1024)
net = blocks.SO3Backbone(params)
net.load_state_dict(synthetic.init_backbone_state(params, seed=0), strict=False)
net = net.to(dev).train()
bucket = dp.FlatGradBucket(net.parameters())
opt = torch.optim.Adam(bucket.params, lr=1e-3, fused=True, capturable=True)
pts = synthetic.synthetic_cloud(8, 1024, 2000).to(dev)

def fwd():
    with torch.no_grad():
        return net(pts).feats.square().mean()
def fwd_bwd():
    bucket.zero_()
    loss = net(pts).feats.square().mean()
    loss.backward()
    return loss
def full():
    l = fwd_bwd()
    bucket.all_reduce_mean()
    opt.step()
    return l

# trace every C-ABI call made while capturing so the offender can be named
orig_call = lib.call
trace = []
def traced(name, device, *a):
    trace.append(name)
    return orig_call(name, device, *a)

for label, fn in (("fwd", fwd), ("fwd_bwd", fwd_bwd), ("full", full)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ops.clear_planes()
    g = torch.cuda.CUDAGraph()
    trace.clear()
    lib.call = traced; ops.call = traced
    try:
        with torch.cuda.graph(g):
            out = fn()
        g.replay(); torch.cuda.synchronize()
        print(label, "captured OK; loss", float(out), flush=True)
    except Exception as e:
        print(label, "FAILED:", str(e).splitlines()[0], "| last C-ABI calls:", trace[-4:], "n calls", len(trace), flush=True)
        # find the first call after which the stream capture is invalid
        break
    finally:
        lib.call = orig_call; ops.call = orig_call
