"""dev: N eager training steps of the bench workload (for ncu captures: `ncu ... python tests/tools/one_step.py 4`)."""
import os, sys
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
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    bucket.zero_()
    loss = net(pts).feats.square().mean()
    loss.backward()
    bucket.all_reduce_mean()
    opt.step()
torch.cuda.synchronize()
print("done", float(loss))
