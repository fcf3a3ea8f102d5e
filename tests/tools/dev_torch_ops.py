"""Developer check (GPU): which torch (non-vgtkb) kernels run inside one backbone step, with the calling op."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from equi_articulated_pose_b200 import blocks, dataparallel as dp
from oracle import so3 as O
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda:0")
params = O.backbone_params(input_num=1024)
net = blocks.SO3Backbone(params)
net.load_state_dict(O.init_backbone_state(params, seed=0), strict=False)
net = net.to(dev).train()
bucket = dp.FlatGradBucket(net.parameters())
opt = torch.optim.Adam(bucket.params, lr=1e-3, fused=True)
pts = O.synthetic_cloud(8, 1024, 2000).to(dev)

def step():
    bucket.zero_()
    out = net(pts)
    loss = out.feats.square().mean()
    loss.backward()
    opt.step()

for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=False) as prof:
    step()
    torch.cuda.synchronize()
from torch.autograd import DeviceType
rows = []
for e in prof.key_averages():
    if e.device_type == DeviceType.CUDA:
        rows.append((e.device_time_total, e.count, e.key))
rows.sort(reverse=True)
mine = sum(t for t, c, k in rows if "vgtkb" in k)
other = [(t, c, k) for t, c, k in rows if "vgtkb" not in k]
print(f"vgtkb kernels {mine:.0f} us, other kernels {sum(t for t, c, k in other):.0f} us")
for t, c, k in other[:25]:
    print(f"{t:9.1f} us  n={c:3d}  {k[:150]}")
# which autograd / aten ops launch them
ops = []
for e in prof.key_averages(group_by_input_shape=True):
    if e.device_type == DeviceType.CPU and e.self_device_time_total > 0 and not e.key.startswith("vgtkb"):
        ops.append((e.self_device_time_total, e.count, e.key, str(e.input_shapes)[:100]))
ops.sort(reverse=True)
print("---- aten ops by self device time")
for t, c, k, sh in ops[:30]:
    print(f"{t:9.1f} us  n={c:3d}  {k[:40]:40s} {sh}")
