"""dev: classic backbone in contraction mode 4 (single-pass bf16) at a tiny size, weight planes off / on, launch-blocking."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from equi_articulated_pose_b200 import blocks, lib, ops, synthetic  # noqa: E402

lib.load()
dev = torch.device("cuda:0")
ops.set_gemm_mode(int(sys.argv[1]) if len(sys.argv) > 1 else 4)
npts = int(sys.argv[2]) if len(sys.argv) > 2 else 256
params = blocks.backbone_params(input_num=npts)
clouds = synthetic.synthetic_cloud(2, npts, 77).to(dev)
for flag in (("1",) if os.environ.get("ONLY_PLANES") else ("0", "1")):
    os.environ["VGTKB_WEIGHT_PLANES"] = flag
    net = blocks.SO3Backbone(params)
    net.load_state_dict(synthetic.init_backbone_state(params, seed=0), strict=False)
    net = net.to(dev).train()
    try:
        out = net(clouds)
        torch.cuda.synchronize()
        print("planes", flag, "forward ok", float(out.feats.abs().max()), flush=True)
        out.feats.square().mean().backward()
        torch.cuda.synchronize()
        print("planes", flag, "backward ok", flush=True)
    except Exception as e:  # noqa: BLE001
        print("planes", flag, "FAILED:", repr(e)[:300], flush=True)
        import traceback
        traceback.print_exc()
        break
