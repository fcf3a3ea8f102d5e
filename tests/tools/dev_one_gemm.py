import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from equi_articulated_pose_b200 import ops
M, N, K = [int(v) for v in sys.argv[1:4]]
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = torch.device("cuda:0")
A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev)
for _ in range(3):
    out = ops.gemm_nt(A, B, None, mode=mode)
torch.cuda.synchronize()
print(float(out[0, 0]))
