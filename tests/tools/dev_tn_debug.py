import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from equi_articulated_pose_b200 import ops
dev = torch.device("cuda:0")
torch.set_printoptions(linewidth=200, precision=1, sci_mode=False)
R, M, N = 64, 64, 128
def run(A, B, name):
    out = ops.gemm_tn(A.to(dev), B.to(dev), mode=2).cpu()
    ref = A.double().t() @ B.double()
    print(name, "max err", float((out - ref).abs().max()), "out[0,:8]", out[0, :8].tolist(), "out[:8,0]", out[:8, 0].tolist(),
          "ref[0,:8]", ref[0, :8].tolist(), "nonzero", int((out != 0).sum()), flush=True)
    return out
ones_A, ones_B = torch.ones(R, M), torch.ones(R, N)
run(ones_A, ones_B, "ones")
run(ones_A, torch.arange(N).float().repeat(R, 1), "B=i")
run(torch.arange(M).float().repeat(R, 1), ones_B, "A=j")
B = torch.zeros(R, N); B[3] = 1
run(ones_A, B, "B row3")
A = torch.zeros(R, M); A[:, 5] = 1
run(A, ones_B, "A col5")
