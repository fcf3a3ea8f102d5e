"""Developer check (GPU): per-parameter gradient / output error of the classic backbone vs the oracle, per GEMM mode."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.test_gpu_parity import _oracle_case, _oracle_run
from tests.helpers import rel_err
from equi_articulated_pose_b200 import ops
dev = torch.device("cuda:0")
n, b, kind = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
sd64 = _oracle_run(n, b, 11, torch.float64, kind)[3]
for mode in [int(m) for m in sys.argv[4:]] or [0, 1, 2]:
    ops.set_gemm_mode(mode)
    sdo, of, oxyz, loss, net, out, l2 = _oracle_case(dev, n, b, 11, kind)
    print(f"mode {mode}: out rel err {rel_err(out.feats, of):.2e} loss rel {abs(float(l2)-float(loss))/abs(float(loss)):.2e}")
    for name, p in net.named_parameters():
        t = sd64[name].grad; sc = float(t.abs().max()) + 1e-30
        print(f"   {name:58s} gpu-vs-fp64 {float((p.grad.double().cpu()-t).abs().max())/sc:.2e}  ref32-vs-fp64 {float((sdo[name].grad.double()-t).abs().max())/sc:.2e}  (max {sc:.1e})")
