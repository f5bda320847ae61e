"""Developer probe: one tcgen05 GEMM per process (a CUDA fault kills the context), prints the error."""
import sys
import torch
sys.path.insert(0, ".")
import swarm_ode_b200 as S
from swarm_ode_b200 import _lib

m, n, k = (int(v) for v in sys.argv[1:4])
torch.manual_seed(0)
a = torch.randn(m, k) * 3
w = torch.randn(n, k) / k ** 0.5
want = a.double() @ w.double().t()
S.set_engine("tc")
try:
    got = S.ops.gemm_nt(a.cuda(), w.cuda())
    _lib.tc_check(torch.device("cuda:0"))
    err = float((got.cpu().double() - want).norm() / want.norm())
    print(f"tc {m}x{n}x{k}: rel-L2 {err:.3e}")
except Exception as e:
    print(f"tc {m}x{n}x{k}: FAILED {str(e)[:300]}")
