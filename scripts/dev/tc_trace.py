"""Developer probe: where the roles of k_gemm_tc (CTA 0) wait, for the Z_0 shape (needs a -DTC_TRACE build)."""
import ctypes as C, sys
import torch
sys.path.insert(0, ".")
import swarm_ode_b200 as S
from swarm_ode_b200 import _lib
dev = torch.device("cuda:0")
m, k, n = 389120, 399, 128
a = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev)
lib = C.CDLL(_lib.LIB_PATH)
buf = (C.c_longlong * 16)()
for _ in range(3):
    S.ops.gemm_nt(a, w)
torch.cuda.synchronize()
lib.gnode_tc_trace(buf, 1)
reps = 5
for _ in range(reps):
    S.ops.gemm_nt(a, w)
torch.cuda.synchronize()
lib.gnode_tc_trace(buf, 1)
v = [x / reps / 1.965e3 for x in buf]
names = {0: "MMA waits weights (b_full)", 1: "MMA waits A operand (aop_full)", 2: "MMA waits accumulator (acc_empty)",
         4: "converter waits raw (raw_full)", 5: "converter waits operand slot (aop_empty)", 7: "A producer waits raw slot",
         8: "epilogue waits accumulator (acc_full)", 9: "B producer waits slot (b_empty)", 10: "MMA warp: issuing the 6 MMAs", 11: "MMA warp: issuing the commits", 15: "kernel (CTA 0)"}
for i, nme in names.items():
    print(f"{nme:45s} {v[i]:9.1f} us")
