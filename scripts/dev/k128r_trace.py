"""Developer probe: phase timestamps of the third block of CTA 0 in the row-major K = 128 engine (needs -DK128_TRACE)."""
import ctypes as C, sys
import torch
sys.path.insert(0, ".")
import swarm_ode_b200 as S
from swarm_ode_b200 import _lib
dev = torch.device("cuda:0")
m, n = 389120, 399
a = torch.randn(m, 128, device=dev); w = torch.randn(n, 128, device=dev); base = torch.randn(m, n, device=dev); bias = torch.randn(n, device=dev)
for _ in range(3):
    out = S.ops.gemm_k128(a, w, bias=bias, base=base)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = S.ops.gemm_k128(a, w, bias=bias, base=base)
e1.record(); torch.cuda.synchronize()
print("ms per call (incl. image pack):", e0.elapsed_time(e1) / 5)
lib = C.CDLL(_lib.LIB_PATH)
buf = (C.c_longlong * 128)()
lib.gnode_k128r_trace(buf)
v = list(buf)
b0 = min(v[8 * q] for q in range(4))
us = lambda x: (x - b0) / 1.965e3
print("quadrant: loop top | acc_full | A fetch issued | slots full | RMW done+arrive | next tile published")
for q in range(4):
    print(q, " ".join(f"{us(x):7.2f}" for x in v[8 * q: 8 * q + 6]))
