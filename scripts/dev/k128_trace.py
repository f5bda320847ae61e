"""Developer probe: phase timestamps of one tile of the K = 128 wide-output engine (needs -DK128_TRACE)."""
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
lib.gnode_k128_trace(buf)
v = list(buf)
b0 = v[0]
us = lambda x: (x - b0) / 1.965e3
print("A load", f"{us(v[1]):.2f}", "residual+arrive", f"{us(v[2]):.2f}")
for c in range(5):
    row = v[3 + 8 * c: 3 + 8 * c + 9]
    print("chunk", c, " ".join(f"{us(x):7.2f}" for x in row))
