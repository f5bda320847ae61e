"""Developer probe: time the tcgen05 GEMM on the two D-wide shapes of the bench (config 2)."""
import sys
import torch
sys.path.insert(0, ".")
import swarm_ode_b200 as S
from swarm_ode_b200 import _lib

M = int(sys.argv[1]) if len(sys.argv) > 1 else 389120
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
S.set_engine("tc")
dev = torch.device("cuda:0")
for (n, k, base) in ((128, 399, False), (399, 128, False), (399, 128, True), (64, 128, False)):
    a = torch.randn(M, k, device=dev)
    w = torch.randn(n, k, device=dev) / k ** 0.5
    b = torch.randn(M, n, device=dev) if base else None
    for _ in range(2):
        S.ops.gemm_nt(a, w, base=b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        S.ops.gemm_nt(a, w, base=b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"M={M} N={n} K={k} base={base}: {ms:.3f} ms  {2.0*M*n*k/ms/1e9:.1f} TFLOP/s  "
          f"{4.0*M*(n+k)/ms/1e6:.0f} GB/s", flush=True)
_lib.tc_check(dev)
