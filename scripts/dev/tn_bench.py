"""Developer probe: time the weight-gradient (TN) contraction on the bench shapes, both engines."""
import sys
import torch
sys.path.insert(0, ".")
import swarm_ode_b200 as S
from swarm_ode_b200 import _lib

M = int(sys.argv[1]) if len(sys.argv) > 1 else 389120
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
for engine in ("tc", "simt"):
    S.set_engine(engine)
    for (p, q) in ((399, 128), (128, 399), (64, 128)):
        a = torch.randn(M, p, device=dev)
        b = torch.randn(M, q, device=dev)
        out = torch.zeros(p, q, device=dev)
        for _ in range(2):
            S.ops.gemm_tn(a, b, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            S.ops.gemm_tn(a, b, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{engine} rows={M} P={p} Q={q}: {ms:.3f} ms  {2.0*M*p*q/ms/1e9:.1f} TFLOP/s  {4.0*M*(p+q)/ms/1e6:.0f} GB/s", flush=True)
_lib.tc_check(dev)
