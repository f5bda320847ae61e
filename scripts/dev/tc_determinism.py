"""Developer probe: bitwise run-to-run determinism of the GEMM engines."""
import sys
import torch
sys.path.insert(0, ".")
import swarm_ode_b200 as S
from swarm_ode_b200 import _lib

dev = torch.device("cuda:0")
for engine in ("tc", "simt"):
    S.set_engine(engine)
    for (m, n, k) in ((389120, 128, 399), (389120, 399, 128), (389120, 64, 128), (50000, 128, 64)):
        a = torch.randn(m, k, device=dev)
        w = torch.randn(n, k, device=dev) / k ** 0.5
        ref = S.ops.gemm_nt(a, w).clone()
        bad, worst = 0, 0.0
        for i in range(10):
            out = S.ops.gemm_nt(a, w)
            d = (out - ref).abs().max().item()
            if d != 0.0:
                bad += 1
                worst = max(worst, d)
                rows = ((out - ref).abs().amax(dim=1) > 0).nonzero().flatten()
                cols = ((out - ref).abs().amax(dim=0) > 0).nonzero().flatten()
                if bad == 1:
                    print("   first mismatch rows", rows[:8].tolist(), "n_rows", rows.numel(), "cols", cols[:8].tolist(), "n_cols", cols.numel())
        print(f"{engine} {m}x{n}x{k}: {bad}/10 runs differ, worst abs diff {worst:.3e} (ref max {ref.abs().max().item():.2f})")
_lib.tc_check(dev)
