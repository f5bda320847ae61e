"""Developer probe: random shapes through the row-major K = 128 engine (gnode_gemm_k128 with one base term) against float64."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import swarm_ode_b200 as S
from swarm_ode_b200 import _lib
dev = torch.device("cuda:0")
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
worst = 0.0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    m = int(rng.choice([rng.integers(1, 300), rng.integers(300, 40000), 128 * 148 * int(rng.integers(1, 4)) + int(rng.integers(0, 200))]))
    n = int(rng.choice([399, 400, rng.integers(16, 401)]))
    g = torch.Generator().manual_seed(it)
    a = torch.randn(m, 128, generator=g); w = torch.randn(n, 128, generator=g) * 0.2
    base = torch.randn(m, n, generator=g) * 3.0
    bias = torch.randn(n, generator=g) if it % 2 else None
    sc, bs, bsc = float(rng.uniform(0.1, 2.0)), float(rng.uniform(-1, 1)), float(rng.uniform(0.5, 1.5))
    want = bsc * base.double() + sc * (a.double() @ w.double().T + (bs * bias.double() if bias is not None else 0.0))
    got = S.ops.gemm_k128(a.to(dev), w.to(dev), bias=None if bias is None else bias.to(dev), bias_scale=bs, base=base.to(dev),
                          base_scale=bsc, scale=sc)
    _lib.tc_check(dev)
    err = (got.double().cpu() - want).abs().max().item() / want.abs().max().item()
    worst = max(worst, err)
    assert err <= 2e-5, (it, m, n, err)
print("ok, worst max-error / max-value:", worst)
