"""The two GEMM engines (fp32 FFMA anchor, tcgen05 3xTF32) against a float64 reference."""
import pytest
import torch

import swarm_ode_b200 as S
from swarm_ode_b200 import _lib
from tests._util import rel_l2

pytestmark = pytest.mark.gpu

SHAPES = [(128, 128, 399), (1000, 128, 399), (300, 399, 128), (4096 + 37, 64, 128), (95, 128, 64), (257, 16, 8),
          (513, 399, 128), (64, 435, 128), (2000, 128, 435), (1, 32, 5)]


@pytest.mark.parametrize("engine", ["simt", "tc"])
@pytest.mark.parametrize("m,n,k", SHAPES)
def test_gemm_nt_plain(cuda, engine, m, n, k):
    if engine == "tc" and m < 4:
        pytest.skip("the tcgen05 engine takes >= 4 rows; AUTO falls back to FFMA below that")
    torch.manual_seed(m + n + k)
    a = torch.randn(m, k) * 3
    w = torch.randn(n, k) / k ** 0.5
    want = a.double() @ w.double().t()
    prev = S.set_engine(engine)
    try:
        got = S.ops.gemm_nt(a.to(cuda), w.to(cuda))
        _lib.tc_check(cuda)
    finally:
        S.set_engine(prev)
    err = rel_l2(got, want)
    print(f"{engine} {m}x{n}x{k}: rel-L2 {err:.3e}")
    assert err <= (1e-5 if engine == "tc" else 2e-6)


@pytest.mark.parametrize("engine", ["simt", "tc"])
@pytest.mark.parametrize("act", ["none", "relu", "tanh"])
def test_gemm_nt_epilogue(cuda, engine, act):
    torch.manual_seed(3)
    m, n, k = 777, 399, 128
    a, w, b = torch.randn(m, k), torch.randn(n, k) / k ** 0.5, torch.randn(n)
    base = torch.randn(m, n)
    v = a.double() @ w.double().t() + b.double()
    v = {"none": v, "relu": torch.relu(v), "tanh": torch.tanh(v)}[act]
    want = base.double() + 0.37 * v
    prev = S.set_engine(engine)
    try:
        got = S.ops.gemm_nt(a.to(cuda), w.to(cuda), bias=b.to(cuda), act=act, base=base.to(cuda), scale=0.37)
        _lib.tc_check(cuda)
    finally:
        S.set_engine(prev)
    assert rel_l2(got, want) <= (1e-5 if engine == "tc" else 2e-6)


def test_tc_engine_matches_ffma_on_warehouse_magnitudes(cuda):
    """Un-normalised coordinates (values up to ~35) times small weights: the 3xTF32 split must stay fp32-grade."""
    batch, _ = S.synthetic.warehouse_batch(32, seed=0)
    x = batch.x.to(cuda)
    w = (torch.rand(128, 399, generator=torch.Generator().manual_seed(1)) * 2 - 1) / 399 ** 0.5
    want = batch.x.double() @ w.double().t()
    errs = {}
    for engine in ("simt", "tc"):
        prev = S.set_engine(engine)
        try:
            errs[engine] = rel_l2(S.ops.gemm_nt(x, w.to(cuda)), want)
            _lib.tc_check(cuda)
        finally:
            S.set_engine(prev)
    print(errs)
    assert errs["tc"] <= 1e-5 and errs["simt"] <= 1e-6
