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


# rows, p, q -- the three weight-gradient shapes of a GraphODEFunc layer stack (D=399 / 435, 2H=128, H=64) plus
# ragged row counts (tail rows go through the FFMA kernel), narrow padding (p, q < 128) and a wide N operand.
TN_SHAPES = [(4096, 399, 128), (4096, 128, 399), (4096, 64, 128), (12345, 435, 128), (1003, 128, 435),
             (64, 128, 64), (9, 16, 24), (8, 399, 128), (50000, 512, 96), (777, 32, 64), (2048, 128, 128)]


@pytest.mark.parametrize("engine", ["simt", "tc"])
@pytest.mark.parametrize("rows,p,q", TN_SHAPES)
def test_gemm_tn(cuda, engine, rows, p, q):
    torch.manual_seed(rows + p + q)
    a = torch.randn(rows, p) * 2
    b = torch.randn(rows, q)
    c0 = torch.randn(p, q)
    want = c0.double() + 0.5 * (a.double().t() @ b.double())
    prev = S.set_engine(engine)
    try:
        got = S.ops.gemm_tn(a.to(cuda), b.to(cuda), out=c0.to(cuda).clone(), scale=0.5)
        _lib.tc_check(cuda)
    finally:
        S.set_engine(prev)
    err = rel_l2(got, want)
    print(f"{engine} tn rows={rows} {p}x{q}: rel-L2 {err:.3e}")
    assert err <= (1e-5 if engine == "tc" else 2e-6)


def test_gemm_tn_deterministic(cuda):
    torch.manual_seed(0)
    a, b = torch.randn(30000, 399, device=cuda), torch.randn(30000, 128, device=cuda)
    outs = [S.ops.gemm_tn(a, b) for _ in range(3)]
    _lib.tc_check(cuda)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
