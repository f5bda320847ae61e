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


@pytest.mark.parametrize("m,n", [(1, 16), (125, 343), (128, 399), (129, 80), (1000, 81), (3000, 400), (257, 435), (40000, 399)])
@pytest.mark.parametrize("terms", ["plain", "bias+base", "all"])
def test_k128_wide_output_engine(cuda, m, n, terms):
    """csrc/gemm_k128.cu: C = base_scale * base + base2 + scale * (A @ W^T + bias_scale * bias), K = 128, any N >= 16, any
    row count, rows of C / base only 4-byte aligned -- against float64, fp32-grade accuracy (three-term tensor-core product)."""
    g = torch.Generator().manual_seed(m * 7 + n)
    # poison the device heap first: the engine must not depend on fresh (zero) memory
    junk = torch.full((8 << 20,), float("nan"), device=cuda)
    del junk
    a = torch.randn(m, 128, generator=g) * 3.0
    w = torch.randn(n, 128, generator=g) * 0.3
    bias = torch.randn(n, generator=g) if terms != "plain" else None
    base = torch.randn(m, n, generator=g) * 5.0 if terms != "plain" else None
    base2 = torch.randn(m, n, generator=g) if terms == "all" else None
    scale, bs, bsc = (0.37, 1.7, 0.5) if terms == "all" else (1.0, 1.0, 1.0)
    want = scale * (a.double() @ w.double().T + (bs * bias.double() if bias is not None else 0.0))
    if base is not None:
        want = want + bsc * base.double()
    if base2 is not None:
        want = want + base2.double()
    dev = lambda t: None if t is None else t.to(cuda)
    got = S.ops.gemm_k128(dev(a), dev(w), bias=dev(bias), bias_scale=bs, base=dev(base), base_scale=bsc, base2=dev(base2), scale=scale)
    assert torch.isfinite(got).all()
    assert rel_l2(got, want) <= 5e-6, rel_l2(got, want)
    from swarm_ode_b200 import _lib
    _lib.tc_check(cuda)


@pytest.mark.parametrize("m,n", [(16, 399), (100, 399), (128 * 148 + 5, 399), (128 * 148 * 3 + 16 * 3 + 9, 399), (5000, 321), (5000, 64),
                                 (5000, 435), (128 * 148 + 21, 440), (3000, 401), (777, 432)])
def test_k128_row_major_engine(cuda, m, n):
    """Row-major variant (k_gemm_k128_rows: dense rows, at most one base term; 80-column chunks for N <= 400, 88-column
    chunks for N <= 440 -- the y_1 / error / dense-output projections): whole 16-row spans through shared memory, a ragged
    last group updated in global memory, several blocks per CTA; element-wise against float64, with and without a base."""
    g = torch.Generator().manual_seed(m + n)
    junk = torch.full((8 << 20,), float("nan"), device=cuda)
    del junk
    a = torch.randn(m, 128, generator=g) * 2.0
    w = torch.randn(n, 128, generator=g) * 0.2
    base = torch.randn(m, n, generator=g) * 4.0
    bias = torch.randn(n, generator=g)
    for kw, want in (
        (dict(base_scale=1.0, scale=0.25), base.double() + 0.25 * (a.double() @ w.double().T)),
        (dict(bias=bias.to(cuda), bias_scale=0.6, base_scale=-0.5, scale=1.5),
         -0.5 * base.double() + 1.5 * (a.double() @ w.double().T + 0.6 * bias.double())),
    ):
        got = S.ops.gemm_k128(a.to(cuda), w.to(cuda), base=base.to(cuda), **kw)
        _lib.tc_check(cuda)
        assert torch.isfinite(got).all()
        assert rel_l2(got, want) <= 5e-6, rel_l2(got, want)
        err = (got.double().cpu() - want).abs().max().item()
        assert err <= 2e-4, err          # no misplaced row / column anywhere (values are O(10))
    # no base term: the span slots are written, not updated (they hold NaN from the poisoned heap or the previous call)
    got = S.ops.gemm_k128(a.to(cuda), w.to(cuda), bias=bias.to(cuda), bias_scale=-1.25, scale=0.75)
    _lib.tc_check(cuda)
    want = 0.75 * (a.double() @ w.double().T - 1.25 * bias.double())
    assert torch.isfinite(got).all()
    assert rel_l2(got, want) <= 5e-6, rel_l2(got, want)
    assert (got.double().cpu() - want).abs().max().item() <= 2e-4
