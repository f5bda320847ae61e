"""Cross-check of the oracle against the REAL third-party libraries wherever they import.

The arithmetic of the reference path lives in ``torchdiffeq`` and ``torch_geometric`` (import sites:
scripts/train_gde.py:8-11), which the reference neither vendors nor pins and which are not installable in the build
container (no network, no wheel).  ``oracle/`` restates their published algorithms; this module is the way out of "parity
unpinned": on any machine where the libraries DO import, every check below compares the restatement with the library
itself.  Each check exists twice -- unmarked (CPU suite) and ``-m gpu`` (so that the GPU box's record shows it was tried
there as well); where a library is absent the skip reason says so.
"""
import importlib

import pytest
import torch

from oracle import torchdiffeq_ref as T
from oracle.pyg_ref import HeteroConvRef, RefBatch, RefData, SAGEConvRef


def _need(name):
    try:
        return importlib.import_module(name)
    except Exception as e:  # ImportError, or a broken binary wheel
        pytest.skip(f"real `{name}` is not importable on this machine ({type(e).__name__}: {e}); "
                    "the oracle's restatement of it stays unpinned here (tried, see DESIGN.md 2)")


class _Field(torch.nn.Module):
    """A small smooth nonlinear field with a call counter."""

    def __init__(self, dim=6, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.w1 = torch.nn.Parameter(torch.randn(dim, 16, generator=g) * 0.4)
        self.w2 = torch.nn.Parameter(torch.randn(16, dim, generator=g) * 0.4)
        self.nfe = 0

    def forward(self, t, y):
        self.nfe += 1
        return torch.tanh(y @ self.w1) @ self.w2 - 0.1 * y * (1.0 + t)


# ---------------------------------------------------------------------------------------------- torchdiffeq
def _check_fixed_solvers():
    tde = _need("torchdiffeq")
    y0 = torch.randn(5, 6, generator=torch.Generator().manual_seed(1))
    t = torch.tensor([0.0, 0.3, 0.7, 1.5])
    for method in ("euler", "midpoint", "rk4"):
        f = _Field()
        want = tde.odeint(f, y0, t, method=method)
        n_real = f.nfe
        f.nfe = 0
        st = T.SolverStats()
        got = T.odeint_ref(f, y0, t, method=method, stats=st)
        assert torch.equal(got, want), (method, float((got - want).abs().max()))     # same operations in the same order
        assert st.nfe == f.nfe == n_real


def _check_dopri5():
    tde = _need("torchdiffeq")
    y0 = torch.randn(5, 6, generator=torch.Generator().manual_seed(2)) * 2.0
    for t, rtol, atol in ((torch.tensor([0.0, 1.0]), 1e-3, 1e-4), (torch.linspace(0, 4, 9), 1e-5, 1e-6),
                          (torch.tensor([0.0, 0.05, 3.0]), 1e-3, 1e-4)):
        f = _Field(seed=3)
        want = tde.odeint(f, y0, t, method="dopri5", rtol=rtol, atol=atol)
        n_real = f.nfe
        f.nfe = 0
        st = T.SolverStats()
        got = T.odeint_ref(f, y0, t, method="dopri5", rtol=rtol, atol=atol, stats=st)
        assert st.nfe == n_real, "the restated controller took a different number of field evaluations"
        assert torch.allclose(got, want, rtol=1e-6, atol=1e-7), float((got - want).abs().max())
    # default method and tolerances (scripts/gnode.py:136-137)
    f = _Field(seed=4)
    want = tde.odeint(f, y0.double(), torch.tensor([0.0, 1.0], dtype=torch.float64))
    n_real, f.nfe = f.nfe, 0
    st = T.SolverStats()
    got = T.odeint_ref(f, y0.double(), torch.tensor([0.0, 1.0], dtype=torch.float64), stats=st)
    assert st.nfe == n_real and torch.allclose(got, want, rtol=1e-9, atol=1e-10)


def _check_dopri5_gradients():
    tde = _need("torchdiffeq")
    y0 = torch.randn(4, 6, generator=torch.Generator().manual_seed(5))
    t = torch.tensor([0.0, 0.5, 2.0])
    grads = []
    for solve in (lambda f: tde.odeint(f, y0, t, method="dopri5", rtol=1e-3, atol=1e-4),
                  lambda f: T.odeint_ref(f, y0, t, method="dopri5", rtol=1e-3, atol=1e-4)):
        f = _Field(seed=6)
        (solve(f)[-1] ** 2).sum().backward()
        grads.append((f.w1.grad.clone(), f.w2.grad.clone()))
    for a, b in zip(*grads):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------------------------------------- torch_geometric
def _check_sage_conv():
    pyg = _need("torch_geometric.nn")
    g = torch.Generator().manual_seed(7)
    x = torch.randn(30, 12, generator=g)
    ei = torch.randint(0, 30, (2, 90), generator=g)
    ei[1, ei[1] == 5] = 6                                   # node 5 has no in-edge: mean of the empty set = 0
    real = pyg.SAGEConv(12, 9)
    ours = SAGEConvRef(12, 9)
    ours.load_state_dict(real.state_dict())                 # same key names as PyG: lin_l.weight, lin_l.bias, lin_r.weight
    assert sorted(real.state_dict().keys()) == sorted(ours.state_dict().keys())
    want, got = real(x, ei), ours(x, ei)
    assert torch.allclose(got, want, rtol=1e-6, atol=1e-6), float((got - want).abs().max())
    # bipartite form used by HeteroConv
    xs, xd = torch.randn(11, 12, generator=g), torch.randn(7, 12, generator=g)
    eb = torch.stack([torch.randint(0, 11, (20,), generator=g), torch.randint(0, 7, (20,), generator=g)])
    real_b = pyg.SAGEConv((12, 12), 9)
    ours_b = SAGEConvRef((12, 12), 9)
    ours_b.load_state_dict(real_b.state_dict())
    assert torch.allclose(ours_b((xs, xd), eb), real_b((xs, xd), eb), rtol=1e-6, atol=1e-6)


def _check_hetero_conv():
    pyg = _need("torch_geometric.nn")
    g = torch.Generator().manual_seed(8)
    x = {"agv": torch.randn(5, 8, generator=g), "picker": torch.randn(3, 8, generator=g), "location": torch.randn(9, 8, generator=g)}
    ets = [("agv", "a2a", "agv"), ("picker", "p2a", "agv"), ("agv", "a2l", "location"), ("location", "l2p", "picker")]
    n = {"agv": 5, "picker": 3, "location": 9}
    ei = {et: torch.stack([torch.randint(0, n[et[0]], (12,), generator=g), torch.randint(0, n[et[2]], (12,), generator=g)]) for et in ets}
    ei[("location", "l2p", "picker")] = torch.empty((2, 0), dtype=torch.long)      # an empty relation still contributes b_l + W_r x
    real = pyg.HeteroConv({et: pyg.SAGEConv((8, 8), 6) for et in ets}, aggr="mean")
    ours = HeteroConvRef({et: SAGEConvRef((8, 8), 6) for et in ets}, aggr="mean")
    for et in ets:
        ours.convs["__".join(et)].load_state_dict(real.convs[et].state_dict())
    want, got = real(x, ei), ours(x, ei)
    assert sorted(want.keys()) == sorted(got.keys())
    for k in want:
        assert torch.allclose(got[k], want[k], rtol=1e-6, atol=1e-6), k


def _check_batch_collate():
    pygd = _need("torch_geometric.data")
    g = torch.Generator().manual_seed(9)
    reals, refs = [], []
    for n in (4, 1, 7):
        x = torch.randn(n, 5, generator=g)
        e = torch.randint(0, n, (2, 2 * n), generator=g)
        m = torch.rand(n, generator=g) > 0.5
        reals.append(pygd.Data(x=x, edge_index=e, is_current_agent=m))
        refs.append(RefData(x=x, edge_index=e, is_current_agent=m))
    want, got = pygd.Batch.from_data_list(reals), RefBatch.from_data_list(refs)
    for k in ("x", "edge_index", "batch", "ptr", "is_current_agent"):
        assert torch.equal(getattr(got, k), getattr(want, k)), k


_CHECKS = {"fixed_solvers": _check_fixed_solvers, "dopri5": _check_dopri5, "dopri5_gradients": _check_dopri5_gradients,
           "sage_conv": _check_sage_conv, "hetero_conv": _check_hetero_conv, "batch_collate": _check_batch_collate}


@pytest.mark.parametrize("name", sorted(_CHECKS))
def test_oracle_matches_real_library(name):
    _CHECKS[name]()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(_CHECKS))
def test_oracle_matches_real_library_on_the_gpu_box(name):
    _CHECKS[name]()
