"""Adjoint backward of the fixed-grid solvers (gnode_integrate_fixed_adjoint) against the oracle's restatement of
torchdiffeq.odeint_adjoint, and against backprop-through-the-solver."""
import pytest
import torch

import swarm_ode_b200 as S
from oracle.torchdiffeq_ref import odeint_adjoint_ref
from oracle.train_gde_ref import GraphODERef
from tests._util import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def _models(solver, dev, graphs=6, seed=5, adjoint=True):
    batch, nxt = S.synthetic.warehouse_batch(graphs, num_agvs=4, num_pickers=3, seed=seed)
    D = batch.x.shape[1]
    ours = S.GraphODE(D, 4, 3, hidden_dim=64, ode_solver=solver, adjoint=adjoint)
    S.synthetic.init_weights(ours, seed=2, conv3_scale=0.05)
    ref = GraphODERef(D, 4, 3, hidden_dim=64, ode_solver=solver)
    ref.load_state_dict(ours.state_dict())
    return ours.to(dev), ref.double(), batch


@pytest.mark.parametrize("solver", ["euler", "midpoint", "rk4"])
def test_adjoint_gradients_match_oracle_adjoint(cuda, solver):
    """Same discretisation of the same augmented system: the native adjoint must agree with the float64 oracle adjoint
    to fp32 accuracy (1e-4 on every parameter gradient and on dL/dy0), for a cotangent on EVERY output time."""
    ours, ref, batch = _models(solver, cuda)
    t = torch.tensor([0.0, 0.4, 1.0])
    g = torch.Generator().manual_seed(0)
    N, D = batch.x.shape
    gsol = torch.randn(3, N, D, generator=g)
    # oracle (float64)
    params = [p for p in ref.ode_func.parameters()]
    ei = batch.edge_index
    f = lambda tt, x: ref.ode_func(tt, x, ei)
    sol_r, gy0_r, gp_r = odeint_adjoint_ref(f, params, batch.x.double(), t.double(), gsol.double(), solver)
    # ours
    b = batch.to(cuda)
    x = b.x.clone().requires_grad_(True)
    graph = S.csr_for(b.edge_index, N)
    sol = S.odeint_adjoint(ours.ode_func.bind(graph), x, t.to(cuda), method=solver)
    sol.backward(gsol.to(cuda))
    assert rel_l2(sol.detach().cpu(), sol_r.float()) <= 1e-5
    assert rel_l2(x.grad.cpu(), gy0_r.float()) <= 1e-4
    names = [n for n, _ in ref.ode_func.named_parameters()]
    ours_p = dict(ours.ode_func.named_parameters())
    worst = 0.0
    for n, gr in zip(names, gp_r):
        e = rel_l2(ours_p[n].grad.cpu(), gr.float())
        worst = max(worst, e)
        assert e <= 1e-4, (n, e)
    print(f"adjoint[{solver}] worst parameter-gradient rel-L2 vs oracle adjoint: {worst:.2e}")


def test_adjoint_close_to_backprop_through_solver(cuda):
    """Both are gradients of the same loss; they differ by the discretisation error of the backward solve (small for
    several short rk4 steps), not by more."""
    ours_adj, _, batch = _models("rk4", cuda, adjoint=True)
    ours_bp, _, _ = _models("rk4", cuda, adjoint=False)
    t = torch.linspace(0.0, 1.0, 6, device=cuda)
    b = batch.to(cuda)
    grads = []
    for m in (ours_adj, ours_bp):
        m.zero_grad(set_to_none=True)
        out = m(b, t)
        (out["trajectories"][-1] ** 2).mean().backward()
        grads.append({n: p.grad.clone() for n, p in m.named_parameters()})
    for n in grads[0]:
        assert rel_l2(grads[0][n], grads[1][n]) <= 2e-2, n


def test_adjoint_rejects_dopri5(cuda):
    ours, _, batch = _models("dopri5", cuda)
    b = batch.to(cuda)
    with pytest.raises(S.GnodeError):
        ours(b, torch.tensor([0.0, 1.0], device=cuda))
