"""Secondary variant: node-wise MLP vector field (ODEFunction, scripts/gnode.py:160-174) through the same
native integrators, vs the torch restatement."""
import pytest
import torch
import torch.nn as nn

import swarm_ode_b200 as S
from oracle.torchdiffeq_ref import SolverStats, odeint_ref
from tests._util import FIXED_TOL, rel_l2

pytestmark = pytest.mark.gpu


def _ref_net(H, h, seed=0, gain=1.0):
    torch.manual_seed(seed)
    net = nn.Sequential(nn.Linear(H, h), nn.Tanh(), nn.Linear(h, h), nn.Tanh(), nn.Linear(h, H))
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(gain)
    return net


@pytest.mark.parametrize("M,H,h", [(268, 64, 32), (19, 128, 32), (1000, 64, 32)])
def test_mlp_rhs_and_fixed_solvers(cuda, M, H, h):
    net = _ref_net(H, h)
    f = S.ODEFunction(H, h)
    f.net.load_state_dict(net.state_dict())
    f = f.to(cuda)
    x = torch.randn(M, H)
    with torch.no_grad():
        assert rel_l2(f(torch.tensor(0.0), x.to(cuda)), net(x)) <= 1e-5
        t = torch.tensor([0.0, 0.5, 1.0])
        for method in ("euler", "midpoint", "rk4"):
            want = odeint_ref(lambda tt, y: net(y), x, t, method=method)
            got = S.odeint(f, x.to(cuda), t.to(cuda), method=method)
            assert rel_l2(got, want) <= FIXED_TOL, method


@pytest.mark.parametrize("gain,rtol,atol", [(1.0, 1e-5, 1e-7), (3.0, 1e-4, 1e-6), (6.0, 1e-3, 1e-5)])
def test_mlp_dopri5_decisions(cuda, gain, rtol, atol):
    H, h, M = 64, 32, 268
    net = _ref_net(H, h, seed=1, gain=gain)
    f = S.ODEFunction(H, h)
    f.net.load_state_dict(net.state_dict())
    f = f.to(cuda)
    x = torch.randn(M, H, generator=torch.Generator().manual_seed(2))
    t = torch.tensor([0.0, 1.0, 2.5])
    rst = SolverStats()
    with torch.no_grad():
        want = odeint_ref(lambda tt, y: net(y), x, t, rtol=rtol, atol=atol, method="dopri5", stats=rst)
        got, st = S.ops.mlp_integrate(x.to(cuda), f.param_list(), t, "dopri5", rtol=rtol, atol=atol)
    print(f"mlp dopri5 gain {gain}: accepted {st.n_accepted}/{rst.n_accepted} attempted {st.n_attempted}/{rst.n_attempted} "
          f"min|ratio-1| {st.min_margin:.3g}")
    assert st.accepted == rst.accepted and st.nfe == rst.nfe
    assert rel_l2(got, want) <= FIXED_TOL


def _grad_check(model_params, ref_params, tol=FIXED_TOL):
    for (n, p), q in zip(model_params, ref_params):
        assert p.grad is not None, n
        assert rel_l2(p.grad, q.grad) <= tol, (n, rel_l2(p.grad, q.grad))


@pytest.mark.parametrize("M,H,h", [(268, 64, 32), (19, 128, 32), (9, 64, 32)])
def test_mlp_field_backward(cuda, M, H, h):
    """ODEFunction.forward under autograd (scripts/gnode.py:173-174): input and parameter gradients."""
    net = _ref_net(H, h, seed=3)
    f = S.ODEFunction(H, h)
    f.net.load_state_dict(net.state_dict())
    f = f.to(cuda)
    x = torch.randn(M, H, generator=torch.Generator().manual_seed(4))
    xr = x.clone().requires_grad_(True)
    net(xr).pow(2).sum().backward()
    xg = x.to(cuda).requires_grad_(True)
    f(torch.tensor(0.0), xg).pow(2).sum().backward()
    assert rel_l2(xg.grad, xr.grad) <= 1e-5
    _grad_check(list(f.net.named_parameters()), list(net.parameters()), 1e-5)


@pytest.mark.parametrize("method", ["euler", "midpoint", "rk4"])
def test_mlp_fixed_solver_backward(cuda, method):
    """loss.backward() through odeint(ODEFunction, ...) on a fixed grid, loss on every output time."""
    H, h, M = 64, 32, 268
    net = _ref_net(H, h, seed=5)
    f = S.ODEFunction(H, h)
    f.net.load_state_dict(net.state_dict())
    f = f.to(cuda)
    x = torch.randn(M, H, generator=torch.Generator().manual_seed(6))
    t = torch.tensor([0.0, 0.3, 1.0])
    w_out = torch.tensor([0.5, 1.0, 2.0]).view(-1, 1, 1)
    xr = x.clone().requires_grad_(True)
    (odeint_ref(lambda tt, y: net(y), xr, t, method=method) * w_out).pow(2).mean().backward()
    xg = x.to(cuda).requires_grad_(True)
    (S.odeint(f, xg, t.to(cuda), method=method) * w_out.to(cuda)).pow(2).mean().backward()
    assert rel_l2(xg.grad, xr.grad) <= FIXED_TOL
    _grad_check(list(f.net.named_parameters()), list(net.parameters()))


@pytest.mark.parametrize("gain,t_points", [(1.0, (0.0, 1.0)), (3.0, (0.0, 0.4, 1.0, 2.5)), (6.0, (0.0, 1.0))])
def test_mlp_dopri5_backward(cuda, gain, t_points):
    """loss.backward() through the adaptive solve of the MLP field (the reference's default odeint method,
    scripts/gnode.py:136-137), against autograd through the oracle replaying the same attempted steps."""
    H, h, M = 64, 32, 268
    net = _ref_net(H, h, seed=7, gain=gain)
    f = S.ODEFunction(H, h)
    f.net.load_state_dict(net.state_dict())
    f = f.to(cuda)
    x = torch.randn(M, H, generator=torch.Generator().manual_seed(8))
    t = torch.tensor(t_points)
    w_out = torch.linspace(1.0, 2.0, len(t_points)).view(-1, 1, 1)
    xg = x.to(cuda).requires_grad_(True)
    sol, st = S.ops.mlp_integrate(xg, f.param_list(), t, "dopri5", rtol=1e-4, atol=1e-6)
    (sol * w_out.to(cuda)).pow(2).mean().backward()
    xr = x.clone().requires_grad_(True)
    rst = SolverStats()
    want = odeint_ref(lambda tt, y: net(y), xr, t, rtol=1e-4, atol=1e-6, method="dopri5", stats=rst,
                      options={"imposed_dts": list(st.dts)})
    (want * w_out).pow(2).mean().backward()
    # float64 oracle on the same steps: with gain 6 the flow amplifies fp32 rounding over its 16 steps, so two fp32
    # evaluations of the gradient differ by more than 1e-4; the bar is then the fp32 oracle's own distance to float64
    net64 = _ref_net(H, h, seed=7, gain=gain).double()
    x64 = x.double().requires_grad_(True)
    want64 = odeint_ref(lambda tt, y: net64(y), x64, t.double(), rtol=1e-4, atol=1e-6, method="dopri5",
                        options={"imposed_dts": list(st.dts)})
    (want64 * w_out.double()).pow(2).mean().backward()
    e_ref = max([rel_l2(xr.grad, x64.grad)] + [rel_l2(p.grad, q.grad) for p, q in zip(net.parameters(), net64.parameters())])
    tol = max(FIXED_TOL, 3.0 * e_ref)
    print(f"mlp dopri5 bwd gain {gain}: accepted {st.n_accepted}/{st.n_attempted}; fp32 oracle vs float64 {e_ref:.2e}, "
          f"ours vs float64 {rel_l2(xg.grad, x64.grad):.2e}")
    assert rst.accepted == st.accepted
    assert rel_l2(sol, want) <= FIXED_TOL
    assert rel_l2(xg.grad, x64.grad) <= tol
    for (n, p), q in zip(f.net.named_parameters(), net64.parameters()):
        assert rel_l2(p.grad, q.grad) <= tol, (n, rel_l2(p.grad, q.grad), tol)
