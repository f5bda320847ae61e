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


def test_mlp_field_refuses_autograd(cuda):
    f = S.ODEFunction(16, 8).to(cuda)
    with pytest.raises(S.GnodeError, match="forward-only"):
        f(torch.tensor(0.0), torch.randn(4, 16, device=cuda))
