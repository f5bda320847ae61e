"""Pins the torchdiffeq restatement (oracle/torchdiffeq_ref.py): tableau vs SciPy's RK45 constants,
fixed-step closed forms, dopri5 values vs analytic / SciPy, controller behaviour."""
import math

import numpy as np
import pytest
import torch

from oracle import torchdiffeq_ref as R
from oracle.torchdiffeq_ref import SolverStats, odeint_ref


def test_dopri5_tableau_matches_scipy_rk45():
    from scipy.integrate._ivp.rk import RK45
    # SciPy: C (nodes), A (stage matrix), B (5th order weights), E (error weights, incl. FSAL stage)
    assert np.allclose(RK45.C[1:], R.DP_ALPHA[:5])
    for i, row in enumerate(R.DP_BETA[:5]):
        assert np.allclose(RK45.A[i + 1, :len(row)], row, rtol=0, atol=1e-15)
    assert np.allclose(RK45.B, R.DP_BETA[5], rtol=0, atol=1e-15)
    assert np.allclose(RK45.B, R.DP_C_SOL[:6], rtol=0, atol=1e-15)
    # SciPy's E = b5 - b4 with Dormand & Prince's original embedded 4th-order weights b4
    # (5179/57600, ...).  torchdiffeq embeds the variant (1951/21600, 22642/50085, 451/720, -12231/42400,
    # 649/6300, 1/60), whose error weights are exactly -2/3 of SciPy's: the same error direction with a
    # different constant.  This pins every entry of c_error against an independent source.
    e = np.array(R.DP_C_ERR)
    assert np.allclose(e, -2.0 / 3.0 * RK45.E, rtol=0, atol=1e-15)
    assert abs(sum(R.DP_C_ERR)) < 1e-15           # both formulas are consistent
    assert abs(sum(R.DP_C_SOL) - 1) < 1e-15
    for a, row in zip(R.DP_ALPHA, R.DP_BETA):     # row-sum condition
        assert abs(sum(row) - a) < 1e-14


def _lin_field(A):
    return lambda t, y: y @ A.t()


@pytest.mark.parametrize("method,order", [("euler", 1), ("midpoint", 2), ("rk4", 4)])
def test_fixed_step_linear_closed_form(method, order):
    torch.manual_seed(0)
    A = torch.randn(5, 5, dtype=torch.float64) * 0.3
    y0 = torch.randn(7, 5, dtype=torch.float64)
    h = 0.37
    t = torch.tensor([0.0, h], dtype=torch.float64)
    y1 = odeint_ref(_lin_field(A), y0, t, method=method)[1]
    # any p-stage p-th order explicit RK applied to y' = Ay gives the degree-p Taylor polynomial of exp(hA)
    P = torch.eye(5, dtype=torch.float64)
    term = torch.eye(5, dtype=torch.float64)
    for k in range(1, order + 1):
        term = term @ (h * A) / k
        P = P + term
    assert torch.allclose(y1, y0 @ P.t(), rtol=1e-12, atol=1e-13)


def test_fixed_grid_outputs_every_time_point_and_nfe():
    st = SolverStats()
    t = torch.arange(0, 6, dtype=torch.float32)
    sol = odeint_ref(lambda t, y: -0.1 * y, torch.ones(3), t, method="rk4", stats=st)
    assert sol.shape == (6, 3) and st.nfe == 20
    assert torch.allclose(sol[:, 0], torch.exp(-0.1 * t), atol=1e-6)
    assert torch.equal(sol[0], torch.ones(3))


def test_rk4_is_three_eighths_rule_not_classic():
    # y' = t^... autonomous trick: use a nonlinear field where classic RK4 and 3/8 rule differ
    f = lambda t, y: torch.sin(y) * y
    y0 = torch.tensor([0.7], dtype=torch.float64)
    h = 0.5
    ours = odeint_ref(f, y0, torch.tensor([0.0, h], dtype=torch.float64), method="rk4")[1]
    k1 = f(0, y0); k2 = f(0, y0 + h * k1 / 3); k3 = f(0, y0 + h * (k2 - k1 / 3)); k4 = f(0, y0 + h * (k1 - k2 + k3))
    three_eighths = y0 + h * (k1 + 3 * (k2 + k3) + k4) / 8
    c1 = f(0, y0); c2 = f(0, y0 + h * c1 / 2); c3 = f(0, y0 + h * c2 / 2); c4 = f(0, y0 + h * c3)
    classic = y0 + h * (c1 + 2 * c2 + 2 * c3 + c4) / 6
    assert torch.allclose(ours, three_eighths, rtol=1e-14)
    assert not torch.allclose(ours, classic, rtol=1e-9)


def test_dopri5_exponential_decay_and_stats():
    st = SolverStats()
    t = torch.tensor([0.0, 0.5, 1.0, 2.5])
    y0 = torch.tensor([[1.0, 2.0], [3.0, -4.0]])
    sol = odeint_ref(lambda t, y: -y, y0, t, rtol=1e-6, atol=1e-8, method="dopri5", stats=st)
    want = y0.unsqueeze(0) * torch.exp(-t).view(-1, 1, 1)
    assert torch.allclose(sol, want, rtol=2e-5, atol=1e-6)
    assert st.n_attempted >= st.n_accepted > 0
    assert st.nfe == 2 + 6 * st.n_attempted          # f0, initial-step probe, 6 new evaluations per attempt (FSAL)
    assert all(r <= 1 for r, a in zip(st.error_ratios, st.accepted) if a)
    assert all(r > 1 for r, a in zip(st.error_ratios, st.accepted) if not a)


def test_dopri5_matches_scipy_values_on_nonlinear_system():
    from scipy.integrate import solve_ivp

    def f_np(t, y):
        return np.array([y[1], -np.sin(y[0]) - 0.1 * y[1]])

    def f_t(t, y):
        return torch.stack([y[1], -torch.sin(y[0]) - 0.1 * y[1]])

    y0 = [1.0, 0.0]
    ts = np.linspace(0, 5, 6)
    ref = solve_ivp(f_np, (0, 5), y0, method="RK45", t_eval=ts, rtol=1e-10, atol=1e-12).y.T
    sol = odeint_ref(f_t, torch.tensor(y0, dtype=torch.float64), torch.tensor(ts), rtol=1e-8, atol=1e-10,
                     method="dopri5")
    assert np.allclose(sol.numpy(), ref, rtol=1e-6, atol=1e-7)


def test_dopri5_default_method_and_rejections_shrink_dt():
    st = SolverStats()
    # stiff-ish start forces at least one rejection with loose initial step
    f = lambda t, y: -50.0 * (y - torch.cos(t.to(y.dtype)))
    sol = odeint_ref(f, torch.zeros(1, dtype=torch.float64), torch.tensor([0.0, 1.0], dtype=torch.float64),
                     rtol=1e-5, atol=1e-7, stats=st)
    assert st.method == "dopri5"
    for i, acc in enumerate(st.accepted[:-1]):
        ratio = st.dts[i + 1] / st.dts[i]
        if acc:
            assert 1.0 - 1e-12 <= ratio <= 10.0 + 1e-9       # never shrinks after an accepted step
        else:
            assert 0.2 - 1e-12 <= ratio < 1.0
    assert math.isfinite(float(sol[-1]))


def test_dopri5_steps_past_output_times_and_interpolates():
    # a very smooth problem: one or two huge steps cover all 11 output times
    st = SolverStats()
    t = torch.linspace(0, 1, 11, dtype=torch.float64)
    sol = odeint_ref(lambda t, y: 0.01 * y, torch.ones(2, dtype=torch.float64), t, rtol=1e-3, atol=1e-4,
                     method="dopri5", stats=st)
    assert st.n_accepted < 10
    assert torch.allclose(sol[:, 0], torch.exp(0.01 * t), rtol=1e-6)


def test_backprop_through_fixed_solver_matches_closed_form():
    A = (torch.randn(3, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(1)) * 0.2).requires_grad_()
    y0 = torch.ones(2, 3, dtype=torch.float64)
    sol = odeint_ref(lambda t, y: y @ A.t(), y0, torch.tensor([0.0, 1.0], dtype=torch.float64), method="euler")
    sol[1].sum().backward()
    # y1 = y0 (I + A)^T  => d sum / dA[i, j] = sum_rows y0[:, j]
    assert torch.allclose(A.grad, y0.sum(0).expand(3, 3))


def test_invalid_method_raises():
    with pytest.raises(ValueError):
        odeint_ref(lambda t, y: y, torch.ones(1), torch.tensor([0.0, 1.0]), method="rk45")


def test_oracle_adjoint_matches_backprop_at_small_steps():
    """odeint_adjoint_ref (restated torchdiffeq adjoint) against autograd through odeint_ref: both are gradients of the
    same loss and differ only by the O(dt^4) error of the backward rk4 solve."""
    from oracle.torchdiffeq_ref import odeint_adjoint_ref
    torch.manual_seed(0)
    W = (torch.randn(5, 5, dtype=torch.float64) * 0.3).requires_grad_(True)
    f = lambda t, x: torch.tanh(x @ W.T)
    y0 = torch.randn(7, 5, dtype=torch.float64)
    t = torch.linspace(0, 1, 41, dtype=torch.float64)
    g = torch.randn(41, 7, 5, dtype=torch.float64)
    sol, gy0, gp = odeint_adjoint_ref(f, [W], y0, t, g, "rk4")
    y0r = y0.clone().requires_grad_(True)
    s2 = odeint_ref(f, y0r, t, method="rk4")
    (s2 * g).sum().backward()
    assert torch.equal(sol, s2.detach())
    assert float((gy0 - y0r.grad).norm() / y0r.grad.norm()) < 1e-8
    assert float((gp[0] - W.grad).norm() / W.grad.norm()) < 1e-8


def _dense_weights(x: float):
    """w_s(x) of csrc/fold.cu:dopri5_dense_weights, restated: the quartic of _interp_fit / _interp_evaluate expanded in the
    seven stage derivatives (d = Kronecker delta, c = c_sol, m = c_mid)."""
    from oracle.torchdiffeq_ref import DP_C_MID, DP_C_SOL
    w = []
    for s in range(7):
        d0, d6, cs, ms = float(s == 0), float(s == 6), DP_C_SOL[s], DP_C_MID[s]
        w.append(x * d0 + x ** 2 * (d6 - 4 * d0 - 5 * cs + 16 * ms) + x ** 3 * (5 * d0 - 3 * d6 + 14 * cs - 32 * ms)
                 + x ** 4 * (2 * d6 - 2 * d0 - 8 * cs + 16 * ms))
    return w


@pytest.mark.parametrize("x", [0.0, 0.137, 0.5, 0.9, 1.0])
def test_dense_output_is_a_stage_combination(x):
    """The identity behind the one-projection dense output of the folded dopri5 (csrc/integrate.cu, gnode_set_dopri5_fsal):
    torchdiffeq's quartic through (y0, y1, y_mid, f0, f1) equals y0 + dt sum_s w_s(x) k_s -- the y0 terms of its coefficients
    cancel -- and w(1) = c_sol.  Checked against the oracle's own _interp_fit / _interp_evaluate in float64."""
    from oracle.torchdiffeq_ref import DP_C_MID, DP_C_SOL, _interp_evaluate, _interp_fit_dopri5
    g = torch.Generator().manual_seed(3)
    y0 = torch.randn(5, 7, generator=g, dtype=torch.float64)
    k = torch.randn(5, 7, 7, generator=g, dtype=torch.float64)            # [..., stage]
    dt = torch.tensor(0.37, dtype=torch.float64)
    y1 = y0 + k.matmul(dt * torch.tensor(DP_C_SOL, dtype=torch.float64))
    coeffs = _interp_fit_dopri5(y0, y1, k, dt, torch.tensor(DP_C_MID, dtype=torch.float64))
    t0, t1 = torch.tensor(2.0, dtype=torch.float64), torch.tensor(2.0, dtype=torch.float64) + dt
    want = _interp_evaluate(coeffs, t0, t1, t0 + x * dt)
    w = torch.tensor(_dense_weights(x), dtype=torch.float64)
    got = y0 + k.matmul(dt * w)
    assert torch.allclose(got, want, rtol=0, atol=1e-12 * float(want.abs().max() + 1))
    if x == 1.0:
        assert torch.allclose(w, torch.tensor(DP_C_SOL, dtype=torch.float64), rtol=0, atol=1e-13)
        assert torch.allclose(got, y1, rtol=0, atol=1e-12)


def test_last_dopri5_stage_is_evaluated_at_the_solution():
    """FSAL: row 6 of the Dormand-Prince tableau equals c_sol, so stage 6 of an attempt sees y1 itself -- what lets the
    folded solver hand Z of stage 6 to the next attempt as Z_0 = y1 @ w1cat^T (csrc/chain_fwd.cu: Args::Zlast)."""
    from oracle.torchdiffeq_ref import DP_BETA, DP_C_SOL
    assert list(DP_BETA[-1]) + [0] == list(DP_C_SOL)
    assert abs(sum(DP_C_SOL) - 1.0) < 1e-15
