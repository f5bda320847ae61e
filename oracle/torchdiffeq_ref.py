"""Oracle restatement of ``torchdiffeq.odeint`` for the solvers the reference can select.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``) -- parity unpinned: torchdiffeq
is an un-vendored, un-pinned dependency of the reference (imported at
``scripts/train_gde.py:11``, ``scripts/gnode.py:10``, ``scripts/run_gnode.py:12``)
and cannot be installed here.  This file restates the published algorithm of
torchdiffeq 0.2.x, anchored on the reference call sites

  * ``odeint(f, x0, time_span, method=self.ode_solver, rtol=1e-3, atol=1e-4)``
                                                     scripts/train_gde.py:78-85
  * ``odeint(self.ode_func, all_embeddings, t)``     scripts/gnode.py:136-137
                                        (default method dopri5, rtol 1e-7, atol 1e-9)
  * ``odeint(..., method='euler')``                  scripts/run_gnode.py:134-135

Semantics restated (all ``[upstream]``):
  fixed grid (euler / midpoint / rk4): the grid is ``t`` itself; ``rk4`` is the
  3/8-rule variant; rtol/atol are ignored; the state keeps its dtype and time
  keeps the dtype of ``t``.
  dopri5: Dormand-Prince 5(4) with FSAL, Hairer's initial-step heuristic, one
  global RMS error norm over the whole state tensor, controller
  (safety .9, ifactor 10, dfactor .2, no shrink after an accepted step), no
  clipping of dt to output times, quartic dense output; time in float64, state
  in its own dtype.
Everything is plain differentiable PyTorch, so ``loss.backward()`` through the
returned solution is the reference's "backprop through the solver".
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional

import torch

# ----------------------------------------------------------------------------
# Dormand-Prince tableau (float64 masters; cast to the state dtype at use).
# ----------------------------------------------------------------------------
DP_ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
DP_BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
DP_C_SOL = [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0]
DP_C_ERR = [
    35 / 384 - 1951 / 21600,
    0,
    500 / 1113 - 22642 / 50085,
    125 / 192 - 451 / 720,
    -2187 / 6784 - -12231 / 42400,
    11 / 84 - 649 / 6300,
    -1.0 / 60.0,
]
DP_C_MID = [
    6025192743 / 30085553152 / 2,
    0,
    51252292925 / 65400821598 / 2,
    -2691868925 / 45128329728 / 2,
    187940372067 / 1594534317056 / 2,
    -1776094331 / 19743644256 / 2,
    11237099 / 235043384 / 2,
]

_ONE_THIRD = 1 / 3
_TWO_THIRDS = 2 / 3


@dataclass
class SolverStats:
    """Bookkeeping torchdiffeq does not expose but the parity gate needs."""
    method: str = ""
    nfe: int = 0
    n_accepted: int = 0
    n_attempted: int = 0
    error_ratios: List[float] = field(default_factory=list)   # one per attempted step
    dts: List[float] = field(default_factory=list)            # dt of each attempted step
    accepted: List[bool] = field(default_factory=list)
    first_step: float = float("nan")


# ----------------------------------------------------------------------------
# Fixed-grid solvers
# ----------------------------------------------------------------------------
def _euler_step(func, t0, dt, t1, y0, st):
    f0 = func(t0, y0); st.nfe += 1
    return dt * f0


def _midpoint_step(func, t0, dt, t1, y0, st):
    half_dt = 0.5 * dt
    f0 = func(t0, y0); st.nfe += 1
    y_mid = y0 + f0 * half_dt
    st.nfe += 1
    return dt * func(t0 + half_dt, y_mid)


def _rk4_38_step(func, t0, dt, t1, y0, st):
    """torchdiffeq ``rk4_alt_step_func`` (3/8 rule), operation order kept."""
    k1 = func(t0, y0)
    k2 = func(t0 + dt * _ONE_THIRD, y0 + dt * k1 * _ONE_THIRD)
    k3 = func(t0 + dt * _TWO_THIRDS, y0 + dt * (k2 - k1 * _ONE_THIRD))
    k4 = func(t1, y0 + dt * (k1 - k2 + k3))
    st.nfe += 4
    return (k1 + 3 * (k2 + k3) + k4) * dt * 0.125


_FIXED = {"euler": _euler_step, "midpoint": _midpoint_step, "rk4": _rk4_38_step}


def _integrate_fixed(func, y0, t, method, st):
    step = _FIXED[method]
    sol = [y0]
    y = y0
    # grid == t (no step_size option at any reference call site) so every grid point is an
    # output point and no interpolation happens.
    for j in range(1, len(t)):
        t0, t1 = t[j - 1], t[j]
        dt = t1 - t0
        y = y + step(func, t0, dt, t1, y, st)
        sol.append(y)
    return torch.stack(sol, dim=0)


# ----------------------------------------------------------------------------
# Adaptive dopri5
# ----------------------------------------------------------------------------
def _rms_norm(x: torch.Tensor) -> torch.Tensor:
    return x.abs().pow(2).mean().sqrt()


def _select_initial_step(func, t0, y0, order, rtol, atol, f0, st):
    dtype = y0.dtype
    scale = atol + torch.abs(y0) * rtol
    d0 = _rms_norm(y0 / scale).abs()
    d1 = _rms_norm(f0 / scale).abs()
    if d0 < 1e-5 or d1 < 1e-5:
        h0 = torch.tensor(1e-6, dtype=dtype)
    else:
        h0 = 0.01 * d0 / d1
    h0 = h0.abs()
    y1 = y0 + h0 * f0
    f1 = func(t0 + h0, y1); st.nfe += 1
    d2 = torch.abs(_rms_norm((f1 - f0) / scale) / h0)
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = torch.max(torch.tensor(1e-6, dtype=dtype), h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1.0 / float(order + 1))
    h1 = h1.abs()
    return torch.min(100 * h0, h1).to(t0.dtype)


def _rk_step_dopri5(func, y0, f0, t0, dt, t1, tab, st):
    """``_runge_kutta_step``: stage inputs are ``y0 + sum_j k_j * (beta_ij * dt)`` summed over a
    trailing stage axis in the state dtype."""
    t0s, dts, t1s = t0.to(y0.dtype), dt.to(y0.dtype), t1.to(y0.dtype)
    ks = [f0]
    yi = y0
    for i, (alpha_i, beta_i) in enumerate(zip(tab["alpha"], tab["beta"])):
        ti = t1s if float(alpha_i) == 1.0 else t0s + alpha_i * dts
        kk = torch.stack(ks, dim=-1)
        yi = y0 + torch.sum(kk * (beta_i * dts), dim=-1).view_as(f0)
        f = func(ti, yi); st.nfe += 1
        ks.append(f)
    k = torch.stack(ks, dim=-1)
    # Dormand-Prince: c_sol == last beta row (+0), so y1 is the last stage input (FSAL).
    y1 = yi
    f1 = ks[-1]
    y1_error = torch.sum(k * (dts * tab["c_error"]), dim=-1)
    return y1, f1, y1_error, k


def _interp_fit_dopri5(y0, y1, k, dt, mid):
    dt = dt.type_as(y0)
    y_mid = y0 + k.matmul(dt * mid).view_as(y0)
    f0 = k[..., 0]
    f1 = k[..., -1]
    a = 2 * dt * (f1 - f0) - 8 * (y1 + y0) + 16 * y_mid
    b = dt * (5 * f0 - 3 * f1) + 18 * y0 + 14 * y1 - 32 * y_mid
    c = dt * (f1 - 4 * f0) - 11 * y0 - 5 * y1 + 16 * y_mid
    d = dt * f0
    e = y0
    return [e, d, c, b, a]


def _interp_evaluate(coeffs, t0, t1, t):
    assert bool((t0 <= t) & (t <= t1)), f"invalid interpolation, fails `t0 <= t <= t1`: {t0}, {t}, {t1}"
    x = ((t - t0) / (t1 - t0)).to(coeffs[0].dtype)
    total = coeffs[0] + x * coeffs[1]
    x_power = x
    for c in coeffs[2:]:
        x_power = x_power * x
        total = total + x_power * c
    return total


def _optimal_step_size(last_step, error_ratio, safety, ifactor, dfactor, order):
    if error_ratio == 0:
        return last_step * ifactor
    if error_ratio < 1:
        dfactor = torch.ones((), dtype=last_step.dtype)
    error_ratio = error_ratio.type_as(last_step)
    exponent = torch.tensor(order, dtype=last_step.dtype).reciprocal()
    factor = torch.min(ifactor, torch.max(safety / error_ratio ** exponent, dfactor))
    return last_step * factor


def _integrate_dopri5(func, y0, t, rtol, atol, st, first_step=None, max_num_steps=2 ** 31 - 1, imposed_dts=None):
    tdtype = torch.promote_types(torch.float64, y0.dtype)
    rtol_t = torch.as_tensor(rtol, dtype=tdtype)
    atol_t = torch.as_tensor(atol, dtype=tdtype)
    safety = torch.as_tensor(0.9, dtype=tdtype)
    ifactor = torch.as_tensor(10.0, dtype=tdtype)
    dfactor = torch.as_tensor(0.2, dtype=tdtype)
    order = 5
    tab = {
        "alpha": torch.tensor(DP_ALPHA, dtype=torch.float64).to(y0.dtype),
        "beta": [torch.tensor(b, dtype=torch.float64).to(y0.dtype) for b in DP_BETA],
        "c_error": torch.tensor(DP_C_ERR, dtype=torch.float64).to(y0.dtype),
    }
    mid = torch.tensor(DP_C_MID, dtype=torch.float64).to(y0.dtype)

    t = t.to(tdtype)
    f0 = func(t[0], y0); st.nfe += 1
    # Step sizes are constants of the differentiation: torchdiffeq computes the next step under @torch.no_grad()
    # (misc._optimal_step_size); the initial step is detached here as well (its derivative would only move the step
    # boundaries: an O(local error) effect), so that autograd through this restatement is exactly "backprop through the
    # accepted steps", which is what libgnode_b200's gnode_integrate_dopri5_bwd computes.
    if first_step is None:
        dt = _select_initial_step(func, t[0], y0, order - 1, rtol_t, atol_t, f0, st).detach()
    else:
        dt = torch.as_tensor(first_step, dtype=tdtype)
    st.first_step = float(dt)

    # rk_state = (y1, f1, t0, t1, dt, interp_coeff)
    y1, f1, s_t0, s_t1, interp = y0, f0, t[0], t[0], [y0] * 5
    sol = [y0]
    for i in range(1, len(t)):
        next_t = t[i]
        n_steps = 0
        while next_t > s_t1:
            assert n_steps < max_num_steps, f"max_num_steps exceeded ({n_steps}>={max_num_steps})"
            # ---- _adaptive_step ----
            ya, fa, ta = y1, f1, s_t1
            tb = ta + dt
            assert ta + dt > ta, f"underflow in dt {float(dt)}"
            assert torch.isfinite(ya).all(), "non-finite values in state `y`"
            yb, fb, y_err, k = _rk_step_dopri5(func, ya, fa, ta, dt, tb, tab, st)
            error_tol = atol_t + rtol_t * torch.max(ya.abs(), yb.abs())
            error_ratio = _rms_norm(y_err / error_tol).abs()
            accept = bool(error_ratio <= 1)
            st.n_attempted += 1
            st.error_ratios.append(float(error_ratio))
            st.dts.append(float(dt))
            st.accepted.append(accept)
            if accept:
                st.n_accepted += 1
                interp = _interp_fit_dopri5(ya, yb, k, dt, mid)
                y1, f1, s_t0, s_t1 = yb, fb, ta, tb
            else:
                s_t0, s_t1 = ta, ta
            dt = _optimal_step_size(dt, error_ratio.detach(), safety, ifactor, dfactor, order).detach()
            n_steps += 1
            if imposed_dts is not None:
                # test hook: replay a given sequence of attempted steps (the controller above is then only bookkeeping),
                # to compare two implementations on the SAME discretisation when a decision is rounding-level close
                assert len(imposed_dts) > st.n_attempted or not next_t > s_t1, "imposed step sequence too short"
                if len(imposed_dts) > st.n_attempted:
                    dt = torch.as_tensor(imposed_dts[st.n_attempted], dtype=tdtype)
        sol.append(_interp_evaluate(interp, s_t0, s_t1, next_t))
    return torch.stack(sol, dim=0)


# ----------------------------------------------------------------------------
# Public entry
# ----------------------------------------------------------------------------
def odeint_ref(func: Callable, y0: torch.Tensor, t: torch.Tensor, *, rtol: float = 1e-7,
               atol: float = 1e-9, method: Optional[str] = None, options: Optional[dict] = None,
               stats: Optional[SolverStats] = None) -> torch.Tensor:
    """``odeint(func, y0, t, rtol=, atol=, method=)`` -> ``[len(t), *y0.shape]``.

    ``method=None`` selects dopri5 like upstream.  ``stats`` (optional) is filled in place.
    """
    st = stats if stats is not None else SolverStats()
    method = method or "dopri5"
    st.method = method
    assert t.dim() == 1 and len(t) >= 1
    if len(t) > 1:
        assert bool((t[1:] > t[:-1]).all()), "t must be strictly increasing"
    if method in _FIXED:
        return _integrate_fixed(func, y0, t, method, st)
    if method == "dopri5":
        opts = options or {}
        dts = opts.get("imposed_dts")
        return _integrate_dopri5(func, y0, t, rtol, atol, st, first_step=opts.get("first_step") if dts is None else dts[0],
                                 imposed_dts=dts)
    raise ValueError(f"Invalid method \"{method}\"")


# ----------------------------------------------------------------------------
# Adjoint method (torchdiffeq.odeint_adjoint, fixed-grid solvers)
# ----------------------------------------------------------------------------
def odeint_adjoint_ref(func: Callable, params: List[torch.Tensor], y0: torch.Tensor, t: torch.Tensor, grad_sol: torch.Tensor,
                       method: str):
    """Restatement of ``OdeintAdjointMethod`` [upstream torchdiffeq 0.2.x, ``_impl/adjoint.py``] for the fixed-grid
    solvers: forward solve under ``no_grad``; backward = for i = len(t)-1 .. 1 integrate the augmented state
    ``(y, adj_y, adj_params)`` with dynamics ``(f, -adj_y^T df/dy, -adj_y^T df/dparams)`` over ``t[i-1:i+1].flip(0)``
    with the same method (one step: the grid is ``t``), then ``y <- y[i-1]`` (the stored solution) and
    ``adj_y += grad_y[i-1]``.  Returns ``(solution, dL/dy0, [dL/dparam])`` for the cotangent ``grad_sol`` of the
    solution.  The time cotangent (``vjp_t``) is not formed: ``t`` does not require grad at the reference's call sites."""
    assert method in _FIXED
    with torch.no_grad():
        sol = _integrate_fixed(func, y0, t, method, SolverStats())
    adj_y = grad_sol[-1].clone()
    adj_p = [torch.zeros_like(p) for p in params]
    sizes = [p.numel() for p in params]

    def pack(y, a, gp):
        return torch.cat([y.reshape(-1), a.reshape(-1)] + [g.reshape(-1) for g in gp])

    n = y0.numel()

    def aug(tt, state):
        y = state[:n].view_as(y0).detach().requires_grad_(True)
        a = state[n:2 * n].view_as(y0)
        with torch.enable_grad():
            f = func(tt, y)
            vjps = torch.autograd.grad(f, [y] + list(params), -a, allow_unused=True)
        vjps = [torch.zeros_like(x) if v is None else v for v, x in zip(vjps, [y] + list(params))]
        return pack(f.detach(), vjps[0], vjps[1:])

    for i in range(len(t) - 1, 0, -1):
        state = pack(sol[i], adj_y, adj_p)
        tt = torch.stack([t[i], t[i - 1]])
        with torch.no_grad():
            out = _integrate_fixed(aug, state, tt, method, SolverStats())[1]
        adj_y = out[n:2 * n].view_as(y0) + grad_sol[i - 1]
        off = 2 * n
        new_p = []
        for p_, sz in zip(params, sizes):
            new_p.append(out[off:off + sz].view_as(p_)); off += sz
        adj_p = new_p
    return sol, adj_y, adj_p
