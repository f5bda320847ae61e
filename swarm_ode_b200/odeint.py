"""``odeint(func, y0, t, *, rtol, atol, method, options)`` with torchdiffeq's call shape
(reference call sites scripts/train_gde.py:78-85, scripts/gnode.py:136-137,
scripts/run_gnode.py:134-135), running natively in ``libgnode_b200.so``.

``func`` must be one of the vector fields the library implements:

* ``GraphODEFunc.bind(edge_index)`` / ``BoundGraphODEFunc`` -- the graph field over a fixed topology
  (what the reference builds with its ``ode_func_wrapper`` closure);
* ``ODEFunction`` -- the node-wise MLP field.

Arbitrary Python callables are rejected: there is deliberately no eager/CPU integration path.
Returns the solution ``[len(t), *y0.shape]``.  Methods: 'euler', 'midpoint', 'rk4' (torchdiffeq's
3/8-rule rk4) and 'dopri5' (also the default when ``method`` is None, as upstream).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from ._lib import GnodeError
from .modules import BoundGraphODEFunc, ODEFunction

_FIXED = ("euler", "midpoint", "rk4")


def odeint(func, y0: torch.Tensor, t: torch.Tensor, *, rtol: float = 1e-7, atol: float = 1e-9,
           method: Optional[str] = None, options: Optional[dict] = None) -> torch.Tensor:
    method = method or "dopri5"
    options = options or {}
    if method not in _FIXED and method != "dopri5":
        raise ValueError(f'Invalid method "{method}". Must be one of {_FIXED + ("dopri5",)}')
    if torch.is_tensor(t):
        if t.dim() != 1:
            raise ValueError("t must be one dimensional")
        t_list = ops.time_grid_to_host(t, torch.float64)
        if any(b <= a for a, b in zip(t_list[:-1], t_list[1:])):
            raise ValueError("t must be strictly increasing")
    sink = options.get("_stats_sink")

    if isinstance(func, BoundGraphODEFunc):
        graph = func.graph_for(y0)
        params = func.func.param_list()
        if method in _FIXED:
            if options.get("_adjoint"):
                return ops.integrate_fixed_adjoint(y0, graph, params, t, method)
            return ops.integrate_fixed(y0, graph, params, t, method)
        if options.get("_adjoint"):
            raise GnodeError("odeint_adjoint: the adjoint backward exists for the fixed-grid solvers (euler, midpoint, rk4); "
                             "dopri5 differentiates by replaying its accepted steps (odeint)")
        sol, stats = ops.integrate_dopri5(y0, graph, params, t, rtol, atol, allreduce=options.get("allreduce"),
                                          max_num_steps=int(options.get("max_num_steps", 0)))
        if sink is not None:
            sink.last_stats = stats
        return sol

    if isinstance(func, ODEFunction):
        sol, stats = ops.mlp_integrate(y0, func.param_list(), t, method, rtol=rtol, atol=atol,
                                       max_num_steps=int(options.get("max_num_steps", 0)))
        if sink is not None:
            sink.last_stats = stats
        return sol

    raise GnodeError(
        f"odeint: unsupported vector field {type(func).__name__}. Pass GraphODEFunc.bind(edge_index) or an "
        "ODEFunction; arbitrary Python callables would need an eager fallback, which this package does not have.")


def odeint_adjoint(func, y0: torch.Tensor, t: torch.Tensor, *, rtol: float = 1e-7, atol: float = 1e-9,
                   method: Optional[str] = None, options: Optional[dict] = None, adjoint_params=None) -> torch.Tensor:
    """``torchdiffeq.odeint_adjoint`` call shape for the graph field with a fixed-grid solver: same forward as ``odeint``,
    gradients by the adjoint method (the augmented system integrated backwards with the same scheme; O(1) memory in the
    number of steps).  The reference trains with plain ``odeint`` (scripts/train_gde.py:78-85); this is opt-in.
    ``adjoint_params`` is accepted for signature compatibility (the field's own parameters are always used)."""
    if not isinstance(func, BoundGraphODEFunc):
        raise GnodeError("odeint_adjoint: only the graph field (GraphODEFunc.bind(edge_index)) has an adjoint backward")
    opts = dict(options or {})
    opts["_adjoint"] = True
    return odeint(func, y0, t, rtol=rtol, atol=atol, method=method or "rk4", options=opts)
