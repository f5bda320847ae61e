"""Drop-in ``nn.Module``s for the reference's GNODE path.

Same class names, constructor arguments, ``forward`` signatures, return dicts and ``state_dict``
keys as the reference, so reference checkpoints load and reference call sites keep working:

* ``SAGEConv(in, out)(x, edge_index)``                      [upstream PyG; scripts/train_gde.py:27-29]
* ``GraphODEFunc(node_dim, hidden_dim=64, num_layers=2).forward(t, x, edge_index)``
                                                            scripts/train_gde.py:20-45
* ``GraphODE(node_dim, num_agvs, num_pickers, hidden_dim=64, ode_solver='euler')
     .forward(batch_data, time_span) -> {'trajectories', 'node_features', 'batch'}``,
  ``.predict_trajectory(batch_data, num_steps, dt)``        scripts/train_gde.py:47-106
* ``ODEFunction(hidden_dim, ode_hidden_dim).forward(t, x)`` scripts/gnode.py:160-174

All arithmetic runs in ``libgnode_b200.so`` (sm_100a); there is no CPU fallback.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from ._lib import GnodeError
from .graph import CSRGraph, csr_for


class _Linear(nn.Module):
    """Parameter holder with PyG ``Linear`` naming / init (weight [out, in])."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1.0 / math.sqrt(self.in_channels) if self.in_channels > 0 else 0.0
            nn.init.uniform_(self.bias, -bound, bound)


class SAGEConv(nn.Module):
    """Mean-aggregating GraphSAGE layer: ``lin_l(mean_j x_j) + lin_r(x_i)``.

    ``edge_index`` may be a PyG-style int64 [2, E] tensor or an already built ``CSRGraph``.
    """

    def __init__(self, in_channels: int, out_channels: int, aggr: str = "mean", normalize: bool = False,
                 root_weight: bool = True, project: bool = False, bias: bool = True):
        super().__init__()
        self.bipartite = isinstance(in_channels, (tuple, list))   # used on (x_src, x_dst) pairs by HeteroConv
        if self.bipartite:
            if in_channels[0] != in_channels[1]:
                raise NotImplementedError("bipartite SAGEConv with different source/target widths")
            in_channels = in_channels[0]
        if aggr != "mean" or normalize or not root_weight or project or not bias:
            raise NotImplementedError("only the PyG defaults used by the reference are implemented")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_l = _Linear(in_channels, out_channels, bias=True)
        self.lin_r = _Linear(in_channels, out_channels, bias=False)

    def forward(self, x: torch.Tensor, edge_index, relu: bool = False) -> torch.Tensor:
        graph = edge_index if isinstance(edge_index, CSRGraph) else csr_for(edge_index, x.size(0))
        return ops.sage_conv(x, self.lin_l.weight, self.lin_l.bias, self.lin_r.weight, graph, relu)


class GraphODEFunc(nn.Module):
    """dx/dt = conv3(relu(conv2(relu(conv1(x)))))   (scripts/train_gde.py:20-45).

    ``num_layers`` is accepted and ignored exactly as in the reference; ``t`` is unused.
    """

    def __init__(self, node_dim: int, hidden_dim: int = 64, num_layers: int = 2):
        super().__init__()
        self.node_dim = node_dim
        self.hidden_dim = hidden_dim
        self.conv1 = SAGEConv(node_dim, hidden_dim)
        self.conv2 = SAGEConv(hidden_dim, hidden_dim)
        self.conv3 = SAGEConv(hidden_dim, node_dim)
        self.activation = nn.ReLU()

    def param_list(self) -> List[torch.Tensor]:
        c1, c2, c3 = self.conv1, self.conv2, self.conv3
        return [c1.lin_l.weight, c1.lin_l.bias, c1.lin_r.weight,
                c2.lin_l.weight, c2.lin_l.bias, c2.lin_r.weight,
                c3.lin_l.weight, c3.lin_l.bias, c3.lin_r.weight]

    def forward(self, t, x: torch.Tensor, edge_index) -> torch.Tensor:
        graph = edge_index if isinstance(edge_index, CSRGraph) else csr_for(edge_index, x.size(0))
        return ops.gnode_rhs(x, graph, self.param_list())

    def bind(self, edge_index, num_nodes: Optional[int] = None) -> "BoundGraphODEFunc":
        """``(t, x)`` closure over a fixed graph that ``swarm_ode_b200.odeint`` runs natively."""
        return BoundGraphODEFunc(self, edge_index, num_nodes)


class BoundGraphODEFunc:
    """The reference's ``ode_func_wrapper(t, x)`` closure (scripts/train_gde.py:74-75) as an object the
    native integrator can recognise."""

    def __init__(self, func: GraphODEFunc, edge_index, num_nodes: Optional[int] = None):
        self.func = func
        self.edge_index = edge_index
        self.num_nodes = num_nodes

    def graph_for(self, x: torch.Tensor) -> CSRGraph:
        if isinstance(self.edge_index, CSRGraph):
            return self.edge_index
        return csr_for(self.edge_index, self.num_nodes if self.num_nodes is not None else x.size(0))

    def __call__(self, t, x):
        return self.func(t, x, self.graph_for(x))


class GraphODE(nn.Module):
    """Graph neural ODE for trajectory prediction (scripts/train_gde.py:47-106)."""

    def __init__(self, node_dim: int, num_agvs: int, num_pickers: int, hidden_dim: int = 64,
                 ode_solver: str = "euler", adjoint: bool = False):
        super().__init__()
        # adjoint=True (extension; the reference has no such switch): gradients by the adjoint method instead of backprop
        # through the solver -- odeint_adjoint semantics, fixed-grid solvers only
        self.adjoint = bool(adjoint)
        self.node_dim = node_dim
        self.num_agvs = num_agvs
        self.num_pickers = num_pickers
        self.ode_solver = ode_solver
        self.ode_func = GraphODEFunc(node_dim=node_dim, hidden_dim=hidden_dim)
        self.position_decoder = nn.Linear(node_dim, 2)
        self.last_stats = None   # Dopri5Stats of the latest adaptive solve
        self.dopri5_allreduce = None  # optional cross-rank hook, see ops.integrate_dopri5

    def forward(self, batch_data, time_span: torch.Tensor) -> Dict[str, torch.Tensor]:
        from .odeint import odeint  # local import: odeint imports this module

        x0 = batch_data.x
        edge_index = batch_data.edge_index
        batch = batch_data.batch
        # deferred index validation: no host synchronisation in the training step (a bad edge list raises at the
        # next call; its edges are skipped on the device meanwhile)
        capturing = x0.is_cuda and torch.cuda.is_current_stream_capturing()
        graph = csr_for(edge_index, x0.size(0), holder=batch_data, validate="device" if capturing else "deferred",
                        graph_ptr=getattr(batch_data, "ptr", None),
                        max_graph_nodes=getattr(batch_data, "max_graph_nodes", None))
        if self.adjoint:
            from .odeint import odeint_adjoint
            solution = odeint_adjoint(self.ode_func.bind(graph), x0, time_span, method=self.ode_solver)
            trajectories = ops.decode_positions(solution, self.position_decoder.weight, self.position_decoder.bias)
            return {"trajectories": trajectories, "node_features": solution, "batch": batch}
        if self.ode_solver in ("euler", "midpoint", "rk4") and self.position_decoder.out_features <= 8:
            # solver + decoder as one autograd node: lets the backward pass use the factored cotangent of the
            # reference's training loss (see ops._IntegrateDecodeFn)
            if torch.is_tensor(time_span) and time_span.dim() != 1:
                raise ValueError("t must be one dimensional")
            solution, trajectories = ops.integrate_fixed_decode(x0, graph, self.ode_func.param_list(), time_span,
                                                                self.ode_solver, self.position_decoder.weight,
                                                                self.position_decoder.bias)
            return {"trajectories": trajectories, "node_features": solution, "batch": batch}
        solution = odeint(self.ode_func.bind(graph), x0, time_span, method=self.ode_solver, rtol=1e-3, atol=1e-4,
                          options={"_stats_sink": self, "allreduce": self.dopri5_allreduce})
        trajectories = ops.decode_positions(solution, self.position_decoder.weight, self.position_decoder.bias)
        return {"trajectories": trajectories, "node_features": solution, "batch": batch}

    def predict_trajectory(self, batch_data, num_steps: int, dt: float = 0.1) -> torch.Tensor:
        time_span = torch.arange(0, num_steps + 1, dtype=torch.float32)
        return self.forward(batch_data, time_span)["trajectories"]


class ODEFunction(nn.Module):
    """Node-wise MLP field ``Linear(H,h) tanh Linear(h,h) tanh Linear(h,H)`` (scripts/gnode.py:160-174)."""

    def __init__(self, hidden_dim: int, ode_hidden_dim: int):
        super().__init__()
        self.net = nn.Sequential(
            nn.Linear(hidden_dim, ode_hidden_dim), nn.Tanh(),
            nn.Linear(ode_hidden_dim, ode_hidden_dim), nn.Tanh(),
            nn.Linear(ode_hidden_dim, hidden_dim))

    def param_list(self) -> List[torch.Tensor]:
        return [self.net[0].weight, self.net[0].bias, self.net[2].weight, self.net[2].bias,
                self.net[4].weight, self.net[4].bias]

    def forward(self, t, x: torch.Tensor) -> torch.Tensor:
        return ops.mlp_rhs(x, self.param_list())
