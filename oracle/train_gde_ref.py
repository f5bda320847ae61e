"""Oracle restatement of the reference's GraphODE path (scripts/train_gde.py).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``) -- parity unpinned.

Follows, clause by clause:
  * ``GraphODEFunc``            scripts/train_gde.py:20-45
  * ``GraphODE``                scripts/train_gde.py:47-106
  * ``GraphConverter``          scripts/train_gde.py:108-271
  * ``collate_trajectory_batches``  scripts/train_gde.py:363-375
  * the training step           scripts/train_gde.py:474-495
on top of ``oracle.pyg_ref`` (SAGEConv / Batch) and ``oracle.torchdiffeq_ref`` (odeint).
"""
from __future__ import annotations

from collections import deque
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .pyg_ref import RefBatch, RefData, SAGEConvRef
from .torchdiffeq_ref import SolverStats, odeint_ref


class GraphODEFuncRef(nn.Module):
    """dx/dt = conv3(relu(conv2(relu(conv1(x)))))  -- scripts/train_gde.py:20-45.
    ``num_layers`` is accepted and ignored exactly as in the reference (:21-29); ``t`` is unused."""

    def __init__(self, node_dim: int, hidden_dim: int = 64, num_layers: int = 2):
        super().__init__()
        self.node_dim = node_dim
        self.hidden_dim = hidden_dim
        self.conv1 = SAGEConvRef(node_dim, hidden_dim)
        self.conv2 = SAGEConvRef(hidden_dim, hidden_dim)
        self.conv3 = SAGEConvRef(hidden_dim, node_dim)

    def forward(self, t, x, edge_index):
        h = torch.relu(self.conv1(x, edge_index))
        h = torch.relu(self.conv2(h, edge_index))
        return self.conv3(h, edge_index)


class GraphODERef(nn.Module):
    """scripts/train_gde.py:47-106.  Returns the same dict; ``last_stats`` additionally keeps the
    solver bookkeeping (accepted / attempted steps) of the latest call."""

    def __init__(self, node_dim: int, num_agvs: int, num_pickers: int, hidden_dim: int = 64,
                 ode_solver: str = "euler"):
        super().__init__()
        self.node_dim = node_dim
        self.num_agvs = num_agvs
        self.num_pickers = num_pickers
        self.ode_solver = ode_solver
        self.ode_func = GraphODEFuncRef(node_dim=node_dim, hidden_dim=hidden_dim)
        self.position_decoder = nn.Linear(node_dim, 2)
        self.last_stats: Optional[SolverStats] = None
        self.solver_options: Optional[dict] = None     # test hook, forwarded to odeint_ref(options=...)

    def forward(self, batch_data, time_span: torch.Tensor) -> Dict[str, torch.Tensor]:
        x0 = batch_data.x
        edge_index = batch_data.edge_index
        stats = SolverStats()
        solution = odeint_ref(lambda t, x: self.ode_func(t, x, edge_index), x0, time_span,
                              method=self.ode_solver, rtol=1e-3, atol=1e-4, stats=stats,
                              options=self.solver_options)
        self.last_stats = stats
        trajectories = torch.stack([self.position_decoder(solution[i]) for i in range(solution.size(0))], dim=0)
        return {"trajectories": trajectories, "node_features": solution, "batch": batch_data.batch}

    def predict_trajectory(self, batch_data, num_steps: int, dt: float = 0.1) -> torch.Tensor:
        time_span = torch.arange(0, num_steps + 1, dtype=torch.float32)
        return self.forward(batch_data, time_span)["trajectories"]


# ----------------------------------------------------------------------------
# Graph construction -- integer / exact
# ----------------------------------------------------------------------------
class GraphConverterRef:
    """scripts/train_gde.py:108-271, restated.

    State: a window (deque, maxlen ``temporal_window``) of per-step snapshots
    ``(x_t [n, D] f32, local spatial edges [2, E_t] int64)``.

    Per call (``:116-184``):
      positions (y, x) = columns (3, 4) of AGV rows (row < num_agvs) else (0, 1)   (:186-208)
      spatial edges: for i < j, if sqrt(sum((p_i - p_j)**2)) < threshold (strict) emit
      [i, j] then [j, i]                                                          (:228-244)
      k = len(window) - 1 after the append (0..W-1, then stays W-1)               (:136)
      x        = concat of the window's x_t                                       (:140-143)
      edges    = spatial(0)+0*n, ..., spatial(k-1)+(k-1)*n  [always appended, even if empty],
                 spatial(k)+k*n  [only if non-empty],
                 temporal (k-1)*n+a -> k*n+a for a = 0..n-1 [only if k > 0]       (:146-166, :246-267)
      is_current_agent = True on [k*n, (k+1)*n)                                   (:174-179)
    """

    def __init__(self, num_agvs: int, num_pickers: int, distance_threshold: float = 3.0,
                 temporal_window: int = 5):
        self.num_agvs = num_agvs
        self.num_pickers = num_pickers
        self.distance_threshold = distance_threshold
        self.temporal_window = temporal_window
        self.graph_history = deque(maxlen=temporal_window)

    def reset_history(self):
        self.graph_history.clear()

    # -- pieces -------------------------------------------------------------
    @staticmethod
    def standardize(observations) -> np.ndarray:
        """:210-226 -- ragged (object array / list) rows are zero-padded to the longest row."""
        if isinstance(observations, np.ndarray) and observations.dtype != object:
            return observations
        rows = observations.tolist() if isinstance(observations, np.ndarray) else list(observations)
        width = max(len(r) for r in rows)
        out = np.zeros((len(rows), width), dtype=np.float32)
        for i, r in enumerate(rows):
            r = np.asarray(r, dtype=np.float32)
            out[i, :len(r)] = r
        return out

    def locations(self, obs: np.ndarray) -> np.ndarray:
        n = obs.shape[0]
        loc = np.empty((n, 2), dtype=obs.dtype)
        is_agv = np.arange(n) < self.num_agvs
        loc[is_agv] = obs[is_agv][:, [3, 4]]
        loc[~is_agv] = obs[~is_agv][:, [0, 1]]
        return loc

    def spatial_edges(self, loc: np.ndarray) -> torch.Tensor:
        n = loc.shape[0]
        pairs = []
        for i in range(n):
            for j in range(i + 1, n):
                # same expression and dtype as the reference (:236): sqrt(sum(diff**2)) in the obs dtype
                d = np.sqrt(np.sum((loc[i] - loc[j]) ** 2))
                if d < self.distance_threshold:
                    pairs.append((i, j))
                    pairs.append((j, i))
        if not pairs:
            return torch.empty((2, 0), dtype=torch.long)
        return torch.tensor(pairs, dtype=torch.long).t()

    # -- the call -------------------------------------------------------------
    def build(self, observations) -> RefData:
        obs = self.standardize(observations)
        n = len(obs)
        x_t = torch.tensor(obs, dtype=torch.float32)
        e_t = self.spatial_edges(self.locations(obs))
        self.graph_history.append((x_t, e_t))
        k = len(self.graph_history) - 1

        x = torch.cat([g[0] for g in self.graph_history], dim=0)
        parts = [self.graph_history[i][1] + i * n for i in range(k)]
        if e_t.shape[1] > 0:
            parts.append(e_t + k * n)
        if k > 0:
            a = torch.arange(n, dtype=torch.long)
            parts.append(torch.stack([(k - 1) * n + a, k * n + a], dim=0))
        edge_index = torch.cat(parts, dim=1) if parts else torch.empty((2, 0), dtype=torch.long)

        mask = torch.zeros(x.size(0), dtype=torch.bool)
        mask[k * n:(k + 1) * n] = True
        return RefData(x=x, edge_index=edge_index, is_current_agent=mask)

    # reference method name
    _build_graph_from_observation = build


def extract_positions_ref(graph: RefData, num_agvs: int, num_pickers: int) -> torch.Tensor:
    """scripts/train_gde.py:336-355 -- (x, y) of the FIRST num_agvs / num_pickers rows of the
    window-concatenated graph (i.e. the oldest snapshot; reference quirk, kept)."""
    parts = []
    if num_agvs > 0:
        parts.append(graph.x[:num_agvs][:, [4, 3]])
    if num_pickers > 0:
        parts.append(graph.x[num_agvs:num_agvs + num_pickers][:, [1, 0]])
    return torch.cat(parts, dim=0)


class TrajectoryBatchRef:
    def __init__(self, graphs, next_positions):
        self.graphs = graphs
        self.next_positions = next_positions


def collate_ref(batch_list: List[TrajectoryBatchRef]) -> TrajectoryBatchRef:
    """scripts/train_gde.py:363-375."""
    graphs = RefBatch.from_data_list([b.graphs for b in batch_list])
    nxt = torch.stack([b.next_positions for b in batch_list], dim=0)
    return TrajectoryBatchRef(graphs, nxt)


def train_step_loss_ref(model: GraphODERef, graphs, next_positions: torch.Tensor,
                        time_span: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Loss of scripts/train_gde.py:482-490 (mse on the current-agent rows of trajectories[1])."""
    if time_span is None:
        time_span = torch.tensor([0.0, 1.0])
    pred = model(graphs, time_span)["trajectories"][1]
    return F.mse_loss(pred[graphs.is_current_agent], next_positions.view(-1, 2))
