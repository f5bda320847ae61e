"""Oracle restatement of the reference's heterogeneous GNODE modules.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``) -- parity unpinned for the third-party numerics.

Follows, clause by clause:
  * ``ODEFunction``                         scripts/gnode.py:160-174, scripts/run_gnode.py:153-167
  * ``HeteroGraphODENetwork`` (joint ODE)    scripts/gnode.py:70-158      (``action_size=None``)
  * ``HeteroGraphODENetwork`` (type ODEs)    scripts/run_gnode.py:67-151  (``action_size`` given)
on top of ``oracle.pyg_ref`` (SAGEConv / HeteroConv) and ``oracle.torchdiffeq_ref`` (odeint).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from .pyg_ref import HeteroConvRef, SAGEConvRef
from .torchdiffeq_ref import odeint_ref

EDGE_TYPES = [
    ("agv", "targets", "location"),
    ("location", "is targeted by", "agv"),
    ("agv", "communicates", "agv"),
    ("picker", "manages", "location"),
    ("agv", "cooperates with", "picker"),
    ("picker", "helps", "agv"),
]


class ODEFunctionRef(nn.Module):
    """Linear(H,h) tanh Linear(h,h) tanh Linear(h,H); ``t`` unused -- scripts/gnode.py:160-174."""

    def __init__(self, hidden_dim: int, ode_hidden_dim: int):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(hidden_dim, ode_hidden_dim), nn.Tanh(),
                                 nn.Linear(ode_hidden_dim, ode_hidden_dim), nn.Tanh(),
                                 nn.Linear(ode_hidden_dim, hidden_dim))

    def forward(self, t, x):
        return self.net(x)


class HeteroGraphODENetworkRef(nn.Module):
    """``action_size is None``: scripts/gnode.py:70-158 -- one ODEFunction over the concatenated embeddings, default
    solver (dopri5, rtol 1e-7 / atol 1e-9), heads emit one value.  ``action_size`` given: scripts/run_gnode.py:67-151
    -- one ODEFunction per agent type, ``method='euler'``, locations are not evolved, heads emit ``action_size``."""

    def __init__(self, node_dims: Dict[str, int], action_size: Optional[int] = None, hidden_dim: int = 64,
                 num_layers: int = 2, ode_hidden_dim: int = 32):
        super().__init__()
        self.hidden_dim, self.ode_hidden_dim, self.action_size = hidden_dim, ode_hidden_dim, action_size
        self.agv_embedding = nn.Linear(node_dims["agv"], hidden_dim)
        self.picker_embedding = nn.Linear(node_dims["picker"], hidden_dim)
        self.location_embedding = nn.Linear(node_dims["location"], hidden_dim)
        self.hetero_convs = nn.ModuleList()
        for _ in range(num_layers):
            convs = {et: SAGEConvRef(hidden_dim if et[0] == et[2] else (hidden_dim, hidden_dim), hidden_dim) for et in EDGE_TYPES}
            self.hetero_convs.append(HeteroConvRef(convs, aggr="mean"))
        out = 1 if action_size is None else action_size
        if action_size is None:
            self.ode_func = ODEFunctionRef(hidden_dim, ode_hidden_dim)
        else:
            self.ode_func_agv = ODEFunctionRef(hidden_dim, ode_hidden_dim)
            self.ode_func_picker = ODEFunctionRef(hidden_dim, ode_hidden_dim)
        self.agv_action_head = nn.Sequential(nn.Linear(hidden_dim, hidden_dim // 2), nn.ReLU(), nn.Linear(hidden_dim // 2, out))
        self.picker_action_head = nn.Sequential(nn.Linear(hidden_dim, hidden_dim // 2), nn.ReLU(), nn.Linear(hidden_dim // 2, out))
        self.solver_options: Optional[dict] = None     # test hook, forwarded to the adaptive odeint_ref(options=...)

    def forward(self, hetero_data, integration_time: float = 1.0):
        x_dict = {"agv": self.agv_embedding(hetero_data["agv"].x),
                  "picker": self.picker_embedding(hetero_data["picker"].x),
                  "location": self.location_embedding(hetero_data["location"].x)}
        for conv in self.hetero_convs:
            x_dict = conv(x_dict, hetero_data.edge_index_dict)
            x_dict = {k: torch.relu(v) for k, v in x_dict.items()}
        t = torch.tensor([0.0, integration_time], dtype=torch.float32)
        if self.action_size is None:
            allx = torch.cat([x_dict["agv"], x_dict["picker"], x_dict["location"]], dim=0)
            ev = odeint_ref(self.ode_func, allx, t, options=self.solver_options)[-1]
            na, npk = hetero_data["agv"].num_nodes, hetero_data["picker"].num_nodes
            agv, picker, loc = ev[:na], ev[na:na + npk], ev[na + npk:]
        else:
            agv = odeint_ref(self.ode_func_agv, x_dict["agv"], t, method="euler")[-1]
            picker = odeint_ref(self.ode_func_picker, x_dict["picker"], t, method="euler")[-1]
            loc = x_dict["location"]
        return {"agv_q_values": self.agv_action_head(agv), "picker_q_values": self.picker_action_head(picker),
                "agv_embeddings": agv, "picker_embeddings": picker, "location_embeddings": loc}
