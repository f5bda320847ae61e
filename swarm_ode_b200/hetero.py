"""Heterogeneous GNODE modules of the reference, B200-native (forward / inference path).

* ``HeteroData``                      light stand-in for ``torch_geometric.data.HeteroData`` (PyG objects work too)
* ``HeteroConv(convs, aggr='mean')``  scripts/gnode.py:99, scripts/run_gnode.py:96          [upstream PyG]
* ``HeteroGraphODENetwork``           scripts/gnode.py:70-158 (joint ODE over all embeddings, default solver) and
                                      scripts/run_gnode.py:67-151 (``action_size`` given: one ODE per agent type,
                                      ``method='euler'``, locations are not evolved)

Every FLOP runs in ``libgnode_b200.so``: the embeddings and action heads through the dense NT contraction, each
relation through ``gnode_sage_bipartite_fwd`` (HeteroConv's mean over relations and the ReLU after it are folded into
the relation calls), the MLP field through the native integrators.  The reference uses this network as an RL
Q-function and trains it with loss.backward(): every piece is an autograd node whose backward runs in the library too
(``gnode_linear_bwd``, ``gnode_sage_bipartite_bwd``, the MLP solver backward of ops.mlp_integrate).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import GnodeError
from .graph import CSRGraph
from .modules import ODEFunction, SAGEConv

EdgeType = Tuple[str, str, str]

EDGE_TYPES: List[EdgeType] = [
    ("agv", "targets", "location"),
    ("location", "is targeted by", "agv"),
    ("agv", "communicates", "agv"),
    ("picker", "manages", "location"),
    ("agv", "cooperates with", "picker"),
    ("picker", "helps", "agv"),
]


class _Store:
    def __init__(self):
        self.__dict__["_d"] = {}

    def __getattr__(self, k):
        try:
            return self.__dict__["_d"][k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self.__dict__["_d"][k] = v

    @property
    def num_nodes(self) -> int:
        return int(self.__dict__["_d"]["x"].size(0))


class HeteroData:
    """``data['agv'].x = ...``, ``data['agv', 'targets', 'location'].edge_index = ...``, ``data.edge_index_dict``."""

    def __init__(self):
        self._stores: Dict[object, _Store] = {}

    def __getitem__(self, key):
        if isinstance(key, list):
            key = tuple(key)
        if key not in self._stores:
            self._stores[key] = _Store()
        return self._stores[key]

    @property
    def edge_index_dict(self) -> Dict[EdgeType, torch.Tensor]:
        return {k: s.edge_index for k, s in self._stores.items()
                if isinstance(k, tuple) and "edge_index" in s.__dict__["_d"]}

    @property
    def node_types(self) -> List[str]:
        return [k for k in self._stores if isinstance(k, str)]

    def to(self, device, non_blocking: bool = False) -> "HeteroData":
        for s in self._stores.values():
            d = s.__dict__["_d"]
            for k, v in list(d.items()):
                if torch.is_tensor(v):
                    d[k] = v.to(device, non_blocking=non_blocking)
        self.__dict__.pop("_gnode_csr", None)
        return self


def _relation_graph(edge_index: torch.Tensor, n_src: int, n_dst: int, cache: Optional[dict], key) -> CSRGraph:
    """Destination-sorted CSR of one relation; rows = destination nodes, column ids = source rows."""
    if cache is not None and key in cache:
        ck, g, _keep = cache[key]
        if ck == (edge_index.data_ptr(), edge_index._version, tuple(edge_index.shape), n_src, n_dst):
            return g
    g = CSRGraph(edge_index, max(n_src, n_dst))
    if cache is not None:   # the entry keeps the keyed tensor alive, so its address cannot be recycled meanwhile
        cache[key] = ((edge_index.data_ptr(), edge_index._version, tuple(edge_index.shape), n_src, n_dst), g, edge_index)
    return g


def sage_bipartite(graph: CSRGraph, x_src: torch.Tensor, x_dst: torch.Tensor, conv: SAGEConv, scale: float = 1.0,
                   accum: Optional[torch.Tensor] = None, post_relu: bool = False,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``post(accum + scale * conv((x_src, x_dst), edge_index))`` on the device (no autograd)."""
    x_src = _lib.require_cuda_f32(x_src.detach(), "x_src")
    x_dst = _lib.require_cuda_f32(x_dst.detach(), "x_dst")
    wl, bl, wr = (_lib.require_cuda_f32(t.detach(), "param") for t in (conv.lin_l.weight, conv.lin_l.bias, conv.lin_r.weight))
    return _sage_bipartite_raw(graph, x_src, x_dst, wl, bl, wr, scale, accum, post_relu, out)


def _sage_bipartite_raw(graph: CSRGraph, x_src, x_dst, wl, bl, wr, scale: float = 1.0, accum=None, post_relu: bool = False,
                        out=None) -> torch.Tensor:
    n_dst, ci = x_dst.shape
    co = wl.shape[0]
    if x_src.shape[1] != ci or wl.shape[1] != ci or wr.shape[1] != ci:
        raise GnodeError("sage_bipartite: source / destination / weight widths differ")
    if out is None:
        out = torch.empty((n_dst, co), dtype=torch.float32, device=x_dst.device)
    L = _lib.lib()
    ws = _lib.WORKSPACE.get(L.gnode_sage_bipartite_workspace_bytes(n_dst, ci, co), x_dst.device)
    with torch.cuda.device(x_dst.device):
        _lib.check(L.gnode_sage_bipartite_fwd(graph.ref(), n_dst, _lib.ptr(x_src), _lib.ptr(x_dst), ci, co, _lib.ptr(wl),
                                              _lib.ptr(bl), _lib.ptr(wr), float(scale), _lib.ptr(accum), int(post_relu),
                                              _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(x_dst.device)),
                   "gnode_sage_bipartite_fwd")
    return out


class HeteroConv(nn.Module):
    """``HeteroConv({edge_type: SAGEConv}, aggr='mean')``: every relation present in both the module dict and
    ``edge_index_dict`` runs on ``(x_src, x_dst)``; results per destination type are averaged.  ``relu=True`` folds the
    ReLU the reference applies right after (scripts/gnode.py:127-128) into the last relation of each type."""

    def __init__(self, convs: Dict[EdgeType, SAGEConv], aggr: str = "mean"):
        super().__init__()
        if aggr != "mean":
            raise NotImplementedError("only aggr='mean' (the reference's setting) is implemented")
        self.edge_types = list(convs.keys())
        self.convs = nn.ModuleDict({"__".join(k): v for k, v in convs.items()})

    def forward(self, x_dict: Dict[str, torch.Tensor], edge_index_dict: Dict[EdgeType, torch.Tensor],
                relu: bool = False, cache: Optional[dict] = None) -> Dict[str, torch.Tensor]:
        active: Dict[str, List[EdgeType]] = {}
        for et in self.edge_types:
            src, _rel, dst = et
            if et in edge_index_dict and src in x_dict and dst in x_dict:
                active.setdefault(dst, []).append(et)
        out: Dict[str, torch.Tensor] = {}
        needs_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or
                                                  any(v.requires_grad for v in x_dict.values()))
        for dst, ets in active.items():
            if needs_grad:
                graphs, rest = [], []
                for et in ets:
                    conv = self.convs["__".join(et)]
                    graphs.append(_relation_graph(edge_index_dict[et], x_dict[et[0]].size(0), x_dict[dst].size(0), cache, et))
                    rest += [x_dict[et[0]], conv.lin_l.weight, conv.lin_l.bias, conv.lin_r.weight]
                out[dst] = _RelationGroupFn.apply(graphs, relu, x_dict[dst], *rest)
                continue
            acc = None
            for i, et in enumerate(ets):
                src = et[0]
                g = _relation_graph(edge_index_dict[et], x_dict[src].size(0), x_dict[dst].size(0), cache, et)
                last = i == len(ets) - 1
                acc = sage_bipartite(g, x_dict[src], x_dict[dst], self.convs["__".join(et)], scale=1.0 / len(ets),
                                     accum=acc, post_relu=relu and last, out=acc)
            out[dst] = acc
        return out


class _LinearFn(torch.autograd.Function):
    """``act(x @ w.T + b)`` with the backward in the library (gnode_linear_bwd)."""

    @staticmethod
    def forward(ctx, x, w, b, act: str):
        xc = _lib.require_cuda_f32(x.detach(), "x")
        wc = _lib.require_cuda_f32(w.detach(), "weight")
        bc = _lib.require_cuda_f32(b.detach(), "bias")
        out = ops.gemm_nt(xc, wc, bias=bc, act=act)
        ctx.act = act
        ctx.save_for_backward(xc, wc, out)
        return out

    @staticmethod
    def backward(ctx, g):
        x, w, out = ctx.saved_tensors
        g = _lib.require_cuda_f32(g.contiguous(), "grad_out")
        m, ci = x.shape
        co = w.shape[0]
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gw = torch.zeros_like(w)
        gb = torch.zeros(co, dtype=torch.float32, device=x.device)
        L = _lib.lib()
        ws = _lib.WORKSPACE.get(L.gnode_linear_bwd_workspace_bytes(m, ci, co), x.device)
        with torch.cuda.device(x.device):
            _lib.check(L.gnode_linear_bwd(_lib.ptr(x), _lib.ptr(w), _lib.ptr(out if ctx.act == "relu" else None), _lib.ptr(g),
                                          m, ci, co, _lib.ptr(gx), _lib.ptr(gw), _lib.ptr(gb), _lib.ptr(ws), ws.numel(),
                                          _lib.stream_ptr(x.device)), "gnode_linear_bwd")
        return gx, gw, gb, None


def _linear(x: torch.Tensor, lin: nn.Linear, act: str = "none") -> torch.Tensor:
    if torch.is_grad_enabled() and (x.requires_grad or lin.weight.requires_grad or lin.bias.requires_grad):
        return _LinearFn.apply(x, lin.weight, lin.bias, act)
    return ops.gemm_nt(x, lin.weight, bias=lin.bias, act=act)


class _RelationGroupFn(torch.autograd.Function):
    """All relations of one destination type of a HeteroConv layer: ``relu?(mean_r conv_r((x_src_r, x_dst), edges_r))``.
    Inputs after the fixed ones: x_dst, then per relation (x_src, lin_l.weight, lin_l.bias, lin_r.weight)."""

    @staticmethod
    def forward(ctx, graphs, relu: bool, x_dst, *rest):
        R = len(graphs)
        xd = _lib.require_cuda_f32(x_dst.detach(), "x_dst")
        tens = [_lib.require_cuda_f32(t.detach(), "relation input") for t in rest]
        acc = None
        for r in range(R):
            xs, wl, bl, wr = tens[4 * r:4 * r + 4]
            acc = _sage_bipartite_raw(graphs[r], xs, xd, wl, bl, wr, scale=1.0 / R, accum=acc,
                                      post_relu=relu and r == R - 1, out=acc)
        ctx.graphs, ctx.relu = graphs, relu
        ctx.save_for_backward(xd, acc, *tens)
        return acc

    @staticmethod
    def backward(ctx, g):
        xd, out, *tens = ctx.saved_tensors
        R = len(ctx.graphs)
        g = _lib.require_cuda_f32(g.contiguous(), "grad_out")
        n_dst, ci = xd.shape
        co = out.shape[1]
        L = _lib.lib()
        ws = _lib.WORKSPACE.get(L.gnode_sage_bipartite_bwd_workspace_bytes(n_dst, ci, co), xd.device)
        gxd_total = None
        grads = []
        with torch.cuda.device(xd.device):
            for r in range(R):
                xs, wl, bl, wr = tens[4 * r:4 * r + 4]
                gxs = torch.empty_like(xs)
                gxd = torch.empty_like(xd)
                gwl, gbl, gwr = torch.zeros_like(wl), torch.zeros_like(bl), torch.zeros_like(wr)
                _lib.check(L.gnode_sage_bipartite_bwd(ctx.graphs[r].ref(), xs.shape[0], n_dst, _lib.ptr(xs), _lib.ptr(xd), ci, co,
                                                      _lib.ptr(wl), _lib.ptr(wr), _lib.ptr(g),
                                                      _lib.ptr(out if ctx.relu else None), 1.0 / R, _lib.ptr(gxs),
                                                      _lib.ptr(gxd), _lib.ptr(gwl), _lib.ptr(gbl), _lib.ptr(gwr), _lib.ptr(ws),
                                                      ws.numel(), _lib.stream_ptr(xd.device)), "gnode_sage_bipartite_bwd")
                gxd_total = gxd if gxd_total is None else gxd_total.add_(gxd)
                grads += [gxs, gwl, gbl, gwr]
        return (None, None, gxd_total, *grads)


class HeteroGraphODENetwork(nn.Module):
    """Heterogeneous graph neural ODE Q-network (scripts/gnode.py:70-158; scripts/run_gnode.py:67-151).

    ``HeteroGraphODENetwork(node_dims, hidden_dim=64, num_layers=2, ode_hidden_dim=32)`` is the scripts/gnode.py form;
    passing ``action_size`` (positionally second, as in scripts/run_gnode.py:70) selects the run_gnode form.
    ``forward(hetero_data, integration_time=1.0)`` returns the reference's dict of five tensors."""

    def __init__(self, node_dims: Dict[str, int], action_size: Optional[int] = None, hidden_dim: int = 64,
                 num_layers: int = 2, ode_hidden_dim: int = 32):
        super().__init__()
        self.hidden_dim, self.ode_hidden_dim, self.action_size = hidden_dim, ode_hidden_dim, action_size
        self.agv_dim, self.picker_dim, self.location_dim = node_dims["agv"], node_dims["picker"], node_dims["location"]
        self.agv_embedding = nn.Linear(self.agv_dim, hidden_dim)
        self.picker_embedding = nn.Linear(self.picker_dim, hidden_dim)
        self.location_embedding = nn.Linear(self.location_dim, hidden_dim)
        self.hetero_convs = nn.ModuleList()
        for _ in range(num_layers):
            convs = {et: SAGEConv(hidden_dim if et[0] == et[2] else (hidden_dim, hidden_dim), hidden_dim) for et in EDGE_TYPES}
            self.hetero_convs.append(HeteroConv(convs, aggr="mean"))
        n_out = 1 if action_size is None else action_size
        if action_size is None:
            self.ode_func = ODEFunction(hidden_dim, ode_hidden_dim)
        else:
            self.ode_func_agv = ODEFunction(hidden_dim, ode_hidden_dim)
            self.ode_func_picker = ODEFunction(hidden_dim, ode_hidden_dim)
        self.agv_action_head = nn.Sequential(nn.Linear(hidden_dim, hidden_dim // 2), nn.ReLU(), nn.Linear(hidden_dim // 2, n_out))
        self.picker_action_head = nn.Sequential(nn.Linear(hidden_dim, hidden_dim // 2), nn.ReLU(), nn.Linear(hidden_dim // 2, n_out))
        self.last_stats = None

    def _head(self, x: torch.Tensor, head: nn.Sequential) -> torch.Tensor:
        return _linear(_linear(x, head[0], "relu"), head[2])

    def forward(self, hetero_data, integration_time: float = 1.0) -> Dict[str, torch.Tensor]:
        x_dict = {"agv": _linear(hetero_data["agv"].x, self.agv_embedding),
                  "picker": _linear(hetero_data["picker"].x, self.picker_embedding),
                  "location": _linear(hetero_data["location"].x, self.location_embedding)}
        eid = hetero_data.edge_index_dict
        cache = getattr(hetero_data, "__dict__", {}).setdefault("_gnode_csr", {}) if hasattr(hetero_data, "__dict__") else None
        for conv in self.hetero_convs:
            # reference: x = conv(x, edges); x = relu(x)  -- the ReLU is fused into the last relation of every type
            x_dict = conv(x_dict, eid, relu=True, cache=cache)
        dev = x_dict["agv"].device
        t = [0.0, float(integration_time)]
        if self.action_size is None:
            na, npk = x_dict["agv"].size(0), x_dict["picker"].size(0)
            allx = torch.cat([x_dict["agv"], x_dict["picker"], x_dict["location"]], dim=0)
            sol, self.last_stats = ops.mlp_integrate(allx, self.ode_func.param_list(), t, "dopri5", rtol=1e-7, atol=1e-9)
            ev = sol[-1]
            agv, picker, loc = ev[:na], ev[na:na + npk], ev[na + npk:]
        else:
            agv = ops.mlp_integrate(x_dict["agv"], self.ode_func_agv.param_list(), t, "euler")[0][-1]
            picker = ops.mlp_integrate(x_dict["picker"], self.ode_func_picker.param_list(), t, "euler")[0][-1]
            loc = x_dict["location"]
        return {"agv_q_values": self._head(agv, self.agv_action_head),
                "picker_q_values": self._head(picker, self.picker_action_head),
                "agv_embeddings": agv, "picker_embeddings": picker, "location_embeddings": loc}
