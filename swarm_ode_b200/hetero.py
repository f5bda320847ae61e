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

    @classmethod
    def from_data_list(cls, data_list: "List[HeteroData]") -> "HeteroData":
        """Disjoint union of many heterogeneous graphs (what ``torch_geometric.data.Batch.from_data_list`` does for
        ``HeteroData`` [upstream PyG]): node features of every type concatenated, every relation's ``edge_index``
        shifted by the node counts of the graphs before it, plus ``ptr`` per node type.  One forward of
        ``HeteroGraphODENetwork`` over the union equals the concatenation of the per-graph forwards (no edge crosses
        graphs) -- the batched form of the reference's one-graph-per-call use (scripts/run_gnode.py:115-151)."""
        out = cls()
        node_types: List[str] = []
        edge_types: List[tuple] = []
        for d in data_list:
            for k in d._stores:
                (node_types if isinstance(k, str) else edge_types).append(k) if k not in node_types and k not in edge_types else None
        counts = {nt: [d[nt].num_nodes if nt in d._stores and "x" in d[nt].__dict__["_d"] else 0 for d in data_list] for nt in node_types}
        for nt in node_types:
            xs = [d[nt].x for d in data_list if nt in d._stores and "x" in d[nt].__dict__["_d"]]
            out[nt].x = torch.cat(xs, dim=0)
            ptr = torch.zeros(len(data_list) + 1, dtype=torch.long)
            ptr[1:] = torch.cumsum(torch.tensor(counts[nt], dtype=torch.long), 0)
            out[nt].ptr = ptr
        for et in edge_types:
            src, _rel, dst = et
            parts = []
            for g, d in enumerate(data_list):
                if et in d._stores and "edge_index" in d[et].__dict__["_d"]:
                    ei = d[et].edge_index
                    off = torch.tensor([[int(out[src].ptr[g])], [int(out[dst].ptr[g])]], dtype=ei.dtype, device=ei.device)
                    parts.append(ei + off)
            out[et].edge_index = torch.cat(parts, dim=1) if parts else torch.empty((2, 0), dtype=torch.long)
        out.num_graphs = len(data_list)
        return out

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


# ----------------------------------------------------------------------------------------------
# MultiAgentGraphConverter (scripts/run_gnode.py:1040-1326)
# ----------------------------------------------------------------------------------------------
class MultiAgentGraphConverter:
    """Drop-in for the reference's observation -> ``HeteroData`` converter (scripts/run_gnode.py:1040-1326): same
    constructor, same ``_build_graph_from_observation(observation, rack_locations)``, same node features, the same six
    relations with the same edges in the same order -- built with vectorised numpy instead of the reference's Python
    pair loops.  Checked bit for bit against golden vectors produced by the reference's own class
    (scripts/make_golden.py, tests/test_hetero_converter.py).

    What the reference's code actually does (and this class reproduces, because a drop-in must):

    * ``position_to_sections`` is emptied at the start of every call and refilled only AFTER the edges were built
      (:1075, :1105-1106), so every section lookup during edge construction returns ``None``:
      - a picker without a target is connected to every requested shelf (``None == None``, :1256-1270);
      - an AGV WITH a target is connected to every picker (``agv_target_in_picker_section``, :1310-1314);
      - two AGVs that both have a target raise ``KeyError`` (``_check_same_rack_group`` indexes the empty dict, :1322);
    * a picker with a target compares a 3-tuple rack entry with a 2-vector (:1267): ``ValueError`` under numpy >= 2;
    * ``rack_locations`` entries must be hashable ``(x, y, group)`` tuples (:1263): ndarray rows raise ``TypeError``;
    * ``_current_agents_info`` / ``_current_shelves_info`` are appended to on every call and never cleared (:1088,
      :1101), while the edge builders index them from 0: a converter that is REUSED keeps building the edges of its
      FIRST observation.  ``fresh=True`` clears them per call instead (the mode a maintainer would want; the reference's
      behaviour is the default).
    """

    def __init__(self, num_agvs, num_pickers, topk_tasks=5, max_comm_distance=5.0, max_task_distance=10.0, fresh: bool = False):
        self.topk_tasks = topk_tasks
        self.max_comm_distance = max_comm_distance
        self.max_task_distance = max_task_distance
        self.num_agv_nodes = num_agvs
        self.num_picker_nodes = num_pickers
        self.num_location_nodes = None
        self.agv_feature_dim = 7       # [carrying_shelf, carrying_requested, toggle_loading, pos_y, pos_x, target_y, target_x]
        self.picker_feature_dim = 4    # [pos_y, pos_x, target_y, target_x]
        self.location_feature_dim = 2  # [has_shelf, is_requested]
        self.fresh = fresh
        self._current_agents_info: list = []
        self._current_shelves_info: list = []
        self._rack_locations: list = []
        self.position_to_sections: dict = {}
        self.edge_list = None

    def reset(self):
        self._current_agents_info, self._current_shelves_info = [], []

    def _build_graph_from_observation(self, observation, rack_locations) -> "HeteroData":
        import numpy as np
        na, npk = self.num_agv_nodes, self.num_picker_nodes
        self.num_location_nodes = len(rack_locations)
        self._rack_locations = rack_locations
        self.position_to_sections = {}
        if self.fresh:
            self.reset()
        obs = [np.asarray(o) for o in observation]
        agv_features = [obs[a][:self.agv_feature_dim].tolist() for a in range(min(na, len(obs)))]
        picker_features = [obs[a][:self.picker_feature_dim].tolist() for a in range(na, len(obs))]
        self._current_agents_info.extend(agv_features + picker_features)
        shelf_data = obs[0][7 + 4 * (na + npk - 1):]
        n_pairs = (len(shelf_data) + 1) // 2
        if len(shelf_data) % 2:
            raise IndexError("index %d is out of bounds for axis 0 with size %d" % (len(shelf_data), len(shelf_data)))
        location_features = [[shelf_data[2 * i], shelf_data[2 * i + 1]] for i in range(n_pairs)]
        self._current_shelves_info.extend([v for pair in location_features for v in pair])
        self.edge_list = self._build_edges()
        for (x, y, group_idx) in self._rack_locations:
            self.position_to_sections[(x, y)] = group_idx

        data = HeteroData()
        for name, feats, n, dim in (("agv", agv_features, na, self.agv_feature_dim), ("picker", picker_features, npk, self.picker_feature_dim),
                                    ("location", location_features, self.num_location_nodes, self.location_feature_dim)):
            if feats:
                data[name].num_nodes = n
                data[name].x = torch.tensor(np.asarray(feats, dtype=np.float64), dtype=torch.float32)
            else:
                data[name].num_nodes = 0
                data[name].x = torch.empty((0, dim), dtype=torch.float32)
        for et, edges in zip(EDGE_TYPES, self.edge_list):
            data[et].edge_index = (torch.from_numpy(edges).t().contiguous() if len(edges)
                                   else torch.empty((2, 0), dtype=torch.long))
        return data

    def _build_edges(self):
        """The six edge lists [agv->location, location->agv, agv<->agv, picker->location, agv->picker, picker->agv] as
        int64 arrays [E, 2], in the reference's emission order, from the FIRST ``num_agvs + num_pickers`` entries of
        ``_current_agents_info`` and the first ``2 * num_locations`` entries of ``_current_shelves_info``."""
        import numpy as np
        na, npk, nl = self.num_agv_nodes, self.num_picker_nodes, len(self._rack_locations)
        info = self._current_agents_info
        agv = np.asarray(info[:na], dtype=np.float64).reshape(na, -1) if na else np.zeros((0, 7))
        pick = np.asarray(info[na:na + npk], dtype=np.float64).reshape(npk, -1) if npk else np.zeros((0, 4))
        shelves = np.asarray(self._current_shelves_info[:2 * nl], dtype=np.float64).reshape(-1, 2)
        requested = np.flatnonzero((shelves[:, 0] != 0) & (shelves[:, 1] != 0)) if len(shelves) else np.zeros(0, dtype=np.int64)
        if len(shelves) < nl and len(requested):   # the reference slices a short list: missing locations unpack an empty slice
            raise IndexError("list index out of range")
        empty = np.zeros((0, 2), dtype=np.int64)

        # ---- AGV -> location (:1199-1221): the first rack at the target, or every requested shelf when idle
        agv_has_t = ~((agv[:, 5] == 0) & (agv[:, 6] == 0)) if na else np.zeros(0, dtype=bool)
        a2l = []
        if na:
            racks_xy = np.asarray([[r[0], r[1]] for r in self._rack_locations], dtype=np.float64).reshape(-1, 2)
            for a in range(na):
                if agv_has_t[a]:
                    hit = np.flatnonzero((racks_xy[:, 0] == agv[a, 6]) & (racks_xy[:, 1] == agv[a, 5]))
                    if len(hit):
                        a2l.append(np.array([[a, hit[0]]], dtype=np.int64))
                elif len(requested):
                    a2l.append(np.stack([np.full(len(requested), a, dtype=np.int64), requested], axis=1))
        a2l = np.concatenate(a2l, axis=0) if a2l else empty
        l2a = a2l[:, ::-1].copy()

        # ---- AGV <-> AGV (:1223-1248): L1 distance <= max_comm_distance; two targets index the empty section map
        a2a = empty
        if na > 1:
            with_t = np.flatnonzero(agv_has_t)
            if len(with_t) >= 2:
                i = int(with_t[0])
                raise KeyError((np.float64(agv[i, 6]), np.float64(agv[i, 5])))
            iu, ju = np.triu_indices(na, k=1)
            dist = np.abs(agv[iu, 4] - agv[ju, 4]) + np.abs(agv[iu, 3] - agv[ju, 3])
            hit = dist <= self.max_comm_distance
            iu, ju = iu[hit], ju[hit]
            a2a = np.empty((2 * len(iu), 2), dtype=np.int64)
            a2a[0::2, 0], a2a[0::2, 1] = iu, ju
            a2a[1::2, 0], a2a[1::2, 1] = ju, iu

        # ---- picker -> location (:1250-1272)
        p2l = []
        if npk and nl:
            for r in self._rack_locations:
                hash(r)                                     # position_to_sections.get(rack_pos): unhashable entries raise TypeError
            pick_has_t = ~((pick[:, 3] == 0) & (pick[:, 2] == 0))
            for p in range(npk):
                if pick_has_t[p]:
                    raise ValueError("operands could not be broadcast together with shapes (%d,) (2,) " % len(self._rack_locations[0])
                                     if len(self._rack_locations[0]) != 2 else "The truth value of an array with more than one element is ambiguous")
                if len(requested):
                    p2l.append(np.stack([np.full(len(requested), p, dtype=np.int64), requested], axis=1))
        p2l = np.concatenate(p2l, axis=0) if p2l else empty

        # ---- AGV -> picker, picker -> AGV (:1274-1318): close, or the AGV has a target (None == None section match)
        a2p = empty
        if na and npk:
            dist = np.abs(agv[:, None, 4] - pick[None, :, 1]) + np.abs(agv[:, None, 3] - pick[None, :, 0])
            hit = (dist <= self.max_comm_distance) | agv_has_t[:, None]
            ai, pi = np.nonzero(hit)                        # row-major: AGV outer, picker inner, as the reference loops
            a2p = np.stack([ai, pi], axis=1).astype(np.int64)
        p2a = a2p[:, ::-1].copy()
        return [a2l, l2a, a2a, p2l, a2p, p2a]
