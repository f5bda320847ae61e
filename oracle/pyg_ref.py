"""Oracle restatement of the torch_geometric pieces the reference path uses.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``) -- parity unpinned: PyG is
not vendored in ``/root/reference`` nor installed here; this file restates the
published semantics of PyG 2.x ``SAGEConv`` / ``Data`` / ``Batch`` and is
anchored on the reference call sites

  * ``SAGEConv(in, out)(x, edge_index)``   scripts/train_gde.py:27-29,36,39,43
  * ``Data(x=, edge_index=, is_current_agent=)``  scripts/train_gde.py:131,182
  * ``Batch.from_data_list(graphs)``        scripts/train_gde.py:367

Everything runs on CPU in plain PyTorch.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn


# ----------------------------------------------------------------------------
# Containers
# ----------------------------------------------------------------------------
class RefData:
    """Minimal stand-in for ``torch_geometric.data.Data`` (attribute bag)."""

    def __init__(self, x=None, edge_index=None, **kwargs):
        self.x = x
        self.edge_index = edge_index
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def num_nodes(self) -> int:
        return int(self.x.size(0))

    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None]

    def to(self, device):
        for k in self.keys():
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


class RefBatch(RefData):
    """Disjoint union of graphs, as ``Batch.from_data_list`` builds it [upstream PyG].

    * node-level tensors are concatenated along dim 0;
    * ``edge_index`` of graph g is shifted by the number of nodes in graphs < g
      and concatenated along dim 1;
    * ``batch[i]`` = graph id of node i, ``ptr`` = cumulative node offsets.
    """

    @classmethod
    def from_data_list(cls, data_list: List[RefData]) -> "RefBatch":
        xs, eis, bvec, masks = [], [], [], []
        ptr = [0]
        for g, d in enumerate(data_list):
            n = d.num_nodes
            xs.append(d.x)
            eis.append(d.edge_index + ptr[-1])
            bvec.append(torch.full((n,), g, dtype=torch.long))
            if getattr(d, "is_current_agent", None) is not None:
                masks.append(d.is_current_agent)
            ptr.append(ptr[-1] + n)
        out = cls(
            x=torch.cat(xs, dim=0),
            edge_index=torch.cat(eis, dim=1) if eis else torch.empty((2, 0), dtype=torch.long),
        )
        out.batch = torch.cat(bvec, dim=0)
        out.ptr = torch.tensor(ptr, dtype=torch.long)
        if masks:
            out.is_current_agent = torch.cat(masks, dim=0)
        out.num_graphs = len(data_list)
        return out


# ----------------------------------------------------------------------------
# SAGEConv
# ----------------------------------------------------------------------------
def scatter_mean_ref(rows: torch.Tensor, index: torch.Tensor, dim_size: int) -> torch.Tensor:
    """``scatter(rows, index, dim=0, dim_size, reduce='mean')`` [upstream PyG utils].

    Sum by destination, divide by ``count.clamp(min=1)``: an isolated node
    aggregates to exactly 0.  Summation order on CPU = edge order.
    """
    out = rows.new_zeros((dim_size,) + tuple(rows.shape[1:]))
    out.index_add_(0, index, rows)
    cnt = rows.new_zeros(dim_size)
    cnt.index_add_(0, index, rows.new_ones(index.numel()))
    cnt = cnt.clamp(min=1)
    return out / cnt.view(-1, *([1] * (rows.dim() - 1)))


def sage_conv_ref(x: torch.Tensor, edge_index: torch.Tensor, w_l: torch.Tensor,
                  b_l: Optional[torch.Tensor], w_r: torch.Tensor) -> torch.Tensor:
    """SAGEConv forward with PyG defaults (aggr='mean', root_weight=True, bias=True,
    normalize=False, project=False) [upstream]:

        agg_i = mean_{e: dst(e)=i} x[src(e)]          (row 0 of edge_index = src j, row 1 = dst i)
        out   = agg @ w_l.T + b_l + x @ w_r.T
    """
    src, dst = edge_index[0], edge_index[1]
    agg = scatter_mean_ref(x.index_select(0, src), dst, x.size(0))
    out = torch.nn.functional.linear(agg, w_l, b_l)
    out = out + torch.nn.functional.linear(x, w_r)
    return out


class _PygLinear(nn.Module):
    """PyG ``Linear`` : weight [out, in]; kaiming_uniform(a=sqrt(5)) weight, U(+-1/sqrt(in)) bias."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1.0 / math.sqrt(in_channels) if in_channels > 0 else 0.0
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x):
        return torch.nn.functional.linear(x, self.weight, self.bias)


class SAGEConvRef(nn.Module):
    """Parameter layout of PyG ``SAGEConv``: ``lin_l.{weight,bias}``, ``lin_r.weight``.

    ``in_channels`` may be a pair ``(src, dst)`` and ``x`` a pair ``(x_src, x_dst)`` (bipartite message passing, the
    form ``HeteroConv`` uses: scripts/gnode.py:92-97): ``lin_l`` acts on the mean of the SOURCE rows gathered per
    destination, ``lin_r`` on the destination's own row [upstream PyG]."""

    def __init__(self, in_channels, out_channels: int):
        super().__init__()
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.lin_l = _PygLinear(in_channels[0], out_channels, bias=True)
        self.lin_r = _PygLinear(in_channels[1], out_channels, bias=False)

    def forward(self, x, edge_index: torch.Tensor) -> torch.Tensor:
        if isinstance(x, torch.Tensor):
            return sage_conv_ref(x, edge_index, self.lin_l.weight, self.lin_l.bias, self.lin_r.weight)
        x_src, x_dst = x
        src, dst = edge_index[0], edge_index[1]
        agg = scatter_mean_ref(x_src.index_select(0, src), dst, x_dst.size(0))
        out = torch.nn.functional.linear(agg, self.lin_l.weight, self.lin_l.bias)
        return out + torch.nn.functional.linear(x_dst, self.lin_r.weight)


class HeteroConvRef(nn.Module):
    """``torch_geometric.nn.HeteroConv(convs, aggr='mean')`` [upstream]: for every edge type present in BOTH the
    module dict and ``edge_index_dict`` (and whose node types are in ``x_dict``), run its conv on
    ``(x_src, x_dst)`` (or ``x`` when src == dst); outputs of the same destination type are stacked and reduced with
    ``mean``.  Destination types that receive nothing are absent from the result."""

    def __init__(self, convs: Dict[Tuple[str, str, str], nn.Module], aggr: str = "mean"):
        super().__init__()
        assert aggr == "mean"
        self.edge_types = list(convs.keys())
        self.convs = nn.ModuleDict({"__".join(k): v for k, v in convs.items()})

    def forward(self, x_dict, edge_index_dict):
        outs: Dict[str, List[torch.Tensor]] = {}
        for et in self.edge_types:
            src, _rel, dst = et
            if et not in edge_index_dict or src not in x_dict or dst not in x_dict:
                continue
            conv = self.convs["__".join(et)]
            ei = edge_index_dict[et]
            out = conv(x_dict[src], ei) if src == dst else conv((x_dict[src], x_dict[dst]), ei)
            outs.setdefault(dst, []).append(out)
        return {k: torch.stack(v, dim=0).mean(dim=0) for k, v in outs.items()}


class _Store:
    """Attribute bag of one node / edge type of ``RefHeteroData``."""

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


class RefHeteroData:
    """Minimal stand-in for ``torch_geometric.data.HeteroData``: ``data['agv'].x``,
    ``data['agv', 'targets', 'location'].edge_index``, ``data.edge_index_dict``."""

    def __init__(self):
        self._stores: Dict[object, _Store] = {}

    def __getitem__(self, key):
        if isinstance(key, list):
            key = tuple(key)
        if key not in self._stores:
            self._stores[key] = _Store()
        return self._stores[key]

    @property
    def edge_index_dict(self):
        return {k: s.edge_index for k, s in self._stores.items() if isinstance(k, tuple) and "edge_index" in s.__dict__["_d"]}

    @property
    def node_types(self):
        return [k for k in self._stores if isinstance(k, str)]
