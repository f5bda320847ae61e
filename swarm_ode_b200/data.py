"""Graph containers and graph construction for the GNODE path.

Host-side mirror of the reference's input contract:

* ``Data`` / ``Batch`` -- the attributes of ``torch_geometric.data.Data`` / ``Batch`` that the
  reference path touches (``x``, ``edge_index``, ``batch``, ``ptr``, ``is_current_agent``, ``.to()``,
  ``Batch.from_data_list``; scripts/train_gde.py:69-71,131,182,367,475).  A real PyG ``Batch`` is
  accepted everywhere these duck types are.
* ``GraphConverter`` -- same constructor / methods / outputs as scripts/train_gde.py:108-271, with the
  O(n^2) Python pair loop replaced by vectorised float32 numpy (bit-exact: same expression
  ``sqrt(sum((p_i - p_j)**2)) < thr`` in the observation dtype, same emission order).
* ``spatial_edges_cuda`` -- the same edges for many snapshots at once on the GPU
  (``gnode_spatial_edges``), bit-exact.
* ``TrajectoryBatch`` / ``collate_trajectory_batches`` -- scripts/train_gde.py:273-276,363-375.
"""
from __future__ import annotations

from collections import deque
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib


class Data:
    """Attribute bag with the ``Data`` surface the reference uses."""

    def __init__(self, x: Optional[torch.Tensor] = None, edge_index: Optional[torch.Tensor] = None, **kwargs):
        self.x = x
        self.edge_index = edge_index
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def num_nodes(self) -> int:
        return int(self.x.size(0))

    @property
    def num_edges(self) -> int:
        return int(self.edge_index.size(1))

    def keys(self):
        return [k for k, v in self.__dict__.items() if not k.startswith("_") and v is not None]

    def to(self, device, non_blocking: bool = False):
        for k in self.keys():
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(self, k, v.to(device, non_blocking=non_blocking))
        # cached device-side CSR (see graph_cache) does not survive a move
        self.__dict__.pop("_gnode_csr", None)
        return self

    def pin_memory(self):
        for k in self.keys():
            v = getattr(self, k)
            if torch.is_tensor(v) and not v.is_cuda:
                setattr(self, k, v.pin_memory())
        return self

    def __repr__(self):
        parts = [f"{k}={list(getattr(self, k).shape)}" if torch.is_tensor(getattr(self, k)) else f"{k}={getattr(self, k)}"
                 for k in self.keys()]
        return f"{type(self).__name__}({', '.join(parts)})"


class Batch(Data):
    """Disjoint union of graphs (``Batch.from_data_list`` semantics [upstream PyG])."""

    @classmethod
    def from_data_list(cls, data_list: Sequence[Data]) -> "Batch":
        sizes = [int(d.x.size(0)) for d in data_list]
        ptr = torch.zeros(len(sizes) + 1, dtype=torch.long)
        ptr[1:] = torch.cumsum(torch.tensor(sizes, dtype=torch.long), 0)
        out = cls(x=torch.cat([d.x for d in data_list], dim=0))
        eis = [d.edge_index + int(off) for d, off in zip(data_list, ptr[:-1])]
        out.edge_index = torch.cat(eis, dim=1) if eis else torch.empty((2, 0), dtype=torch.long)
        out.batch = torch.repeat_interleave(torch.arange(len(sizes), dtype=torch.long),
                                            torch.tensor(sizes, dtype=torch.long))
        out.ptr = ptr
        masks = [getattr(d, "is_current_agent", None) for d in data_list]
        if all(m is not None for m in masks) and masks:
            out.is_current_agent = torch.cat(masks, dim=0)
        out.num_graphs = len(sizes)
        out.max_graph_nodes = max(sizes) if sizes else 0     # host-side fact: enables the graph-resident kernels
        return out

    def shard(self, rank: int, world_size: int) -> "Batch":
        """Contiguous range of whole graphs for one data-parallel rank (graphs never span ranks)."""
        G = int(self.ptr.numel() - 1)
        lo, hi = (G * rank) // world_size, (G * (rank + 1)) // world_size
        n0, n1 = int(self.ptr[lo]), int(self.ptr[hi])
        ei = self.edge_index
        sel = (ei[1] >= n0) & (ei[1] < n1)
        out = Batch(x=self.x[n0:n1], edge_index=ei[:, sel] - n0)
        out.batch = self.batch[n0:n1] - lo
        out.ptr = self.ptr[lo:hi + 1] - n0
        if getattr(self, "is_current_agent", None) is not None:
            out.is_current_agent = self.is_current_agent[n0:n1]
        out.num_graphs = hi - lo
        if getattr(self, "max_graph_nodes", None) is not None:
            out.max_graph_nodes = self.max_graph_nodes
        return out


# ----------------------------------------------------------------------------------------------
# Lossless narrow transport of a collated batch (host -> device)
# ----------------------------------------------------------------------------------------------
_PACK_TYPES = ((torch.uint8, _lib.PACK_U8, "u8"), (torch.int16, _lib.PACK_I16, "i16"), (torch.float16, _lib.PACK_F16, "f16"))


def narrowest_exact_dtype(x: torch.Tensor):
    """``(torch dtype, GNODE_PACK_* code, name)`` of the narrowest type that reproduces every element of the fp32 tensor
    ``x`` exactly (u8, i16, f16, else f32).  The check is the round trip itself, element by element."""
    if x.dtype != torch.float32:
        raise _lib.GnodeError(f"node features must be float32 (got {x.dtype})")
    finite = bool(torch.isfinite(x).all()) if x.numel() else True
    if finite:
        lo, hi = (float(x.min()), float(x.max())) if x.numel() else (0.0, 0.0)
        for dt, code, name in _PACK_TYPES:
            if dt == torch.uint8 and (lo < 0.0 or hi > 255.0):
                continue
            if dt == torch.int16 and (lo < -32768.0 or hi > 32767.0):
                continue
            if dt == torch.float16 and max(abs(lo), abs(hi)) > 65504.0:
                continue
            if torch.equal(x.to(dt).to(torch.float32), x) and not (dt != torch.float16 and bool(((x == 0) & torch.signbit(x)).any())):
                return dt, code, name
    return torch.float32, _lib.PACK_F32, "f32"


BITS_MAX_COLS, BITS_MAX_ROW_BYTES = 2048, 1024      # limits of gnode_unpack_bits


def pack_bits(x: torch.Tensor):
    """Column-wise bit packing of an fp32 matrix whose entries are integers in [0, 255] (warehouse observations: flags and
    grid coordinates).  Column ``c`` gets ``w_c = bit_length(max of the column)`` bits; a row is the little-endian bit string
    of its columns, padded to a multiple of 16 bytes with at least one spare byte.  Returns ``(packed uint8 [rows,
    row_bytes], bit_offsets int32 [cols + 1])`` or ``None`` when the matrix is not representable (negative, fractional,
    > 255, -0.0, non-finite values; too many columns) -- the caller then falls back to a wider transport type.
    Exactness is by construction (every value is an integer below 2^w_c) and is re-checked by ``unpack_bits_host``."""
    if x.dim() != 2 or x.dtype != torch.float32 or x.shape[1] < 1 or x.shape[1] > BITS_MAX_COLS:
        return None
    rows, cols = x.shape
    if rows == 0:
        return None
    if not bool(torch.isfinite(x).all()) or float(x.min()) < 0.0 or float(x.max()) > 255.0:
        return None
    xi = x.to(torch.uint8)
    if not torch.equal(xi.to(torch.float32), x) or bool(((x == 0) & torch.signbit(x)).any()):
        return None
    xt = np.ascontiguousarray(xi.numpy().T)                  # [cols, rows]: every column contiguous
    width = np.array([int(v).bit_length() for v in xt.max(axis=1)], dtype=np.int64)
    off = np.zeros(cols + 1, dtype=np.int64)
    off[1:] = np.cumsum(width)
    row_bytes = int(((off[-1] + 7) // 8 + 1 + 15) // 16 * 16)
    if row_bytes > BITS_MAX_ROW_BYTES:
        return None
    out_t = np.zeros((row_bytes, rows), dtype=np.uint8)      # byte-major while it is built, transposed once at the end
    for c in range(cols):
        if width[c] == 0:
            continue
        b, sh = int(off[c] >> 3), int(off[c] & 7)
        v = xt[c].astype(np.uint16) << sh
        out_t[b] |= v.astype(np.uint8)                        # low byte (astype truncates)
        if sh + width[c] > 8:
            out_t[b + 1] |= (v >> 8).astype(np.uint8)
    return torch.from_numpy(np.ascontiguousarray(out_t.T)), torch.from_numpy(off.astype(np.int32))


def unpack_bits_host(packed: torch.Tensor, bit_offsets: torch.Tensor) -> torch.Tensor:
    """Host inverse of ``pack_bits`` (the arithmetic of ``gnode_unpack_bits``): used to verify a packed batch."""
    pt = np.ascontiguousarray(packed.numpy().T)              # [row_bytes, rows]
    off = bit_offsets.numpy().astype(np.int64)
    cols = off.shape[0] - 1
    out_t = np.zeros((cols, pt.shape[1]), dtype=np.uint8)
    for c in range(cols):
        w = int(off[c + 1] - off[c])
        if w == 0:
            continue
        b, sh = int(off[c] >> 3), int(off[c] & 7)
        v = pt[b].astype(np.uint16) | (pt[b + 1].astype(np.uint16) << 8)
        out_t[c] = ((v >> sh) & ((1 << w) - 1)).astype(np.uint8)
    return torch.from_numpy(np.ascontiguousarray(out_t.T)).to(torch.float32)


class PackedBatch:
    """A collated batch in its transport format: what ``batch.to(device)`` of the reference (scripts/train_gde.py:475)
    moves over PCIe, made as small as it can be WITHOUT changing a single bit of what arrives.

    * ``x`` travels in the narrowest form that reproduces every fp32 value exactly -- column-wise bit packing for small
      non-negative integers (warehouse observations are flags and grid coordinates: 105 bytes per 399-column row), else
      u8 / i16 / f16 / f32; the choice is verified element by element when the batch is packed, i.e. at dataset-build
      time -- and is widened to fp32 on the device (``gnode_unpack_bits`` / ``gnode_unpack_features``);
    * ``edge_index`` travels as int32 (node ids < 2^31) and is widened to int64 (``gnode_unpack_edges``);
    * ``batch`` is not transported at all: it is a function of ``ptr`` (``gnode_batch_vector``);
    * ``ptr``, ``is_current_agent`` and the targets travel as they are.

    ``to(device)`` returns an ordinary :class:`Batch` whose tensors are bit-identical to ``batch.to(device)``."""

    def __init__(self, batch: "Batch", next_positions: Optional[torch.Tensor] = None, bits: bool = True):
        if batch.x.is_cuda:
            raise _lib.GnodeError("PackedBatch packs a HOST batch")
        x = batch.x.contiguous()
        dt, self.kind, self.kind_name = narrowest_exact_dtype(x)
        self.x_shape = tuple(x.shape)
        self.x_packed = x if dt == torch.float32 else x.to(dt)
        self.bit_offsets = None
        if bits and self.kind == _lib.PACK_U8:
            # small non-negative integers: column-wise bit widths (flags 1 bit, coordinates 5 bits) beat one byte per value
            pb = pack_bits(x)
            if pb is not None and pb[0].shape[1] < x.shape[1]:
                if not torch.equal(unpack_bits_host(*pb), x):          # the round trip itself, element by element
                    raise _lib.GnodeError("pack_bits: round trip is not exact")
                self.x_packed, self.bit_offsets = pb
                self.kind, self.kind_name = _lib.PACK_BITS, "bits"
        ei = batch.edge_index
        self.num_nodes = int(x.shape[0])
        self.edges_int32 = self.num_nodes < 2 ** 31 and (ei.numel() == 0 or (int(ei.min()) >= 0 and int(ei.max()) < 2 ** 31))
        self.edge_index = ei.to(torch.int32).contiguous() if self.edges_int32 else ei.contiguous()
        self.ptr = getattr(batch, "ptr", None)
        self.batch = None if self.ptr is not None else batch.batch      # without offsets the vector itself must travel
        self.is_current_agent = getattr(batch, "is_current_agent", None)
        self.num_graphs = getattr(batch, "num_graphs", None)
        self.max_graph_nodes = getattr(batch, "max_graph_nodes", None)
        self.next_positions = next_positions

    def _tensors(self):
        return [t for t in (self.x_packed, self.bit_offsets, self.edge_index, self.ptr, self.batch, self.is_current_agent,
                            self.next_positions) if t is not None]

    @property
    def nbytes(self) -> int:
        """Bytes one ``to(device)`` moves over the link."""
        return sum(t.numel() * t.element_size() for t in self._tensors())

    def pin_memory(self) -> "PackedBatch":
        for k in ("x_packed", "bit_offsets", "edge_index", "ptr", "batch", "is_current_agent", "next_positions"):
            v = getattr(self, k)
            if v is not None and not v.is_pinned():
                setattr(self, k, v.pin_memory())
        return self

    def to(self, device, non_blocking: bool = False):
        """Upload and widen on ``device``'s current stream.  Returns ``batch`` or ``(batch, next_positions)``."""
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.GnodeError("PackedBatch.to() targets a CUDA device; libgnode_b200 has no CPU path")
        L = _lib.lib()
        with torch.cuda.device(device):
            s = _lib.stream_ptr(device)
            xp = self.x_packed.to(device, non_blocking=non_blocking)
            bo = None
            if self.kind == _lib.PACK_F32:
                x = xp
            elif self.kind == _lib.PACK_BITS:
                bo = self.bit_offsets.to(device, non_blocking=non_blocking)
                x = torch.empty(self.x_shape, dtype=torch.float32, device=device)
                _lib.check(L.gnode_unpack_bits(_lib.ptr(xp), self.x_shape[0], self.x_shape[1], int(xp.shape[1]), _lib.ptr(bo),
                                               _lib.ptr(x), s), "gnode_unpack_bits")
            else:
                x = torch.empty(self.x_shape, dtype=torch.float32, device=device)
                _lib.check(L.gnode_unpack_features(_lib.ptr(xp), self.kind, x.numel(), _lib.ptr(x), s), "gnode_unpack_features")
            ep = self.edge_index.to(device, non_blocking=non_blocking)
            if self.edges_int32:
                ei = torch.empty(tuple(ep.shape), dtype=torch.int64, device=device)
                _lib.check(L.gnode_unpack_edges(_lib.ptr(ep), ep.numel(), _lib.ptr(ei), s), "gnode_unpack_edges")
            else:
                ei = ep
            out = Batch(x=x, edge_index=ei)
            if self.ptr is not None:
                out.ptr = self.ptr.to(device, non_blocking=non_blocking)
                out.batch = torch.empty(self.num_nodes, dtype=torch.int64, device=device)
                _lib.check(L.gnode_batch_vector(_lib.ptr(out.ptr), out.ptr.numel() - 1, self.num_nodes, _lib.ptr(out.batch), s),
                           "gnode_batch_vector")
            else:
                out.batch = self.batch.to(device, non_blocking=non_blocking)
            if self.is_current_agent is not None:
                out.is_current_agent = self.is_current_agent.to(device, non_blocking=non_blocking)
            if self.num_graphs is not None:
                out.num_graphs = self.num_graphs
            if self.max_graph_nodes is not None:
                out.max_graph_nodes = self.max_graph_nodes
            # the packed device copies are read by kernels enqueued on this stream: keep the allocator from recycling them
            # under a different stream before those kernels ran
            for t in (xp, ep, bo):
                if t is not None:
                    t.record_stream(torch.cuda.current_stream(device))
        if self.next_positions is not None:
            return out, self.next_positions.to(device, non_blocking=non_blocking)
        return out


# ----------------------------------------------------------------------------------------------
# GraphConverter
# ----------------------------------------------------------------------------------------------
def _pair_edges(loc: np.ndarray, threshold: float) -> np.ndarray:
    """[2, E] int64 spatial edges of one snapshot, order (i,j),(j,i) for i<j lexicographic."""
    n = loc.shape[0]
    if n < 2:
        return np.empty((2, 0), dtype=np.int64)
    iu, ju = np.triu_indices(n, k=1)
    diff = loc[iu] - loc[ju]
    # same arithmetic as np.sqrt(np.sum((a - b) ** 2)) on a length-2 vector in loc.dtype
    d = np.sqrt((diff[:, 0] ** 2) + (diff[:, 1] ** 2))
    hit = d < threshold  # python-float threshold is a weak scalar: compared in loc.dtype, as in the reference
    iu, ju = iu[hit], ju[hit]
    e = np.empty((2, 2 * iu.size), dtype=np.int64)
    e[0, 0::2], e[1, 0::2] = iu, ju
    e[0, 1::2], e[1, 1::2] = ju, iu
    return e


class GraphConverter:
    """Drop-in for scripts/train_gde.py:108-271 (same ctor, methods and output ``Data``)."""

    def __init__(self, num_agvs, num_pickers, distance_threshold: float = 3.0, temporal_window: int = 5):
        self.num_agvs = num_agvs
        self.num_pickers = num_pickers
        self.distance_threshold = distance_threshold
        self.temporal_window = temporal_window
        self.graph_history = deque(maxlen=temporal_window)

    def reset_history(self):
        self.graph_history.clear()

    def _standardize_observations(self, observations) -> np.ndarray:
        if isinstance(observations, np.ndarray) and observations.dtype == object:
            rows = observations.tolist()
        elif isinstance(observations, list):
            rows = observations
        else:
            return observations
        width = max(len(r) for r in rows)
        out = np.zeros((len(rows), width), dtype=np.float32)
        for i, r in enumerate(rows):
            r = np.array(r, dtype=np.float32)
            out[i, :len(r)] = r
        return out

    def _extract_locations_by_agent_type(self, observations: np.ndarray) -> np.ndarray:
        n = len(observations)
        agv = np.arange(n) < self.num_agvs
        loc = np.empty((n, 2), dtype=observations.dtype)
        loc[agv] = observations[agv][:, 3:5]
        loc[~agv] = observations[~agv][:, 0:2]
        return loc

    def _compute_spatial_edges(self, locations: np.ndarray) -> torch.Tensor:
        return torch.from_numpy(_pair_edges(np.asarray(locations), self.distance_threshold))

    def _compute_temporal_edges_with_window(self, num_agents: int) -> torch.Tensor:
        k = len(self.graph_history) - 1
        if k <= 0:
            return torch.empty((2, 0), dtype=torch.long)
        a = torch.arange(num_agents, dtype=torch.long)
        return torch.stack([(k - 1) * num_agents + a, k * num_agents + a], dim=0)

    def _build_graph_from_observation(self, observations) -> Data:
        obs = self._standardize_observations(observations)
        n = len(obs)
        x_t = torch.tensor(obs, dtype=torch.float32)
        e_t = self._compute_spatial_edges(self._extract_locations_by_agent_type(obs))
        self.graph_history.append(Data(x=x_t, edge_index=e_t))
        k = len(self.graph_history) - 1
        x = torch.cat([g.x for g in self.graph_history], dim=0)
        parts = [self.graph_history[i].edge_index + i * n for i in range(k)]
        if e_t.shape[1] > 0:
            parts.append(e_t + k * n)
        t_e = self._compute_temporal_edges_with_window(n)
        if t_e.shape[1] > 0:
            parts.append(t_e)
        edge_index = torch.cat(parts, dim=1) if parts else torch.empty((2, 0), dtype=torch.long)
        mask = torch.zeros(x.size(0), dtype=torch.bool)
        mask[k * n:(k + 1) * n] = True
        return Data(x=x, edge_index=edge_index, is_current_agent=mask)


def spatial_edges_cuda(pos: torch.Tensor, threshold: float):
    """Spatial edges of many snapshots at once on the GPU (bit-exact with ``GraphConverter``).

    pos: CUDA float32 [n_snap, n_agents, 2] of (y, x).  Returns ``(counts int32 [n_snap],
    edges int32 [n_snap, n_agents*(n_agents-1), 2])``; only the first ``counts[s]`` rows of
    ``edges[s]`` are valid, each row = (src, dst), in the reference's emission order.
    """
    pos = _lib.require_cuda_f32(pos, "pos")
    n_snap, n, two = pos.shape
    assert two == 2
    counts = torch.empty(n_snap, dtype=torch.int32, device=pos.device)
    edges = torch.empty((n_snap, max(n * (n - 1), 1), 2), dtype=torch.int32, device=pos.device)
    _lib.check(_lib.lib().gnode_spatial_edges(_lib.ptr(pos), n_snap, n, float(threshold), _lib.ptr(counts),
                                              _lib.ptr(edges), _lib.stream_ptr(pos.device)), "gnode_spatial_edges")
    return counts, edges


def build_episode_batch(observations: torch.Tensor, num_agvs: int, num_pickers: int, distance_threshold: float = 3.0,
                        temporal_window: int = 5) -> "Batch":
    """All window graphs of ONE episode, built on the GPU, as the ``Batch`` the reference gets from a fresh
    ``GraphConverter`` fed step by step (scripts/train_gde.py:308-314) and ``Batch.from_data_list`` (:367) -- bit for bit.

    observations: CUDA float32 ``[n_steps, num_agvs + num_pickers, D]`` (zero-padded rows, as collect_data.py writes them).
    The O(n^2) Python pair loop per step and the per-batch host collation disappear; one small device->host read (the
    total edge count) sizes the returned ``edge_index``."""
    obs = _lib.require_cuda_f32(observations, "observations")
    if obs.dim() != 3 or obs.shape[1] != num_agvs + num_pickers:
        raise _lib.GnodeError(f"observations must be [n_steps, {num_agvs + num_pickers}, D] (got {list(obs.shape)})")
    T, n, D = obs.shape
    dev = obs.device
    L = _lib.lib()
    N = int(L.gnode_window_graphs_nodes(T, n, temporal_window))
    cap = int(L.gnode_window_graphs_edge_capacity(T, n, temporal_window))
    x = torch.empty((N, D), dtype=torch.float32, device=dev)
    ei = torch.empty((2, cap), dtype=torch.int64, device=dev)
    batch = torch.empty(N, dtype=torch.int64, device=dev)
    cur = torch.empty(N, dtype=torch.bool, device=dev)
    ptr = torch.empty(T + 1, dtype=torch.int64, device=dev)
    eoff = torch.empty(T + 1, dtype=torch.int64, device=dev)
    ws = _lib.WORKSPACE.get(L.gnode_window_graphs_workspace_bytes(T, n), dev)
    with torch.cuda.device(dev):
        _lib.check(L.gnode_window_graphs(_lib.ptr(obs), T, n, D, int(num_agvs), float(distance_threshold), int(temporal_window),
                                         _lib.ptr(x), _lib.ptr(ei), cap, _lib.ptr(batch), _lib.ptr(cur), _lib.ptr(ptr),
                                         _lib.ptr(eoff), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)), "gnode_window_graphs")
    E = int(eoff[-1])
    out = Batch(x=x, edge_index=ei[:, :E].contiguous(), is_current_agent=cur)
    out.batch, out.ptr = batch, ptr
    out.num_graphs = T
    out.max_graph_nodes = min(T, temporal_window) * n
    out.edge_ptr = eoff
    return out


# ----------------------------------------------------------------------------------------------
# Batching (scripts/train_gde.py:273-276, 336-375)
# ----------------------------------------------------------------------------------------------
class TrajectoryBatch:
    def __init__(self, graphs, next_positions: torch.Tensor):
        self.graphs = graphs
        self.next_positions = next_positions


def extract_positions_from_graph(graph: Data, num_agvs: int, num_pickers: int) -> torch.Tensor:
    """scripts/train_gde.py:336-355 (reads the first n rows of the window graph -- reference quirk kept)."""
    parts = []
    if num_agvs > 0:
        parts.append(graph.x[:num_agvs][:, [4, 3]])
    if num_pickers > 0:
        parts.append(graph.x[num_agvs:num_agvs + num_pickers][:, [1, 0]])
    return torch.cat(parts, dim=0)


def collate_trajectory_batches(batch_list: List[TrajectoryBatch]) -> TrajectoryBatch:
    graphs = Batch.from_data_list([b.graphs for b in batch_list])
    nxt = torch.stack([b.next_positions for b in batch_list], dim=0)
    return TrajectoryBatch(graphs=graphs, next_positions=nxt)
