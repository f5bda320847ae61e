"""Device-side CSR of a PyG-style ``edge_index`` (built once per batch, cached).

The reference hands an unsorted int64 ``edge_index`` [2, E] to every SAGEConv of every RK stage
(scripts/train_gde.py:74-75 -> :36,39,43).  Here the topology is bucketed once into
destination-sorted and source-sorted int32 CSR (``gnode_csr_build``) and reused by every kernel.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Optional

import torch

from . import _lib


class CSRGraph:
    """int32 CSR (by destination) + transpose CSR (by source) living on one CUDA device."""

    def __init__(self, edge_index: torch.Tensor, num_nodes: int, validate: str = "sync",
                 graph_ptr: Optional[torch.Tensor] = None, max_graph_nodes: Optional[int] = None):
        if not edge_index.is_cuda:
            raise _lib.GnodeError("edge_index must live on a CUDA device; libgnode_b200 has no CPU path")
        if edge_index.dim() != 2 or edge_index.size(0) != 2:
            raise _lib.GnodeError(f"edge_index must be [2, E] (got {list(edge_index.shape)})")
        if edge_index.dtype != torch.int64:
            edge_index = edge_index.to(torch.int64)
        edge_index = edge_index.contiguous()
        dev = edge_index.device
        N, E = int(num_nodes), int(edge_index.size(1))
        self.num_nodes, self.num_edges, self.device = N, E, dev
        self.rowptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
        self.col = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
        self.t_rowptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
        self.t_col = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
        L = _lib.lib()
        nbytes = L.gnode_csr_workspace_bytes(N, E)
        ws = _lib.WORKSPACE.get(nbytes, dev, "csr")
        self._err_dev = torch.empty(1, dtype=torch.int32, device=dev)
        self._err_host = None
        self._err_event = None
        self._device_checks = validate == "device"
        with torch.cuda.device(dev):
            if validate == "sync":
                _lib.check(L.gnode_csr_build(_lib.ptr(edge_index), E, N, _lib.ptr(self.rowptr), _lib.ptr(self.col),
                                             _lib.ptr(self.t_rowptr), _lib.ptr(self.t_col), _lib.ptr(ws), ws.numel(),
                                             _lib.stream_ptr(dev)), "gnode_csr_build")
            elif validate == "device":
                # CUDA-graph capture (graphed.GraphedTrainStep): nothing but kernel launches.  The error flag stays on
                # the device; whoever replays the graph reads it (CSRGraph.device_flags) when it wants a verdict.
                self._err_dev.zero_()
                _lib.check(L.gnode_csr_build_async(_lib.ptr(edge_index), E, N, _lib.ptr(self.rowptr), _lib.ptr(self.col),
                                                   _lib.ptr(self.t_rowptr), _lib.ptr(self.t_col), _lib.ptr(self._err_dev),
                                                   _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)), "gnode_csr_build_async")
            else:
                # no host synchronisation: out-of-range edges are skipped on the device and flagged; the flag travels
                # to pinned host memory asynchronously and is looked at by poll() / validate()
                _lib.check(L.gnode_csr_build_async(_lib.ptr(edge_index), E, N, _lib.ptr(self.rowptr), _lib.ptr(self.col),
                                                   _lib.ptr(self.t_rowptr), _lib.ptr(self.t_col), _lib.ptr(self._err_dev),
                                                   _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)), "gnode_csr_build_async")
                self._err_host = torch.empty(1, dtype=torch.int32, pin_memory=True)
                self._err_host.copy_(self._err_dev, non_blocking=True)
                self._err_event = torch.cuda.Event()
                self._err_event.record(torch.cuda.current_stream(dev))
                _PENDING.append(self)
        # Whole-graph row tiles: let the integrators keep a tile on chip across the stages of a step.  Needs the batch's
        # graph offsets and the host-known size of its largest graph: tiles hold at most 128 rows, or -- for batches with
        # graphs of 129 .. 256 nodes -- 140 / 256 rows, processed as two 128-row blocks by the chain kernels.
        self.tiles = None
        self.tile_err = None
        self.tile_rows = 0
        if (graph_ptr is not None and max_graph_nodes is not None and 0 < int(max_graph_nodes) <= 256
                and graph_ptr.is_cuda and graph_ptr.numel() >= 2):
            m = int(max_graph_nodes)
            self.tile_rows = 128 if m <= 128 else (140 if m <= 140 else 256)
            gp = graph_ptr.to(torch.int64).contiguous()
            self.tiles = torch.empty(gp.numel() + 1, dtype=torch.int32, device=dev)
            self.tile_err = torch.zeros(1, dtype=torch.int32, device=dev)
            if validate == "device":
                self.tile_err.zero_()       # part of the captured graph: every replay starts from a clean flag
            with torch.cuda.device(dev):
                _lib.check(L.gnode_tiles_build_rows(_lib.ptr(gp), gp.numel() - 1, self.tile_rows, _lib.ptr(self.tiles),
                                                    _lib.stream_ptr(dev)), "gnode_tiles_build_rows")
        self.struct = _lib.GnodeGraph(N, E, self.rowptr.data_ptr(), self.col.data_ptr(), self.t_rowptr.data_ptr(),
                                      self.t_col.data_ptr(), _lib.ptr(self.tiles), _lib.ptr(self.tile_err), self.tile_rows)

    def ref(self):
        return C.byref(self.struct)

    def poll(self, wait: bool = False) -> bool:
        """Deferred index validation: returns True once the verdict is known; raises GnodeError for a bad edge list."""
        if self._err_event is None:
            return True
        if wait:
            self._err_event.synchronize()
        elif not self._err_event.query():
            return False
        bad = int(self._err_host[0]) != 0
        self._err_event = None
        if bad:
            raise _lib.GnodeError(f"gnode_csr_build: edge_index holds node ids outside [0, {self.num_nodes}) "
                                  "(deferred check; the offending edges were skipped)")
        return True

    def validate(self) -> None:
        self.poll(wait=True)
        # this call gives the verdict for the graph: drop its deferred checks so that they do not raise again later
        _PENDING[:] = [c for c in _PENDING if not (isinstance(c, _TileCheck) and c.graph is self)]
        if self.tile_err is not None:
            _raise_tile_error(int(self.tile_err.item()), deferred=False)
        _lib.tc_check(self.device)

    def device_flags(self) -> torch.Tensor:
        """[csr index error, tile error] as a device tensor (no synchronisation): for graphs built with
        ``validate="device"``, whose checks are not scheduled automatically."""
        te = self.tile_err if self.tile_err is not None else torch.zeros_like(self._err_dev)
        return torch.cat([self._err_dev, te])

    def schedule_tile_check(self) -> None:
        """Called after the graph-resident kernels ran: ship the tile error flag and the tcgen05 barrier status word to
        pinned memory without blocking; a later poll_pending() raises if an edge was found outside its tile, a graph
        exceeded the announced tile capacity, or a barrier wait expired.  Completed checks are looked at here (loops
        that reuse one batch never reach csr_for's cache-miss poll), and a graph keeps at most ONE outstanding check."""
        if self._device_checks:
            return                      # captured graph: the flags are read by the replaying side (device_flags)
        poll_pending()
        if self.tile_err is None:
            return
        _PENDING[:] = [c for c in _PENDING if not (isinstance(c, _TileCheck) and c.graph is self)]
        _PENDING.append(_TileCheck(self))


def _raise_tile_error(code: int, deferred: bool) -> None:
    tail = " (deferred check; results of that call are invalid)" if deferred else ""
    if code == 2:
        raise _lib.GnodeError("a graph of the batch has more nodes than Batch.max_graph_nodes announced: the whole-graph "
                              "tiles could not be built and the integrator produced nothing" + tail)
    if code != 0:
        raise _lib.GnodeError("an edge leaves its whole-graph tile: batch.ptr does not describe a disjoint union of "
                              "the graphs in edge_index" + tail)


class _TileCheck:
    def __init__(self, g: "CSRGraph"):
        self.graph = g
        dev = g.device
        self.host = torch.zeros(2, dtype=torch.int32, pin_memory=True)     # [tile error flag, tcgen05 status word]
        stream = torch.cuda.current_stream(dev)
        self.host[:1].copy_(g.tile_err, non_blocking=True)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().gnode_tc_status_async(self.host.data_ptr() + 4, stream.cuda_stream), "gnode_tc_status_async")
        self.event = torch.cuda.Event()
        self.event.record(stream)

    def poll(self) -> bool:
        if not self.event.query():
            return False
        _raise_tile_error(int(self.host[0]), deferred=True)
        if int(self.host[1]) != 0:
            with torch.cuda.device(self.graph.device):      # read-and-clear, so that the next call starts healthy
                _lib.lib().gnode_tc_status(_lib.stream_ptr(self.graph.device))
            raise _lib.GnodeError(f"a tcgen05 kernel gave up on a barrier wait (code {int(self.host[1])}): the results of "
                                  "that call are invalid (deferred check)")
        return True


_PENDING: list = []


def poll_pending() -> None:
    """Look (without blocking) at the deferred validations that have completed; raises the first failure found (every
    completed check leaves the list either way)."""
    keep, first = [], None
    for g in list(_PENDING):
        try:
            if not g.poll():
                keep.append(g)
        except _lib.GnodeError as e:
            first = first or e
    _PENDING[:] = keep
    if first is not None:
        raise first


_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()  # key -> (keyed edge_index tensor, CSRGraph)
_CACHE_MAX = 8


def csr_for(edge_index: torch.Tensor, num_nodes: int, holder: Optional[object] = None,
            validate: str = "sync", graph_ptr: Optional[torch.Tensor] = None,
            max_graph_nodes: Optional[int] = None) -> CSRGraph:
    """CSR of ``edge_index``; cached on ``holder`` (e.g. the batch object) and in a small LRU keyed by
    the tensor's storage, shape and version counter.  Every cache entry keeps the keyed tensor alive,
    so its address cannot be recycled for a different edge list while the entry exists."""
    # the tiling is part of the graph: the same edge list with another `ptr` (or another announced capacity) is a
    # different entry.  A caller that passes NO offsets accepts whatever entry exists for the edge list (a tiled graph
    # serves every un-tiled use).
    base = (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version, int(num_nodes), str(edge_index.device))
    key = base + (graph_ptr.data_ptr() if graph_ptr is not None else 0, int(max_graph_nodes) if max_graph_nodes is not None else 0)

    def matches(k):
        return k == key or (graph_ptr is None and k[:5] == base)

    if holder is not None:
        cached = getattr(holder, "__dict__", {}).get("_gnode_csr")
        if cached is not None and matches(cached[0]):
            return cached[1]
    entry = _CACHE.get(key)
    if entry is None and graph_ptr is None:
        for k in reversed(_CACHE):
            if k[:5] == base:
                key, entry = k, _CACHE[k]
                break
    if entry is None or entry[0] is not edge_index and entry[0].data_ptr() != edge_index.data_ptr():
        if validate == "deferred":
            poll_pending()
        entry = (edge_index, CSRGraph(edge_index, num_nodes, validate=validate, graph_ptr=graph_ptr,
                                      max_graph_nodes=max_graph_nodes))
        _CACHE[key] = entry
        while len(_CACHE) > _CACHE_MAX:
            _CACHE.popitem(last=False)
    else:
        _CACHE.move_to_end(key)
    g = entry[1]
    if holder is not None:
        try:
            holder.__dict__["_gnode_csr"] = (key, g, edge_index)
        except Exception:
            pass
    return g


def clear_cache():
    _CACHE.clear()
