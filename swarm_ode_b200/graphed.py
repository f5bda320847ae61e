"""The training step of scripts/train_gde.py:478-495 as ONE CUDA-graph launch per batch.

At the reference's own batch size (``batch_size: 32``, scripts/train_gde.py:437-445) a step is a few thousand nodes: the
GPU work is tens of microseconds while the eager step issues ~60 library launches plus the Python autograd / optimizer
glue, so the step is bound by the host (1.47 ms at 32 graphs).  ``GraphedTrainStep`` captures the whole step -- CSR
build, tiling, forward through the solver, masked MSE, backward through the solver, gradient clipping, Adam -- once into
a CUDA graph over static input buffers and replays it per batch: one launch, no Python in the loop.

What makes the step capturable: the fixed-grid integrators never synchronise, every workspace is caller-owned and
grow-only (allocated during the warm-up steps), index / tile validation stays on the device in capture mode
(``CSRGraph(validate="device")``) and is read back by ``check()``.  Batches differ in their number of edges, a graph has
fixed shapes: the edge buffer has a fixed capacity and is padded with the (-1, -1) padding edge, which the CSR build
skips.  The number of nodes, graphs and masked nodes per batch must stay what it was at capture (true for the
reference's loader: every window graph has ``n_agents * window`` nodes); ``step`` falls back to the eager step
otherwise.  The graph holds the ADDRESSES of the parameters and of the optimizer's state tensors: change them in place
(``load_state_dict`` on the model, ``tensor.copy_`` / ``zero_`` on optimizer state), never by rebinding.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import graph as G
from ._lib import GnodeError
from .data import Batch
from .dist import masked_mse_train_step


class GraphedTrainStep:
    def __init__(self, model, optimizer, example: Batch, example_next: torch.Tensor,
                 time_span: Optional[torch.Tensor] = None, max_norm: float = 1.0,
                 edge_capacity: Optional[int] = None, warmup: int = 3, preserve_state: bool = True):
        # The captured step is LOCAL to this process: no gradient all-reduce is captured.  Under torch.distributed use it only
        # for per-rank work (e.g. rank-0 evaluation / side benches); data-parallel training uses dist.masked_mse_train_step.
        if not example.x.is_cuda:
            raise GnodeError("GraphedTrainStep needs a CUDA batch")
        if model.ode_solver not in ("euler", "midpoint", "rk4"):
            raise GnodeError("GraphedTrainStep needs a fixed-grid solver (dopri5's step control runs on the host)")
        for g in optimizer.param_groups:
            if "capturable" in g and not g["capturable"]:
                raise GnodeError("the optimizer must be built with capturable=True to be captured in a CUDA graph")
        dev = example.x.device
        self.model, self.optimizer, self.max_norm, self.device = model, optimizer, max_norm, dev
        self.time_span = time_span if time_span is not None else torch.tensor([0.0, 1.0], device=dev)
        E = int(example.edge_index.size(1))
        self.edge_capacity = int(edge_capacity) if edge_capacity is not None else E + E // 2 + 64
        if self.edge_capacity < E:
            raise GnodeError("edge_capacity is smaller than the example batch's edge count")
        # static inputs of the captured step
        sb = Batch(x=torch.empty_like(example.x),
                   edge_index=torch.full((2, self.edge_capacity), -1, dtype=torch.int64, device=dev))
        sb.batch = example.batch.clone()
        sb.is_current_agent = example.is_current_agent.clone()
        if getattr(example, "ptr", None) is not None:
            sb.ptr = example.ptr.clone()
            sb.max_graph_nodes = getattr(example, "max_graph_nodes", None)
        sb.num_graphs = getattr(example, "num_graphs", None)
        self.static = sb
        self.static_next = torch.empty_like(example_next)
        self._shape = (tuple(example.x.shape), int(example_next.numel()), int(example.batch.numel()))
        self._load(example, example_next)
        # The warm-up steps below are real optimizer steps: with preserve_state the weights and the optimizer state are put
        # back afterwards, IN PLACE (the graph holds their addresses), so that training continues from where it was.
        params = [p for g in optimizer.param_groups for p in g["params"]]
        snap_p = [p.detach().clone() for p in params] if preserve_state else None
        snap_s = ({id(p): {k: v.detach().clone() for k, v in optimizer.state[p].items() if torch.is_tensor(v)}
                   for p in params if p in optimizer.state} if preserve_state else None)
        # warm-up on a side stream: grows every workspace, runs every first-use attribute call, fills the time-grid cache
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._eager_on_static()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        self._drop_csr()
        with torch.cuda.graph(self.graph):
            self.loss = self._eager_on_static()
        self._csr = self.static.__dict__.get("_gnode_csr")
        self.replays = 0
        if preserve_state:
            with torch.no_grad():
                for p, p0 in zip(params, snap_p):
                    p.copy_(p0)
                for p in params:
                    for k, v in optimizer.state.get(p, {}).items():
                        if torch.is_tensor(v):
                            old = snap_s.get(id(p), {}).get(k)
                            v.copy_(old) if old is not None else v.zero_()     # a state born in the warm-up starts from zero
            optimizer.zero_grad(set_to_none=False)

    # -- helpers --------------------------------------------------------------------------------------------------
    def _drop_csr(self):
        G.clear_cache()
        self.static.__dict__.pop("_gnode_csr", None)

    def _eager_on_static(self):
        self._drop_csr()
        return masked_mse_train_step(self.model, self.optimizer, self.static, self.static_next, self.time_span, self.max_norm,
                                     distributed=False)

    def _load(self, batch: Batch, nxt: torch.Tensor):
        E = int(batch.edge_index.size(1))
        self.static.x.copy_(batch.x, non_blocking=True)
        self.static.edge_index[:, :E].copy_(batch.edge_index, non_blocking=True)
        if E < self.edge_capacity:
            self.static.edge_index[:, E:].fill_(-1)
        self.static.batch.copy_(batch.batch, non_blocking=True)
        self.static.is_current_agent.copy_(batch.is_current_agent, non_blocking=True)
        if getattr(self.static, "ptr", None) is not None:
            self.static.ptr.copy_(batch.ptr, non_blocking=True)
        self.static_next.copy_(nxt, non_blocking=True)

    def fits(self, batch: Batch, nxt: torch.Tensor) -> bool:
        return ((tuple(batch.x.shape), int(nxt.numel()), int(batch.batch.numel())) == self._shape
                and int(batch.edge_index.size(1)) <= self.edge_capacity
                and (getattr(self.static, "ptr", None) is None or getattr(batch, "ptr", None) is not None)
                and (getattr(batch, "max_graph_nodes", None) or 0) <= (getattr(self.static, "max_graph_nodes", None) or 0))

    # -- the step -------------------------------------------------------------------------------------------------
    def step(self, batch: Batch, next_positions: torch.Tensor) -> torch.Tensor:
        """Copies the batch into the static buffers and replays the captured step; returns the loss (a device scalar that
        the next replay overwrites).  A batch that does not fit the captured shapes takes the eager step."""
        if not self.fits(batch, next_positions):
            return masked_mse_train_step(self.model, self.optimizer, batch, next_positions, self.time_span, self.max_norm,
                                         distributed=False)
        self._load(batch, next_positions)
        self.graph.replay()
        self.replays += 1
        return self.loss

    def check(self) -> None:
        """Reads the device-side validation flags of the captured step (one synchronisation): raises GnodeError for an
        edge list with out-of-range ids, an edge leaving its graph, a graph larger than announced or an expired tcgen05
        barrier wait.  Call it every few hundred steps or at the end of an epoch."""
        from . import _lib
        if self._csr is not None:
            flags = self._csr[1].device_flags().tolist()
            if flags[0]:
                raise GnodeError("edge_index holds node ids outside [0, N) (captured step)")
            if flags[1]:
                G._raise_tile_error(int(flags[1]), deferred=True)
        _lib.tc_check(self.device)
