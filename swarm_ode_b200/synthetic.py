"""Deterministic synthetic inputs of the reference's shapes (no dataset / simulator needed).

The reference trains on HDF5 rollouts of the TA-RWARE simulator, which cannot run here (its A*
extension is absent).  These generators reproduce the *shape and statistics* of what
``WarehouseDataset`` + ``GraphConverter`` + ``collate_trajectory_batches`` hand to the model
(scripts/train_gde.py:278-375):

* warehouse layout formulas                      tarware/warehouse.py:215-256
* agents spawn on distinct highway cells         tarware/warehouse.py:640-651
* AGV / picker observation row layout            tarware/spaces/MultiAgentPartialObservationSpace.py:35-111
* picker rows zero-padded to the AGV row length  scripts/collect_data.py:99-118
* un-normalised integer coordinates              tarware/spaces/MultiAgentBaseObservationSpace.py:21 (ctor mix-up)
* 5-snapshot window graph, threshold 5.0         scripts/train_gde.py:116-184,308
* "next positions" = (x, y) of the first n rows of the next window graph   scripts/train_gde.py:319,336-355

Everything is vectorised numpy; edges are produced with the same float32 expression as
``GraphConverter`` so the result is identical to running the converter per trajectory.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from .data import Batch

_SIZES = {"tiny": (1, 3), "small": (2, 3), "medium": (2, 5), "large": (3, 5), "extralarge": (4, 7)}
_REQUESTS = {"tiny": 20, "small": 20, "medium": 20, "large": 40, "extralarge": 60}


@dataclass
class Layout:
    rows: int
    cols: int
    highway: np.ndarray      # [n_hw, 2] (y, x)
    shelves: np.ndarray      # [n_shelf, 2] (y, x)
    n_requests: int


def warehouse_layout(size: str = "medium", column_height: int = 8) -> Layout:
    shelf_rows, shelf_columns = _SIZES[size]
    lanes, col_w, bottom = 2, 2, 2
    rows = lanes + (column_height + lanes) * shelf_rows + bottom + 1
    cols = lanes + (col_w + lanes) * shelf_columns
    hy = {i + j for i in range(0, rows, column_height + lanes) for j in range(lanes)}
    hx = {i + j for i in range(0, cols, col_w + lanes) for j in range(lanes)}
    yy, xx = np.indices((rows, cols))
    hw = np.isin(xx, list(hx)) | np.isin(yy, list(hy)) | (yy >= rows - 1 - bottom)
    highway = np.stack([yy[hw], xx[hw]], axis=1).astype(np.int32)
    shelves = np.stack([yy[~hw], xx[~hw]], axis=1).astype(np.int32)
    return Layout(rows, cols, highway, shelves, _REQUESTS[size])


def node_dim(num_agvs: int, num_pickers: int, layout: Layout) -> int:
    return 3 + 4 * num_agvs + 4 * num_pickers + 2 * len(layout.shelves)


def warehouse_batch(num_graphs: int, num_agvs: int = 12, num_pickers: int = 7, size: str = "medium",
                    window: int = 5, threshold: float = 5.0, seed: int = 0,
                    return_positions: bool = False):
    """A collated batch of ``num_graphs`` full-window graphs.

    Returns ``(Batch, next_positions [B, n, 2])`` (CPU tensors), plus the integer positions
    ``[B, window, n, 2]`` when ``return_positions``.
    """
    rng = np.random.default_rng(seed)
    lay = warehouse_layout(size)
    B, W, n = num_graphs, window, num_agvs + num_pickers
    S = len(lay.shelves)
    D = node_dim(num_agvs, num_pickers, lay)

    # ---- positions: distinct highway cells, then a lazy 4-neighbour random walk clipped to the grid
    keys = rng.random((B, len(lay.highway)))
    start = np.argpartition(keys, n, axis=1)[:, :n]
    pos = np.empty((B, W, n, 2), dtype=np.int32)
    pos[:, 0] = lay.highway[start]
    steps = np.array([[0, 0], [1, 0], [-1, 0], [0, 1], [0, -1]], dtype=np.int32)
    for w in range(1, W):
        move = rng.random((B, n)) < 0.7
        direction = rng.integers(1, 5, size=(B, n))
        delta = steps[np.where(move, direction, 0)]
        nxt = pos[:, w - 1] + delta
        nxt[..., 0] = np.clip(nxt[..., 0], 0, lay.rows - 1)
        nxt[..., 1] = np.clip(nxt[..., 1], 0, lay.cols - 1)
        pos[:, w] = nxt
    # ---- targets (constant over the window), AGV status bits, shelf bits
    tgt = lay.shelves[rng.integers(0, S, size=(B, n))].astype(np.int32)
    tgt[rng.random((B, n)) < 0.3] = 0
    carry = (rng.random((B, num_agvs)) < 0.5)
    carry_req = carry & (rng.random((B, num_agvs)) < 0.5)
    toggle = rng.random((B, W, num_agvs)) < 0.2
    has_shelf = (rng.random((B, S)) < 0.9).astype(np.float32)
    req_keys = rng.random((B, S))
    requested = np.zeros((B, S), dtype=np.float32)
    np.put_along_axis(requested, np.argpartition(req_keys, lay.n_requests, axis=1)[:, :lay.n_requests], 1.0, axis=1)
    shelf_info = np.stack([has_shelf, requested * has_shelf], axis=2).reshape(B, 2 * S)

    # ---- per-agent info vectors
    base4 = np.concatenate([pos, np.broadcast_to(tgt[:, None], (B, W, n, 2))], axis=3).astype(np.float32)  # y,x,ty,tx
    info7 = np.concatenate([
        np.broadcast_to(carry[:, None, :, None], (B, W, num_agvs, 1)).astype(np.float32),
        np.broadcast_to(carry_req[:, None, :, None], (B, W, num_agvs, 1)).astype(np.float32),
        toggle[..., None].astype(np.float32),
        base4[:, :, :num_agvs]], axis=3)                                                        # [B,W,n_agv,7]

    x = np.zeros((B, W, n, D), dtype=np.float32)
    others = np.array([[o for o in range(n) if o != a] for a in range(n)], dtype=np.int64)       # [n, n-1]
    # AGV rows: own 7, every other agent's (y,x,ty,tx), shelf bits
    if num_agvs:
        x[:, :, :num_agvs, 0:7] = info7
        oth = base4[:, :, others[:num_agvs]]                                                     # [B,W,n_agv,n-1,4]
        x[:, :, :num_agvs, 7:7 + 4 * (n - 1)] = oth.reshape(B, W, num_agvs, 4 * (n - 1))
        x[:, :, :num_agvs, 7 + 4 * (n - 1):] = shelf_info[:, None, None, :]
    # picker rows: own 4, every AGV's 7-vector, every other picker's 4-vector, zero padding
    if num_pickers:
        x[:, :, num_agvs:, 0:4] = base4[:, :, num_agvs:]
        agv_block = info7.reshape(B, W, 1, 7 * num_agvs)
        x[:, :, num_agvs:, 4:4 + 7 * num_agvs] = agv_block
        if num_pickers > 1:
            pk = np.array([[o for o in range(num_agvs, n) if o != a] for a in range(num_agvs, n)], dtype=np.int64)
            x[:, :, num_agvs:, 4 + 7 * num_agvs:4 + 7 * num_agvs + 4 * (num_pickers - 1)] = \
                base4[:, :, pk].reshape(B, W, num_pickers, 4 * (num_pickers - 1))

    # ---- edges, GraphConverter order: spatial(k=0..W-1) then temporal (W-2 -> W-1), per graph
    posf = pos.reshape(B * W, n, 2).astype(np.float32)
    iu, ju = np.triu_indices(n, k=1)
    diff = posf[:, iu] - posf[:, ju]
    dist = np.sqrt((diff[..., 0] ** 2) + (diff[..., 1] ** 2))
    s_idx, p_idx = np.nonzero(dist < threshold)
    off = (s_idx.astype(np.int64)) * n                       # snapshot s = b*W + w starts at node s*n
    src = np.empty(2 * s_idx.size, dtype=np.int64)
    dst = np.empty(2 * s_idx.size, dtype=np.int64)
    src[0::2], dst[0::2] = off + iu[p_idx], off + ju[p_idx]
    src[1::2], dst[1::2] = off + ju[p_idx], off + iu[p_idx]
    gid = np.repeat(s_idx // W, 2)
    if W > 1:
        a = np.arange(n, dtype=np.int64)
        gb = np.arange(B, dtype=np.int64)[:, None] * (W * n)
        t_src = (gb + (W - 2) * n + a[None]).reshape(-1)
        t_dst = (gb + (W - 1) * n + a[None]).reshape(-1)
        src = np.concatenate([src, t_src]); dst = np.concatenate([dst, t_dst])
        gid = np.concatenate([gid, np.repeat(np.arange(B, dtype=np.int64), n)])
    order = np.argsort(gid, kind="stable")
    edge_index = torch.from_numpy(np.stack([src[order], dst[order]], axis=0))

    batch = Batch(x=torch.from_numpy(x.reshape(B * W * n, D)), edge_index=edge_index)
    batch.batch = torch.arange(B, dtype=torch.long).repeat_interleave(W * n)
    batch.ptr = torch.arange(B + 1, dtype=torch.long) * (W * n)
    mask = torch.zeros(B, W, n, dtype=torch.bool)
    mask[:, W - 1] = True
    batch.is_current_agent = mask.reshape(-1)
    batch.num_graphs = B
    batch.max_graph_nodes = W * n
    # (x, y) of snapshot 1 == first n rows of the next step's window graph (reference quirk)
    nxt_src = pos[:, min(1, W - 1)]
    next_positions = torch.from_numpy(np.stack([nxt_src[..., 1], nxt_src[..., 0]], axis=-1).astype(np.float32))
    if return_positions:
        return batch, next_positions, pos
    return batch, next_positions


def dense_batch(num_graphs: int, num_agents: int = 256, node_dim_: int = 64, seed: int = 0) -> Batch:
    """Config 4: ``num_agents`` nodes per graph, complete directed graph without self loops, N(0,1) features."""
    g = torch.Generator().manual_seed(seed)
    B, n = num_graphs, num_agents
    x = torch.randn(B * n, node_dim_, generator=g)
    i, j = torch.meshgrid(torch.arange(n), torch.arange(n), indexing="ij")
    keep = i != j
    local = torch.stack([i[keep], j[keep]], dim=0)                                   # [2, n(n-1)]
    offs = (torch.arange(B, dtype=torch.long) * n).view(B, 1, 1)
    edge_index = (local.unsqueeze(0) + offs).permute(1, 0, 2).reshape(2, -1)
    out = Batch(x=x, edge_index=edge_index)
    out.batch = torch.arange(B, dtype=torch.long).repeat_interleave(n)
    out.ptr = torch.arange(B + 1, dtype=torch.long) * n
    out.is_current_agent = torch.ones(B * n, dtype=torch.bool)
    out.num_graphs = B
    out.max_graph_nodes = n
    return out


def geometric_batch(num_graphs: int, num_agents: int, node_dim_: int = 128, threshold: float = 5.0,
                    seed: int = 0) -> Batch:
    """Config 5: random geometric graph on a ceil(sqrt(20 n))^2 grid (degree roughly constant in n)."""
    rng = np.random.default_rng(seed)
    B, n = num_graphs, num_agents
    side = int(np.ceil(np.sqrt(20 * n)))
    pos = rng.integers(0, side, size=(B, n, 2)).astype(np.float32)
    srcs, dsts = [], []
    chunk = max(1, (1 << 22) // (n * n))
    for b0 in range(0, B, chunk):
        p = pos[b0:b0 + chunk]
        d = np.sqrt(((p[:, :, None, :] - p[:, None, :, :]) ** 2).sum(-1))
        hit = (d < threshold) & ~np.eye(n, dtype=bool)[None]
        bb, ii, jj = np.nonzero(hit)
        srcs.append((bb + b0) * n + ii); dsts.append((bb + b0) * n + jj)
    edge_index = torch.from_numpy(np.stack([np.concatenate(srcs), np.concatenate(dsts)]).astype(np.int64))
    g = torch.Generator().manual_seed(seed)
    out = Batch(x=torch.randn(B * n, node_dim_, generator=g), edge_index=edge_index)
    out.batch = torch.arange(B, dtype=torch.long).repeat_interleave(n)
    out.ptr = torch.arange(B + 1, dtype=torch.long) * n
    out.is_current_agent = torch.ones(B * n, dtype=torch.bool)
    out.num_graphs = B
    out.max_graph_nodes = n
    return out


def init_weights(module: torch.nn.Module, seed: int = 1, conv3_scale: float = 1.0) -> None:
    """U(+-1/sqrt(fan_in)) for every weight and bias (seeded); optionally damp conv3 (keeps the field
    non-stiff for the dopri5 configurations).  Same values whatever the device."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(module.named_parameters()):
            if p.dim() == 2:
                fan_in = p.shape[1]
            else:
                parent = name.rsplit(".", 1)[0]
                w = dict(module.named_parameters()).get(parent + ".weight")
                fan_in = w.shape[1] if w is not None else p.shape[0]
            bound = 1.0 / np.sqrt(max(fan_in, 1))
            vals = (torch.rand(p.shape, generator=g) * 2 - 1) * bound
            if conv3_scale != 1.0 and ".conv3." in "." + name:
                vals = vals * conv3_scale
            p.copy_(vals.to(p.device))


def multi_agent_observation(rng, num_agvs: int, num_pickers: int, racks, grid=(25, 22), p_requested: float = 0.35,
                            agv_target_prob: float = 0.0):
    """One joint observation in the row layout ``MultiAgentGraphConverter`` reads (scripts/run_gnode.py:1079-1100):
    AGV rows ``[carrying, carrying_requested, toggle, y, x, target_y, target_x, ...]``, picker rows ``[y, x, target_y,
    target_x, ...]``, and in row 0, after the ``7 + 4 (n - 1)`` agent entries, one ``(has_shelf, is_requested)`` pair per
    rack.  ``racks``: list of ``(x, y, group)`` tuples.  At most ONE AGV gets a target (``agv_target_prob``): two AGVs with
    targets make the reference's converter raise (hetero.MultiAgentGraphConverter docstring)."""
    import numpy as np
    n, n_loc = num_agvs + num_pickers, len(racks)
    width = 7 + 4 * (n - 1) + 2 * n_loc
    obs = np.zeros((n, width), dtype=np.float32)
    obs[:num_agvs, 0:3] = rng.integers(0, 2, (num_agvs, 3))
    obs[:num_agvs, 3] = rng.integers(0, grid[1], num_agvs)
    obs[:num_agvs, 4] = rng.integers(0, grid[0], num_agvs)
    obs[num_agvs:, 0] = rng.integers(0, grid[1], num_pickers)
    obs[num_agvs:, 1] = rng.integers(0, grid[0], num_pickers)
    if num_agvs and rng.random() < agv_target_prob:
        a, r = int(rng.integers(0, num_agvs)), int(rng.integers(0, n_loc))
        obs[a, 5], obs[a, 6] = racks[r][1], racks[r][0]
    tail = np.zeros(2 * n_loc, dtype=np.float32)
    tail[0::2] = rng.random(n_loc) < 0.8
    tail[1::2] = rng.random(n_loc) < p_requested
    obs[0, 7 + 4 * (n - 1):] = tail
    return obs


def rack_locations(rng, n_loc: int, grid=(25, 22), groups: int = 4):
    """``n_loc`` distinct rack cells as hashable ``(x, y, group)`` tuples (what the converter's callers pass)."""
    cells = rng.permutation(grid[0] * grid[1])[:n_loc]
    return [(int(c % grid[0]) + 1, int(c // grid[0]) + 1, int(g)) for c, g in zip(cells, rng.integers(0, groups, n_loc))]
