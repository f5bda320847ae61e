"""Device-resident replacement of the reference's dataset / loader (scripts/train_gde.py:278-375).

The reference builds every window graph at dataset-construction time with a Python O(n^2) pair loop per step
(``WarehouseDataset._load_all_sequences`` -> ``GraphConverter._build_graph_from_observation``), keeps the graphs on the
host, and collates + uploads a batch per training step.  Here an episode's observations go to the GPU once, all of its
window graphs are built there in one call (``build_episode_batch`` -> ``gnode_window_graphs``, bit-exact), and a batch is
a gather of whole graphs on the device: no host collation, no per-step upload of node features.

Sources of episodes (``observations [n_steps, n_agents, D]`` float32 + ``num_agvs`` / ``num_pickers``):
  * ``.npz`` shards written by ``scripts/h5_to_npz.py`` from the reference's HDF5 files (collect_data.py:20-44,137-170);
  * the HDF5 files themselves when ``h5py`` is importable (it is not in this image);
  * synthetic episodes (``synthetic_episode``) in the reference's observation layout.
"""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .data import Batch, TrajectoryBatch, build_episode_batch
from . import synthetic


class Episode:
    def __init__(self, observations: np.ndarray, num_agvs: int, num_pickers: int):
        obs = np.asarray(observations, dtype=np.float32)
        if obs.ndim != 3 or obs.shape[1] != num_agvs + num_pickers:
            raise ValueError(f"episode observations must be [n_steps, {num_agvs + num_pickers}, D], got {obs.shape}")
        self.observations, self.num_agvs, self.num_pickers = obs, int(num_agvs), int(num_pickers)


def synthetic_episode(n_steps: int, num_agvs: int = 12, num_pickers: int = 7, size: str = "medium", seed: int = 0) -> Episode:
    """A lazy random walk of the agents over ``n_steps`` steps in the reference's partial-observation row layout."""
    b, _ = synthetic.warehouse_batch(1, num_agvs=num_agvs, num_pickers=num_pickers, size=size, window=n_steps, seed=seed)
    n = num_agvs + num_pickers
    return Episode(b.x.numpy().reshape(n_steps, n, -1), num_agvs, num_pickers)


def load_episodes(path: str) -> List[Episode]:
    """Episodes of one file: ``.npz`` shard (``scripts/h5_to_npz.py``) or the reference's ``.h5`` (needs h5py)."""
    if path.endswith(".npz"):
        z = np.load(path)
        ids = sorted({k.split("/")[0] for k in z.files if k.startswith("episode_")})
        return [Episode(z[f"{e}/observations"], int(z[f"{e}/num_agvs"]), int(z[f"{e}/num_pickers"])) for e in ids]
    if path.endswith((".h5", ".hdf5")):
        try:
            import h5py
        except ImportError as e:
            raise RuntimeError("reading the reference's HDF5 files needs h5py; convert them with scripts/h5_to_npz.py "
                               "where h5py is available and pass the .npz shards") from e
        out = []
        with h5py.File(path, "r") as f:   # schema of collect_data.py:20-44,137-170, read order of train_gde.py:296-316
            for key in sorted(k for k in f.keys() if k.startswith("episode_")):
                g = f[key]
                steps = g["steps"]
                obs = np.stack([_pad_rows(steps[s]["observations"][:]) for s in sorted(steps.keys())], axis=0)
                out.append(Episode(obs, int(g["metadata"].attrs["num_agvs"]), int(g["metadata"].attrs["num_pickers"])))
        return out
    raise ValueError(f"unknown episode file type: {path}")


def _pad_rows(rows) -> np.ndarray:
    rows = np.asarray(rows)
    if rows.dtype != object:
        return rows.astype(np.float32)
    width = max(len(r) for r in rows)
    out = np.zeros((len(rows), width), dtype=np.float32)
    for i, r in enumerate(rows):
        out[i, :len(r)] = np.asarray(r, dtype=np.float32)
    return out


class WarehouseDataset:
    """``len(ds)`` (current graph, next positions) pairs like scripts/train_gde.py:278-334, resident on ``device``.

    ``ds[i]`` returns the reference's ``TrajectoryBatch(graph, next_positions)`` of one sample (device tensors);
    ``ds.collate(indices)`` returns the collated ``TrajectoryBatch`` of ``collate_trajectory_batches`` (:363-375) for a list
    of sample indices -- built by gathering whole graphs on the device."""

    def __init__(self, episodes: Sequence[Episode], device, distance_threshold: float = 5.0, temporal_window: int = 5):
        if not episodes:
            raise ValueError("no episodes")
        self.device = torch.device(device)
        self.num_agvs, self.num_pickers = episodes[0].num_agvs, episodes[0].num_pickers
        n = self.num_agvs + self.num_pickers
        xs, eis, curs, ptrs, eptrs, poss, samples = [], [], [], [0], [0], [], []
        node_off = edge_off = graph_off = 0
        for ep in episodes:
            if (ep.num_agvs, ep.num_pickers) != (self.num_agvs, self.num_pickers):
                raise ValueError("all episodes of a dataset must have the same agents")
            obs = torch.from_numpy(ep.observations).to(self.device)
            b = build_episode_batch(obs, ep.num_agvs, ep.num_pickers, distance_threshold, temporal_window)
            T = int(obs.shape[0])
            xs.append(b.x)
            eis.append(b.edge_index + node_off)
            curs.append(b.is_current_agent)
            p, e = b.ptr.cpu().numpy(), b.edge_ptr.cpu().numpy()
            ptrs += (p[1:] + node_off).tolist()
            eptrs += (e[1:] + edge_off).tolist()
            # positions the reference reads back from a window graph: its FIRST n rows (oldest snapshot of the window;
            # reference quirk kept, train_gde.py:336-355): (x, y) = columns (4, 3) of AGV rows, (1, 0) of picker rows
            first = b.x[b.ptr[:-1].unsqueeze(1) + torch.arange(n, device=self.device).unsqueeze(0)]      # [T, n, D]
            agv = first[:, :self.num_agvs][:, :, [4, 3]]
            pick = first[:, self.num_agvs:][:, :, [1, 0]]
            poss.append(torch.cat([agv, pick], dim=1))
            samples += [graph_off + t for t in range(T - 1)]      # the last step of an episode has no successor
            node_off += int(p[-1]); edge_off += int(e[-1]); graph_off += T
        self.x = torch.cat(xs, dim=0)
        self.edge_index = torch.cat(eis, dim=1)
        self.is_current_agent = torch.cat(curs, dim=0)
        self.positions = torch.cat(poss, dim=0)                   # [graphs, n, 2]
        self._ptr = np.asarray(ptrs, dtype=np.int64)              # host copies: batch sizes are known without a device sync
        self._eptr = np.asarray(eptrs, dtype=np.int64)
        self._samples = np.asarray(samples, dtype=np.int64)
        self.node_dim = int(self.x.shape[1])
        self.max_graph_nodes = int((self._ptr[1:] - self._ptr[:-1]).max())

    def __len__(self) -> int:
        return int(self._samples.size)

    def __getitem__(self, i: int) -> TrajectoryBatch:
        tb = self.collate([i])
        tb.next_positions = tb.next_positions[0]
        return tb

    def collate(self, indices: Sequence[int]) -> TrajectoryBatch:
        g = self._samples[np.asarray(indices, dtype=np.int64)]
        n0, n1, e0, e1 = self._ptr[g], self._ptr[g + 1], self._eptr[g], self._eptr[g + 1]
        ncnt, ecnt = n1 - n0, e1 - e0
        new_ptr = np.concatenate([[0], np.cumsum(ncnt)])
        dev = self.device
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=True)
        B, N, E = len(g), int(new_ptr[-1]), int(ecnt.sum())
        ncnt_d, ecnt_d = up(ncnt), up(ecnt)
        # node gather: graph b's rows [n0[b], n1[b]) land at [new_ptr[b], new_ptr[b+1])
        shift = torch.repeat_interleave(up(n0 - new_ptr[:-1]), ncnt_d, output_size=N)
        nidx = torch.arange(N, device=dev) + shift
        eshift = torch.repeat_interleave(up(e0 - np.concatenate([[0], np.cumsum(ecnt)])[:-1]), ecnt_d, output_size=E)
        eidx = torch.arange(E, device=dev) + eshift
        out = Batch(x=self.x[nidx], edge_index=self.edge_index[:, eidx] - torch.repeat_interleave(up(n0 - new_ptr[:-1]), ecnt_d, output_size=E))
        out.batch = torch.repeat_interleave(torch.arange(B, device=dev), ncnt_d, output_size=N)
        out.ptr = up(new_ptr)
        out.is_current_agent = self.is_current_agent[nidx]
        out.num_graphs, out.max_graph_nodes = B, int(ncnt.max())
        return TrajectoryBatch(graphs=out, next_positions=self.positions[up(g + 1)])

    @classmethod
    def from_files(cls, paths: Sequence[str], device, **kw) -> "WarehouseDataset":
        eps: List[Episode] = []
        for p in paths:
            if not os.path.exists(p):
                raise FileNotFoundError(p)
            eps += load_episodes(p)
        return cls(eps, device, **kw)


def split_indices(n: int, train_fraction: float = 0.8, seed: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Random train / validation split of ``range(n)`` (scripts/train_gde.py:448-450; the remainder of the reference's two
    ``int(...)`` sizes goes to the validation part instead of raising)."""
    perm = np.random.default_rng(seed).permutation(n)
    k = int(train_fraction * n)
    return perm[:k], perm[k:]
