#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own code (``/root/reference/scripts/train_gde.py``).

The reference cannot be imported as-is here: its arithmetic lives in ``torch_geometric`` and ``torchdiffeq``,
which are not installed (and not vendored in the reference).  This script imports ``scripts/train_gde.py``
with those two libraries (plus the unused h5py / wandb / matplotlib) replaced in ``sys.modules``:

* ``torch_geometric.data.Data`` / ``Batch``  -> ``oracle.pyg_ref.RefData`` / ``RefBatch`` (containers only)
* ``torch_geometric.nn.SAGEConv``           -> ``oracle.pyg_ref.SAGEConvRef``   (restated arithmetic)
* ``torchdiffeq.odeint``                    -> ``oracle.torchdiffeq_ref.odeint_ref`` (restated arithmetic)

so that what runs is

* **graph_converter.npz** -- the reference's ``GraphConverter`` (scripts/train_gde.py:108-271) and
  ``collate_trajectory_batches`` (:363-375), numpy/torch code of the REFERENCE ITSELF, bit for bit; only
  the ``Data`` container is a stand-in.  These vectors pin edge construction to the real reference.
* **graph_ode.npz** -- the reference's ``GraphODE`` / ``GraphODEFunc`` modules (:20-106: closure, solver
  call, decoder loop, return dict) on top of the restated third-party numerics.  These pin the wiring of
  the reference modules; the SAGEConv / odeint arithmetic underneath is the oracle's ("parity unpinned").

Runs only where ``/root/reference`` exists (the build container); the vectors it writes are committed.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("GNODE_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")


def import_reference_train_gde():
    from oracle.pyg_ref import RefBatch, RefData, SAGEConvRef
    from oracle.torchdiffeq_ref import odeint_ref

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Unused:  # names the reference imports but never touches on this path
        def __init__(self, *a, **k):
            raise RuntimeError("stubbed third-party class was instantiated")

    stub("torch_geometric")
    stub("torch_geometric.nn", HeteroConv=_Unused, SAGEConv=SAGEConvRef, GATConv=_Unused, Linear=_Unused)
    stub("torch_geometric.data", HeteroData=_Unused, Batch=RefBatch, Data=RefData)
    stub("torchdiffeq", odeint=odeint_ref)
    for name in ("h5py", "wandb", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                stub(name)
    spec = importlib.util.spec_from_file_location("ref_train_gde", os.path.join(REF, "scripts", "train_gde.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)  # module level only defines classes; the training run sits under __main__
    return mod


def observations(rng, n_agv, n_pick, D, grid=(25, 22), move_p=0.7, prev=None):
    """One snapshot of per-agent observation rows in the reference layout (AGV: y,x at cols 3,4; picker: cols 0,1),
    integer coordinates; agents random-walk between snapshots."""
    n = n_agv + n_pick
    if prev is None:
        pos = np.stack([rng.integers(0, grid[0], n), rng.integers(0, grid[1], n)], 1)
    else:
        step = rng.integers(0, 4, n)
        d = np.array([[1, 0], [-1, 0], [0, 1], [0, -1]])[step] * (rng.random(n) < move_p)[:, None]
        pos = np.clip(prev + d, 0, np.array(grid) - 1)
    obs = rng.integers(0, 3, (n, D)).astype(np.float32)
    obs[:n_agv, 3:5] = pos[:n_agv]
    obs[n_agv:, 0:2] = pos[n_agv:]
    return obs, pos


def golden_converter(ref):
    out = {}
    cases = [  # name, n_agv, n_pick, D, threshold, window, steps
        ("medium", 12, 7, 40, 5.0, 5, 8),
        ("default_thr", 3, 2, 12, 3.0, 5, 7),
        ("one_agent", 1, 0, 8, 5.0, 5, 3),
        ("window2", 4, 3, 16, 5.0, 2, 5),
        ("dense", 6, 6, 10, 100.0, 5, 6),
        ("boundary", 5, 0, 8, 5.0, 3, 1),
    ]
    for name, na, npk, D, thr, win, steps in cases:
        rng = np.random.default_rng({"medium": 0, "default_thr": 1, "one_agent": 2, "window2": 3, "dense": 4, "boundary": 5}[name])
        conv = ref.GraphConverter(na, npk, distance_threshold=thr, temporal_window=win)
        pos = None
        graphs = []
        for s in range(steps):
            obs, pos = observations(rng, na, npk, D, prev=pos)
            if name == "boundary":  # distances exactly 5 (3-4-5 triangle, axis aligned) must NOT connect; 4.9.. must
                obs[:, 3:5] = np.array([[0, 0], [3, 4], [5, 0], [0, 5], [4, 2]], dtype=np.float32)
            g = conv._build_graph_from_observation(obs)
            graphs.append(g)
            out[f"{name}/obs{s}"] = obs
            out[f"{name}/x{s}"] = g.x.numpy()
            out[f"{name}/edge_index{s}"] = g.edge_index.numpy()
            out[f"{name}/is_current_agent{s}"] = g.is_current_agent.numpy()
        out[f"{name}/meta"] = np.array([na, npk, D, win, steps], dtype=np.int64)
        out[f"{name}/threshold"] = np.array([thr], dtype=np.float64)
        if name == "medium":  # collate (scripts/train_gde.py:363-375) over the last 4 graphs
            items = [ref.TrajectoryBatch(g, torch.full((na + npk, 2), float(i))) for i, g in enumerate(graphs[-4:])]
            b = ref.collate_trajectory_batches(items)
            out["medium/collate_x"] = b.graphs.x.numpy()
            out["medium/collate_edge_index"] = b.graphs.edge_index.numpy()
            out["medium/collate_batch"] = b.graphs.batch.numpy()
            out["medium/collate_mask"] = b.graphs.is_current_agent.numpy()
            out["medium/collate_next"] = b.next_positions.numpy()
    # ragged rows (picker rows shorter than AGV rows) are zero padded by _standardize_observations
    conv = ref.GraphConverter(2, 2, distance_threshold=5.0)
    ragged = [np.arange(9, dtype=np.float32), np.arange(9, dtype=np.float32) + 1, np.arange(5, dtype=np.float32),
              np.arange(6, dtype=np.float32)]
    g = conv._build_graph_from_observation(ragged)
    out["ragged/x"] = g.x.numpy()
    out["ragged/edge_index"] = g.edge_index.numpy()
    for i, r in enumerate(ragged):
        out[f"ragged/row{i}"] = r
    np.savez_compressed(os.path.join(OUT, "graph_converter.npz"), **out)
    print("graph_converter.npz:", len(out), "arrays")


def golden_graph_ode(ref):
    import swarm_ode_b200.synthetic as syn

    out = {}
    batch, nxt = syn.warehouse_batch(3, num_agvs=4, num_pickers=3, seed=7)
    D = batch.x.shape[1]
    from oracle.pyg_ref import RefBatch
    rb = RefBatch(x=batch.x.clone(), edge_index=batch.edge_index.clone())
    rb.batch, rb.is_current_agent = batch.batch, batch.is_current_agent
    out["x"], out["edge_index"], out["batch"] = batch.x.numpy(), batch.edge_index.numpy(), batch.batch.numpy()
    out["is_current_agent"], out["next_positions"] = batch.is_current_agent.numpy(), nxt.numpy()
    for solver, tspan in (("euler", [0.0, 1.0]), ("midpoint", [0.0, 0.5, 1.0]), ("rk4", [0.0, 1.0]),
                          ("rk4_multi", [0.0, 1.0, 2.0, 3.0]), ("dopri5", [0.0, 1.0])):
        method = solver.split("_")[0]
        torch.manual_seed(0)
        model = ref.GraphODE(node_dim=D, num_agvs=4, num_pickers=3, hidden_dim=32, ode_solver=method)
        syn.init_weights(model, seed=11, conv3_scale=0.1)
        res = model(rb, torch.tensor(tspan))
        assert set(res) == {"trajectories", "node_features", "batch"}
        for k, v in model.state_dict().items():   # same seed for every solver: stored once
            out[f"param/{k}"] = v.detach().numpy()
        out[f"{solver}/t"] = np.array(tspan, dtype=np.float32)
        out[f"{solver}/trajectories"] = res["trajectories"].detach().numpy()
        out[f"{solver}/node_features"] = res["node_features"].detach().numpy()
        if method != "dopri5":
            # the training step of scripts/train_gde.py:484-493 (device-mismatch bug at :476/:490 aside)
            loss = torch.nn.functional.mse_loss(res["trajectories"][1][rb.is_current_agent], nxt.view(-1, 2))
            loss.backward()
            out[f"{solver}/loss"] = np.array([float(loss.detach())])
            for k, p in model.named_parameters():
                out[f"{solver}/grad/{k}"] = p.grad.detach().numpy()
    model = ref.GraphODE(node_dim=D, num_agvs=4, num_pickers=3, hidden_dim=32, ode_solver="euler")
    syn.init_weights(model, seed=11, conv3_scale=0.1)
    out["predict_trajectory_3"] = model.predict_trajectory(rb, 3).detach().numpy()
    np.savez_compressed(os.path.join(OUT, "graph_ode.npz"), **out)
    print("graph_ode.npz:", len(out), "arrays")


def reference_classes(rel_path, names):
    """The reference's hetero scripts cannot be imported (module-level argparse / gym / rware code), so the class
    definitions are taken from the file by name and executed with the stand-ins in scope.  Nothing is copied into the
    repository: the source text is read from /root/reference at generation time."""
    import ast
    import torch.nn as nn
    from oracle.pyg_ref import HeteroConvRef, SAGEConvRef
    from oracle.torchdiffeq_ref import odeint_ref
    path = os.path.join(REF, rel_path)
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "nn": nn, "SAGEConv": SAGEConvRef, "HeteroConv": HeteroConvRef, "odeint": odeint_ref}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns


def hetero_case(seed=3, n_agv=6, n_pick=3, n_loc=20, dims=(11, 8, 5)):
    from oracle.pyg_ref import RefHeteroData
    g = torch.Generator().manual_seed(seed)
    d = RefHeteroData()
    d["agv"].x = torch.randn(n_agv, dims[0], generator=g)
    d["picker"].x = torch.randn(n_pick, dims[1], generator=g)
    d["location"].x = torch.randn(n_loc, dims[2], generator=g)
    tgt = torch.randint(0, n_loc, (n_agv,), generator=g)
    ar = torch.arange(n_agv)
    d["agv", "targets", "location"].edge_index = torch.stack([ar, tgt])
    d["location", "is targeted by", "agv"].edge_index = torch.stack([tgt, ar])
    pairs = [(i, j) for i in range(n_agv) for j in range(n_agv) if i != j and (i + j) % 3 == 0]
    d["agv", "communicates", "agv"].edge_index = torch.tensor(pairs).t().contiguous()
    mg = torch.randint(0, n_pick, (n_loc,), generator=g)
    d["picker", "manages", "location"].edge_index = torch.stack([mg, torch.arange(n_loc)])
    coop = torch.randint(0, n_pick, (n_agv,), generator=g)
    d["agv", "cooperates with", "picker"].edge_index = torch.stack([ar, coop])     # picker 'coop' may receive several AGVs
    d["picker", "helps", "agv"].edge_index = torch.stack([coop[:4], ar[:4]])       # AGVs 4, 5 get no help edge
    return d


def golden_hetero():
    out = {}
    d = hetero_case()
    for k in ("agv", "picker", "location"):
        out[f"x/{k}"] = d[k].x.numpy()
    for et, ei in d.edge_index_dict.items():
        out["edge/" + "__".join(et)] = ei.numpy()
    dims = {"agv": 11, "picker": 8, "location": 5}
    for tag, rel, kwargs in (("joint", "scripts/gnode.py", dict(hidden_dim=32, num_layers=2, ode_hidden_dim=16)),
                             ("typed", "scripts/run_gnode.py", dict(action_size=5, hidden_dim=32, num_layers=2, ode_hidden_dim=16))):
        ns = reference_classes(rel, {"HeteroGraphODENetwork", "ODEFunction"})
        torch.manual_seed(21)
        model = ns["HeteroGraphODENetwork"](dims, **kwargs)
        with torch.no_grad():
            for p in model.parameters():      # damp the MLP field a little so that default-tolerance dopri5 stays cheap
                p.mul_(0.7)
            res = model(d, integration_time=1.0 if tag == "typed" else 0.5)
        for k, v in model.state_dict().items():
            out[f"{tag}/param/{k}"] = v.numpy()
        for k, v in res.items():
            out[f"{tag}/{k}"] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "hetero.npz"), **out)
    print("hetero.npz:", len(out), "arrays")


def mac_observation(rng, n_agv, n_pick, n_loc, racks, agv_targets=(), picker_targets=(), grid=(10, 10), p_requested=0.35):
    """One observation array [n_agv + n_pick, 7 + 4 (n - 1) + 2 n_loc] with the row layout the reference's converter
    reads (scripts/run_gnode.py:1079-1100): AGV rows [carry, carry_req, toggle, y, x, ty, tx, ...], picker rows
    [y, x, ty, tx, ...]; the shelf pairs are read from row 0.  `agv_targets` / `picker_targets`: {agent: rack index}."""
    D = 7 + 4 * (n_agv + n_pick - 1) + 2 * n_loc
    obs = np.zeros((n_agv + n_pick, D), dtype=np.float32)
    for a in range(n_agv):
        obs[a, :3] = rng.integers(0, 2, 3)
        obs[a, 3:5] = rng.integers(0, grid[0], 2)
    for p in range(n_pick):
        obs[n_agv + p, 0:2] = rng.integers(0, grid[0], 2)
    for a, r in dict(agv_targets).items():
        obs[a, 5], obs[a, 6] = (racks[r][1], racks[r][0]) if r >= 0 else (grid[0] + 3, grid[1] + 3)   # (ty, tx); -1: off every rack
    for p, r in dict(picker_targets).items():
        obs[n_agv + p, 2], obs[n_agv + p, 3] = racks[r][1], racks[r][0]
    shelf = np.zeros(2 * n_loc, dtype=np.float32)
    shelf[0::2] = rng.random(n_loc) < 0.8
    shelf[1::2] = rng.random(n_loc) < p_requested
    obs[0, 7 + 4 * (n_agv + n_pick - 1):] = shelf
    obs[1:, 7:] = rng.integers(0, 5, size=(n_agv + n_pick - 1, D - 7))      # other rows' tails are never read by the converter
    return obs


def golden_multi_agent_converter():
    """tests/golden/multi_agent_converter.npz: the reference's MultiAgentGraphConverter (scripts/run_gnode.py:1040-1326),
    executed from its source text (the module itself cannot be imported: module-level wandb / gym / argparse code), on
    synthetic observations -- node features and the six relations per case, and for the inputs the reference's code
    cannot process, the type of the exception it raises."""
    import ast
    from oracle.pyg_ref import RefHeteroData
    path = os.path.join(REF, "scripts/run_gnode.py")
    tree = ast.parse(open(path).read())
    ns = {"np": np, "torch": torch, "HeteroData": RefHeteroData}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == "MultiAgentGraphConverter":
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    Conv = ns["MultiAgentGraphConverter"]
    rng = np.random.default_rng(5)
    out = {}

    def racks_for(n_loc, grid):
        cells = rng.permutation(grid[0] * grid[1])[:n_loc]
        return [(int(c % grid[0]) + 1, int(c // grid[0]) + 1, int(g)) for c, g in zip(cells, rng.integers(0, 4, n_loc))]

    def record(tag, conv, obs, racks):
        out[f"{tag}/obs"] = obs
        out[f"{tag}/racks"] = np.asarray(racks, dtype=np.int64)
        try:
            d = conv._build_graph_from_observation(obs, racks)
        except Exception as e:  # noqa: BLE001 -- the exception TYPE is the golden value
            out[f"{tag}/raises"] = np.array(type(e).__name__)
            return
        for k in ("agv", "picker", "location"):
            out[f"{tag}/x/{k}"] = d[k].x.numpy()
        for et, ei in d.edge_index_dict.items():
            out[f"{tag}/edge/" + "__".join(et)] = ei.numpy()

    shapes = {"small": (4, 3, 12, (10, 10)), "medium": (19, 9, 160, (25, 22))}
    for name, (na, npk, nl, grid) in shapes.items():
        racks = racks_for(nl, grid)
        record(f"{name}/idle", Conv(na, npk), mac_observation(rng, na, npk, nl, racks, grid=grid), racks)
        record(f"{name}/one_agv_target", Conv(na, npk), mac_observation(rng, na, npk, nl, racks, agv_targets={1: 2}, grid=grid), racks)
        record(f"{name}/agv_target_off_rack", Conv(na, npk), mac_observation(rng, na, npk, nl, racks, agv_targets={na - 1: -1}, grid=grid), racks)
        record(f"{name}/nothing_requested", Conv(na, npk), mac_observation(rng, na, npk, nl, racks, grid=grid, p_requested=0.0), racks)
        record(f"{name}/two_agv_targets", Conv(na, npk), mac_observation(rng, na, npk, nl, racks, agv_targets={0: 1, 2: 3}, grid=grid), racks)
        record(f"{name}/picker_target", Conv(na, npk), mac_observation(rng, na, npk, nl, racks, picker_targets={1: 4}, grid=grid), racks)
        # a REUSED converter: the second call's edges are those of the first observation (stale agent / shelf info)
        conv = Conv(na, npk)
        record(f"{name}/reuse_first", conv, mac_observation(rng, na, npk, nl, racks, agv_targets={0: 5}, grid=grid), racks)
        record(f"{name}/reuse_second", conv, mac_observation(rng, na, npk, nl, racks, grid=grid), racks)
    # rack entries that are ndarray rows (what h5py hands back): unhashable in position_to_sections.get
    na, npk, nl, grid = shapes["small"]
    racks = racks_for(nl, grid)
    obs = mac_observation(rng, na, npk, nl, racks, grid=grid)
    out["small/ndarray_racks/obs"], out["small/ndarray_racks/racks"] = obs, np.asarray(racks, dtype=np.int64)
    try:
        Conv(na, npk)._build_graph_from_observation(obs, np.asarray(racks))
    except Exception as e:  # noqa: BLE001
        out["small/ndarray_racks/raises"] = np.array(type(e).__name__)
    np.savez_compressed(os.path.join(OUT, "multi_agent_converter.npz"), **out)
    print("multi_agent_converter.npz:", len(out), "arrays;", sorted(k for k in out if k.endswith("raises")),
          [str(out[k]) for k in sorted(out) if k.endswith("raises")])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["converter", "graph_ode", "hetero", "multi_agent_converter"]
    if "multi_agent_converter" in which:
        golden_multi_agent_converter()
    if "converter" in which or "graph_ode" in which:
        ref = import_reference_train_gde()
        if "converter" in which:
            golden_converter(ref)
        if "graph_ode" in which:
            golden_graph_ode(ref)
    if "hetero" in which:
        golden_hetero()
