"""Device-resident dataset / loader (swarm_ode_b200/dataset.py) against the reference's construction: a fresh
GraphConverter per episode fed step by step, extract_positions_from_graph of the NEXT step's graph, and
collate_trajectory_batches (scripts/train_gde.py:296-334, 336-375) -- bit for bit."""
import numpy as np
import pytest
import torch

import swarm_ode_b200 as S
from swarm_ode_b200.data import extract_positions_from_graph
from swarm_ode_b200.dataset import Episode, WarehouseDataset, load_episodes, split_indices, synthetic_episode


def _reference_pairs(ep, thr=5.0, W=5):
    conv = S.GraphConverter(ep.num_agvs, ep.num_pickers, distance_threshold=thr, temporal_window=W)
    steps = [conv._build_graph_from_observation(ep.observations[t]) for t in range(ep.observations.shape[0])]
    pos = [extract_positions_from_graph(g, ep.num_agvs, ep.num_pickers) for g in steps]
    return [S.TrajectoryBatch(steps[i], pos[i + 1]) for i in range(len(steps) - 1)]


def test_npz_round_trip_and_split(tmp_path):
    eps = [synthetic_episode(7, 3, 2, seed=s) for s in range(2)]
    out = {}
    for i, ep in enumerate(eps):
        out[f"episode_{i:06d}/observations"] = ep.observations
        out[f"episode_{i:06d}/num_agvs"] = np.int64(ep.num_agvs)
        out[f"episode_{i:06d}/num_pickers"] = np.int64(ep.num_pickers)
    p = str(tmp_path / "shard.npz")
    np.savez_compressed(p, **out)
    back = load_episodes(p)
    assert len(back) == 2 and all(np.array_equal(a.observations, b.observations) for a, b in zip(eps, back))
    assert (back[0].num_agvs, back[0].num_pickers) == (3, 2)
    tr, va = split_indices(11, 0.8, seed=0)
    assert len(tr) == 8 and len(va) == 3 and sorted(np.concatenate([tr, va]).tolist()) == list(range(11))
    with pytest.raises(ValueError):
        Episode(np.zeros((4, 3, 9), dtype=np.float32), 3, 2)


@pytest.mark.gpu
def test_device_dataset_equals_reference_construction(cuda):
    eps = [synthetic_episode(T, 4, 3, seed=s) for s, T in ((0, 9), (1, 3), (2, 14))]
    ds = WarehouseDataset(eps, cuda, distance_threshold=5.0, temporal_window=5)
    ref = sum((_reference_pairs(ep) for ep in eps), [])
    assert len(ds) == len(ref) == 8 + 2 + 13
    for idx in ([0], [5, 1, 22, 9], list(range(len(ref))), [10, 10, 3]):
        got = ds.collate(idx)
        want = S.collate_trajectory_batches([ref[i] for i in idx])
        assert torch.equal(got.graphs.x.cpu(), want.graphs.x)
        assert torch.equal(got.graphs.edge_index.cpu(), want.graphs.edge_index)
        assert torch.equal(got.graphs.batch.cpu(), want.graphs.batch) and torch.equal(got.graphs.ptr.cpu(), want.graphs.ptr)
        assert torch.equal(got.graphs.is_current_agent.cpu(), want.graphs.is_current_agent)
        assert torch.equal(got.next_positions.cpu(), want.next_positions)
        assert got.graphs.max_graph_nodes == want.graphs.max_graph_nodes
    one = ds[4]
    assert torch.equal(one.graphs.x.cpu(), ref[4].graphs.x) and torch.equal(one.next_positions.cpu(), ref[4].next_positions)


@pytest.mark.gpu
@pytest.mark.parametrize("extra", [[], ["--cuda-graph", "--ode-solver", "rk4"]])
def test_train_gde_entry_point_runs(cuda, tmp_path, monkeypatch, extra):
    """scripts/train_gde.py on synthetic episodes: two epochs, finite decreasing-or-equal best loss, a loadable checkpoint
    with the reference's state_dict keys; also with the training step replayed as a CUDA graph."""
    import importlib.util, os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("train_gde_entry", os.path.join(root, "scripts", "train_gde.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.setattr(sys, "argv", ["train_gde.py", "--synthetic", "2", "--steps-per-episode", "12", "--num-agvs", "3",
                                      "--num-pickers", "2", "--num-epochs", "2", "--batch-size", "8", "--save-dir", str(tmp_path)] + extra)
    best = mod.main()
    assert np.isfinite(best)
    runs = os.listdir(tmp_path)
    sd = torch.load(os.path.join(tmp_path, runs[0], "best_model.pth"), map_location="cpu")
    assert "ode_func.conv1.lin_l.weight" in sd and "position_decoder.bias" in sd


@pytest.mark.gpu
def test_run_gnode_entry_point_runs(cuda, monkeypatch, capsys):
    """scripts/run_gnode.py (stand-in for the reference's heterogeneous entry point): converter -> HeteroGraphODENetwork ->
    TD(0) updates on synthetic joint observations; finite losses, Q-value shapes of the 19 AGV + 9 picker warehouse."""
    import importlib.util, os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("run_gnode_entry", os.path.join(root, "scripts", "run_gnode.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.setattr(sys, "argv", ["run_gnode.py", "--steps", "3", "--batch-size", "4", "--locations", "24", "--action-size", "25",
                                      "--hidden-dim", "64"])
    mod.main()
    out = capsys.readouterr().out
    assert "agv_q (76, 25)" in out and "picker_q (36, 25)" in out and "nan" not in out.lower()
