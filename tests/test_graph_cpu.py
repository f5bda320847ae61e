"""Graph construction is integer/exact: the product's vectorised GraphConverter and the synthetic
generator must reproduce the restated reference converter bit for bit (scripts/train_gde.py:108-271)."""
import numpy as np
import pytest
import torch

import swarm_ode_b200 as S
from oracle.pyg_ref import RefBatch
from oracle.train_gde_ref import GraphConverterRef, collate_ref, extract_positions_ref, TrajectoryBatchRef


def _obs(rng, n_agv, n_pick, D=20, grid=12):
    n = n_agv + n_pick
    obs = rng.random((n, D)).astype(np.float32)
    pos = rng.integers(0, grid, size=(n, 2)).astype(np.float32)
    obs[:n_agv, 3:5] = pos[:n_agv]
    obs[n_agv:, 0:2] = pos[n_agv:]
    return obs


def test_oracle_converter_hand_built_window():
    conv = GraphConverterRef(2, 1, distance_threshold=5.0, temporal_window=3)
    D = 6
    o = np.zeros((3, D), dtype=np.float32)
    # AGVs read (y, x) at cols (3, 4); the picker at cols (0, 1)
    o[0, 3:5] = (0, 0); o[1, 3:5] = (3, 4); o[2, 0:2] = (0, 4)
    g0 = conv.build(o)
    # d(0,1) = 5 -> excluded (strict <); d(0,2) = 4, d(1,2) = 3 -> included
    assert g0.edge_index.tolist() == [[0, 2, 1, 2], [2, 0, 2, 1]]
    assert g0.is_current_agent.tolist() == [True] * 3 and g0.x.shape == (3, D)
    o2 = o.copy(); o2[0, 3:5] = (20, 20); o2[1, 3:5] = (0, 0); o2[2, 0:2] = (10, 10)     # nobody close
    g1 = conv.build(o2)
    # window k=1: previous spatial (offset 0), current spatial omitted (empty), temporal 0->3, 1->4, 2->5
    assert g1.edge_index.tolist() == [[0, 2, 1, 2, 0, 1, 2], [2, 0, 2, 1, 3, 4, 5]]
    assert g1.is_current_agent.tolist() == [False] * 3 + [True] * 3
    g2 = conv.build(o)
    # k=2: spatial(0)+0, spatial(1) (empty, still "appended"), spatial(2)+6, temporal 3->6..5->8 only
    assert g2.edge_index.tolist() == [[0, 2, 1, 2, 6, 8, 7, 8, 3, 4, 5], [2, 0, 2, 1, 8, 6, 8, 7, 6, 7, 8]]
    g3 = conv.build(o2)
    # window is full (maxlen 3): oldest snapshot dropped, k stays 2
    assert g3.x.shape == (9, D) and g3.is_current_agent.tolist() == [False] * 6 + [True] * 3
    assert g3.edge_index.tolist() == [[3, 5, 4, 5, 3, 4, 5], [5, 3, 5, 4, 6, 7, 8]]


def test_oracle_converter_ragged_observations_are_zero_padded():
    conv = GraphConverterRef(1, 1, distance_threshold=5.0)
    obs = np.empty(2, dtype=object)
    obs[0] = [0, 0, 0, 1.0, 2.0, 0, 0, 9.0]
    obs[1] = [1.0, 3.0, 0, 0]
    g = conv.build(obs)
    assert g.x.shape == (2, 8) and g.x[1, 4:].abs().sum() == 0
    assert g.edge_index.tolist() == [[0, 1], [1, 0]]        # d((1,2),(1,3)) = 1


@pytest.mark.parametrize("n_agv,n_pick,window,steps", [(12, 7, 5, 9), (3, 2, 5, 7), (19, 9, 5, 6), (1, 0, 2, 4), (4, 4, 1, 3)])
def test_product_converter_bit_exact_vs_oracle(n_agv, n_pick, window, steps):
    rng = np.random.default_rng(n_agv * 31 + n_pick)
    a = S.GraphConverter(n_agv, n_pick, distance_threshold=5.0, temporal_window=window)
    b = GraphConverterRef(n_agv, n_pick, distance_threshold=5.0, temporal_window=window)
    for _ in range(steps):
        o = _obs(rng, n_agv, n_pick)
        ga, gb = a._build_graph_from_observation(o), b.build(o)
        assert torch.equal(ga.x, gb.x)
        assert ga.edge_index.dtype == torch.int64 and torch.equal(ga.edge_index, gb.edge_index)
        assert torch.equal(ga.is_current_agent, gb.is_current_agent)
    a.reset_history(); b.reset_history()
    o = _obs(rng, n_agv, n_pick)
    assert torch.equal(a._build_graph_from_observation(o).edge_index, b.build(o).edge_index)


def test_product_converter_fractional_positions_and_default_threshold():
    rng = np.random.default_rng(5)
    a, b = S.GraphConverter(5, 3), GraphConverterRef(5, 3)          # default threshold 3.0
    for _ in range(6):
        o = _obs(rng, 5, 3)
        o[:5, 3:5] += rng.random((5, 2)).astype(np.float32)
        assert torch.equal(a._build_graph_from_observation(o).edge_index, b.build(o).edge_index)


def test_collate_matches_oracle_batching():
    rng = np.random.default_rng(3)
    convs = (S.GraphConverter(4, 2, 5.0), GraphConverterRef(4, 2, 5.0))
    items_a, items_b = [], []
    for _ in range(5):
        o = _obs(rng, 4, 2)
        ga, gb = convs[0]._build_graph_from_observation(o), convs[1].build(o)
        items_a.append(S.TrajectoryBatch(ga, S.extract_positions_from_graph(ga, 4, 2)))
        items_b.append(TrajectoryBatchRef(gb, extract_positions_ref(gb, 4, 2)))
    ca, cb = S.collate_trajectory_batches(items_a), collate_ref(items_b)
    for k in ("x", "edge_index", "batch", "ptr", "is_current_agent"):
        assert torch.equal(getattr(ca.graphs, k), getattr(cb.graphs, k)), k
    assert torch.equal(ca.next_positions, cb.next_positions)


@pytest.mark.parametrize("n_agv,n_pick,size,D", [(12, 7, "medium", 399), (19, 9, "medium", 435), (19, 9, "large", 595)])
def test_synthetic_batch_equals_converter_per_trajectory(n_agv, n_pick, size, D):
    B, W = 5, 5
    batch, nxt, pos = S.synthetic.warehouse_batch(B, n_agv, n_pick, size=size, seed=11, return_positions=True)
    n = n_agv + n_pick
    assert batch.x.shape == (B * W * n, D)
    x = batch.x.view(B, W, n, D).numpy()
    graphs = []
    for b in range(B):
        conv = GraphConverterRef(n_agv, n_pick, distance_threshold=5.0, temporal_window=W)
        for w in range(W):
            g = conv.build(x[b, w])
        graphs.append(g)
        # the row layout puts (y, x) where the converter reads it
        assert np.array_equal(conv.locations(x[b, W - 1]), pos[b, W - 1].astype(np.float32))
    ref = RefBatch.from_data_list(graphs)
    assert torch.equal(batch.x, ref.x)
    assert torch.equal(batch.edge_index, ref.edge_index)
    assert torch.equal(batch.batch, ref.batch) and torch.equal(batch.ptr, ref.ptr)
    assert torch.equal(batch.is_current_agent, ref.is_current_agent)
    # picker rows are zero-padded beyond 7*n_agv + 4*n_pick (scripts/collect_data.py:99-118)
    assert np.all(x[:, :, n_agv:, 7 * n_agv + 4 * n_pick:] == 0)
    assert nxt.shape == (B, n, 2)


def test_layout_matches_reference_formulas():
    med, large = S.synthetic.warehouse_layout("medium"), S.synthetic.warehouse_layout("large")
    assert (med.rows, med.cols, len(med.highway), len(med.shelves)) == (25, 22, 390, 160)
    assert (large.rows, large.cols, len(large.highway), len(large.shelves)) == (35, 22, 530, 240)


def test_batch_shard_partitions_graphs():
    batch, _ = S.synthetic.warehouse_batch(7, seed=2)
    parts = [batch.shard(r, 3) for r in range(3)]
    assert sum(p.num_graphs for p in parts) == 7
    assert torch.equal(torch.cat([p.x for p in parts]), batch.x)
    off, e_tot = 0, 0
    for p in parts:
        assert int(p.edge_index.min()) >= 0 and int(p.edge_index.max()) < p.x.shape[0]
        e_tot += p.edge_index.shape[1]
        assert p.ptr[0] == 0 and p.ptr[-1] == p.x.shape[0]
    assert e_tot == batch.edge_index.shape[1]
