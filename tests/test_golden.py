"""Golden vectors produced by the REFERENCE's own code (scripts/make_golden.py imports
/root/reference/scripts/train_gde.py with torch_geometric / torchdiffeq stubbed by the oracle's restatements).

* graph_converter.npz -- reference GraphConverter / collate, bit-exact targets for the oracle's restatement, the
  product's vectorised converter and (GPU) the spatial-edge kernel;
* graph_ode.npz       -- reference GraphODE modules (closure, solver call, decoder, dict) on the restated
  third-party numerics: targets for the oracle's GraphODERef (CPU) and for the CUDA path (GPU, rel-L2 <= 1e-4).
"""
import os

import numpy as np
import pytest
import torch

import swarm_ode_b200 as S
from oracle.pyg_ref import RefBatch
from oracle.train_gde_ref import GraphConverterRef, GraphODERef, collate_ref, TrajectoryBatchRef
from tests._util import FIXED_TOL, rel_l2

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CONV_CASES = ["medium", "default_thr", "one_agent", "window2", "dense", "boundary"]
SOLVERS = ["euler", "midpoint", "rk4", "rk4_multi"]


@pytest.fixture(scope="module")
def conv_gold():
    return np.load(os.path.join(GOLDEN, "graph_converter.npz"))


@pytest.fixture(scope="module")
def ode_gold():
    return np.load(os.path.join(GOLDEN, "graph_ode.npz"))


@pytest.mark.parametrize("impl", ["oracle", "product"])
@pytest.mark.parametrize("case", CONV_CASES)
def test_graph_converter_bit_exact_vs_reference(conv_gold, case, impl):
    na, npk, D, win, steps = [int(v) for v in conv_gold[f"{case}/meta"]]
    thr = float(conv_gold[f"{case}/threshold"][0])
    if impl == "oracle":
        conv = GraphConverterRef(na, npk, distance_threshold=thr, temporal_window=win)
        build = conv.build
    else:
        conv = S.GraphConverter(na, npk, distance_threshold=thr, temporal_window=win)
        build = conv._build_graph_from_observation
    for s in range(steps):
        g = build(conv_gold[f"{case}/obs{s}"])
        assert g.edge_index.dtype == torch.int64
        assert np.array_equal(g.edge_index.numpy(), conv_gold[f"{case}/edge_index{s}"]), (case, s)
        assert np.array_equal(g.x.numpy(), conv_gold[f"{case}/x{s}"])
        assert np.array_equal(g.is_current_agent.numpy(), conv_gold[f"{case}/is_current_agent{s}"])


def test_boundary_distance_is_strict(conv_gold):
    # pairs at distance exactly 5.0 (0-1, 0-2, 0-3) are NOT connected; 1-4 (sqrt 5), 2-4 (sqrt 5), 0-4 (sqrt 20) are
    ei = conv_gold["boundary/edge_index0"]
    pairs = {tuple(p) for p in ei.T.tolist()}
    assert (0, 1) not in pairs and (0, 2) not in pairs and (0, 3) not in pairs
    assert (1, 4) in pairs and (4, 1) in pairs and (0, 4) in pairs


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_ragged_rows_and_collate_vs_reference(conv_gold, impl):
    rows = [conv_gold[f"ragged/row{i}"] for i in range(4)]
    if impl == "oracle":
        g = GraphConverterRef(2, 2, distance_threshold=5.0).build(rows)
    else:
        g = S.GraphConverter(2, 2, distance_threshold=5.0)._build_graph_from_observation(rows)
    assert np.array_equal(g.x.numpy(), conv_gold["ragged/x"])
    assert np.array_equal(g.edge_index.numpy(), conv_gold["ragged/edge_index"])
    # collate of the last four "medium" graphs
    na, npk, D, win, steps = [int(v) for v in conv_gold["medium/meta"]]
    if impl == "oracle":
        conv = GraphConverterRef(na, npk, distance_threshold=5.0, temporal_window=win)
        graphs = [conv.build(conv_gold[f"medium/obs{s}"]) for s in range(steps)]
        items = [TrajectoryBatchRef(g, torch.full((na + npk, 2), float(i))) for i, g in enumerate(graphs[-4:])]
        b = collate_ref(items)
    else:
        conv = S.GraphConverter(na, npk, distance_threshold=5.0, temporal_window=win)
        graphs = [conv._build_graph_from_observation(conv_gold[f"medium/obs{s}"]) for s in range(steps)]
        items = [S.TrajectoryBatch(g, torch.full((na + npk, 2), float(i))) for i, g in enumerate(graphs[-4:])]
        b = S.collate_trajectory_batches(items)
    assert np.array_equal(b.graphs.x.numpy(), conv_gold["medium/collate_x"])
    assert np.array_equal(b.graphs.edge_index.numpy(), conv_gold["medium/collate_edge_index"])
    assert np.array_equal(b.graphs.batch.numpy(), conv_gold["medium/collate_batch"])
    assert np.array_equal(b.graphs.is_current_agent.numpy(), conv_gold["medium/collate_mask"])
    assert np.array_equal(b.next_positions.numpy(), conv_gold["medium/collate_next"])


def _load_model(cls, gold, method, **kw):
    D = gold["x"].shape[1]
    model = cls(D, 4, 3, hidden_dim=32, ode_solver=method, **kw)
    sd = {k[len("param/"):]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith("param/")}
    model.load_state_dict(sd)           # reference state_dict keys load unchanged (checkpoint compatibility)
    return model


def _batch(gold, cls):
    b = cls(x=torch.from_numpy(gold["x"]), edge_index=torch.from_numpy(gold["edge_index"]))
    b.batch = torch.from_numpy(gold["batch"])
    b.is_current_agent = torch.from_numpy(gold["is_current_agent"])
    return b


@pytest.mark.parametrize("solver", SOLVERS + ["dopri5"])
def test_oracle_graph_ode_matches_reference_modules(ode_gold, solver):
    method = solver.split("_")[0]
    model = _load_model(GraphODERef, ode_gold, method)
    res = model(_batch(ode_gold, RefBatch), torch.from_numpy(ode_gold[f"{solver}/t"]))
    # same restated numerics underneath -> the reference's module wiring must reproduce the vectors to fp32 rounding
    # (bit for bit on the machine that generated them; another CPU's BLAS may round the last bit differently)
    np.testing.assert_allclose(res["node_features"].detach().numpy(), ode_gold[f"{solver}/node_features"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(res["trajectories"].detach().numpy(), ode_gold[f"{solver}/trajectories"], rtol=2e-5, atol=2e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["simt", "auto"])
@pytest.mark.parametrize("solver", SOLVERS)
def test_cuda_graph_ode_matches_reference_modules(cuda, ode_gold, solver, engine):
    method = solver.split("_")[0]
    prev = S.set_engine(engine)
    try:
        model = _load_model(S.GraphODE, ode_gold, method).to(cuda)
        batch = _batch(ode_gold, S.Batch).to(cuda)
        res = model(batch, torch.from_numpy(ode_gold[f"{solver}/t"]).to(cuda))
        assert rel_l2(res["node_features"], torch.from_numpy(ode_gold[f"{solver}/node_features"])) <= FIXED_TOL
        assert rel_l2(res["trajectories"], torch.from_numpy(ode_gold[f"{solver}/trajectories"])) <= FIXED_TOL
        nxt = torch.from_numpy(ode_gold["next_positions"]).to(cuda)
        loss = torch.nn.functional.mse_loss(res["trajectories"][1][batch.is_current_agent], nxt.view(-1, 2))
        loss.backward()
        assert abs(float(loss.detach()) - float(ode_gold[f"{solver}/loss"][0])) <= 1e-4 * abs(float(ode_gold[f"{solver}/loss"][0]))
        for name, p in model.named_parameters():
            want = torch.from_numpy(ode_gold[f"{solver}/grad/{name}"])
            assert rel_l2(p.grad, want) <= FIXED_TOL, (name, rel_l2(p.grad, want))
    finally:
        S.set_engine(prev)


@pytest.mark.gpu
def test_cuda_dopri5_and_predict_trajectory_match_reference_modules(cuda, ode_gold):
    model = _load_model(S.GraphODE, ode_gold, "dopri5").to(cuda)
    batch = _batch(ode_gold, S.Batch).to(cuda)
    with torch.no_grad():
        res = model(batch, torch.from_numpy(ode_gold["dopri5/t"]).to(cuda))
    assert rel_l2(res["node_features"], torch.from_numpy(ode_gold["dopri5/node_features"])) <= FIXED_TOL
    model = _load_model(S.GraphODE, ode_gold, "euler").to(cuda)
    with torch.no_grad():
        traj = model.predict_trajectory(batch, 3)
    assert traj.shape == (4, batch.x.shape[0], 2)
    assert rel_l2(traj, torch.from_numpy(ode_gold["predict_trajectory_3"])) <= FIXED_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("case", CONV_CASES)
def test_cuda_spatial_edges_bit_exact_vs_reference(cuda, conv_gold, case):
    """gnode_spatial_edges on the current snapshot of every golden step == the reference's spatial edges
    (the leading block of the step-0 edge list, and the current-snapshot block of later steps)."""
    na, npk, D, win, steps = [int(v) for v in conv_gold[f"{case}/meta"]]
    thr = float(conv_gold[f"{case}/threshold"][0])
    n = na + npk
    obs0 = conv_gold[f"{case}/obs0"]
    pos = np.concatenate([obs0[:na, 3:5], obs0[na:, 0:2]], 0).astype(np.float32)
    counts, edges = S.spatial_edges_cuda(torch.from_numpy(pos)[None].to(cuda), thr)
    got = edges[0, : int(counts[0])].t().cpu().numpy().astype(np.int64)   # rows (src, dst) -> [2, E]
    want = conv_gold[f"{case}/edge_index0"]          # window position 0: spatial edges only
    assert np.array_equal(got, want.reshape(2, -1))
