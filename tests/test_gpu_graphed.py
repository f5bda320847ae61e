"""The CUDA-graph training step (swarm_ode_b200/graphed.py) against the eager step of dist.masked_mse_train_step."""
import copy

import pytest
import torch

import swarm_ode_b200 as S
from swarm_ode_b200.dist import masked_mse_train_step
from swarm_ode_b200.graphed import GraphedTrainStep

pytestmark = pytest.mark.gpu


@pytest.fixture
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def _setup(dev, solver, graphs=32, seed=1):
    batch, nxt = S.synthetic.warehouse_batch(graphs, seed=3)
    model = S.GraphODE(batch.x.shape[1], 12, 7, hidden_dim=64, ode_solver=solver)
    S.synthetic.init_weights(model, seed=seed, conv3_scale=0.1)
    return model.to(dev), batch, nxt


@pytest.mark.parametrize("solver", ["rk4", "euler"])
def test_graphed_step_equals_eager_step(cuda, solver):
    """Same batches (different edge counts, so the padding edges are exercised), same initial weights, Adam with
    capturable=True on both sides: the captured step must reproduce the eager losses and end at the same weights."""
    model_e, _, _ = _setup(cuda, solver)
    model_g = copy.deepcopy(model_e)
    opt_e = torch.optim.Adam(model_e.parameters(), lr=1e-3, weight_decay=1e-4, capturable=True)
    opt_g = torch.optim.Adam(model_g.parameters(), lr=1e-3, weight_decay=1e-4, capturable=True)
    t = torch.tensor([0.0, 1.0], device=cuda)
    batches = []
    for seed in (3, 4, 5, 6):
        b, nx = S.synthetic.warehouse_batch(32, seed=seed)
        batches.append((b.to(cuda), nx.to(cuda)))
    edge_counts = {int(b.edge_index.size(1)) for b, _ in batches}
    assert len(edge_counts) > 1
    gs = GraphedTrainStep(model_g, opt_g, batches[0][0], batches[0][1], t, edge_capacity=max(edge_counts) + 17)
    # preserve_state (default): the capture's warm-up steps left the weights and Adam's state where they were
    for (n, pe), (_, pg) in zip(model_e.named_parameters(), model_g.named_parameters()):
        assert torch.equal(pe, pg), n
    for b, nx in batches:
        le = masked_mse_train_step(model_e, opt_e, b, nx, t)
        lg = gs.step(b, nx)
        assert torch.allclose(le, lg, rtol=1e-6, atol=0.0), (float(le), float(lg))
    gs.check()
    assert gs.replays == len(batches)
    for (n, pe), (_, pg) in zip(model_e.named_parameters(), model_g.named_parameters()):
        assert torch.allclose(pe, pg, rtol=1e-5, atol=1e-7), n


def test_graphed_step_falls_back_for_other_shapes(cuda):
    model, batch, nxt = _setup(cuda, "rk4")
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True)
    t = torch.tensor([0.0, 1.0], device=cuda)
    gs = GraphedTrainStep(model, opt, batch.to(cuda), nxt.to(cuda), t)
    other, onxt = S.synthetic.warehouse_batch(8, seed=9)
    loss = gs.step(other.to(cuda), onxt.to(cuda))
    assert torch.isfinite(loss) and gs.replays == 0


def test_graphed_step_reports_bad_edges(cuda):
    model, batch, nxt = _setup(cuda, "rk4")
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True)
    t = torch.tensor([0.0, 1.0], device=cuda)
    b = batch.to(cuda)
    gs = GraphedTrainStep(model, opt, b, nxt.to(cuda), t)
    bad = S.Batch(x=b.x, edge_index=b.edge_index.clone())
    bad.batch, bad.is_current_agent, bad.ptr, bad.max_graph_nodes = b.batch, b.is_current_agent, b.ptr, b.max_graph_nodes
    bad.edge_index[0, 0] = b.x.shape[0] + 5
    gs.step(bad, nxt.to(cuda))
    with pytest.raises(S.GnodeError):
        gs.check()
