"""HeteroGraphODENetwork (scripts/gnode.py:70-158 and scripts/run_gnode.py:67-151): golden vectors produced by the
reference's own class definitions (scripts/make_golden.py executes them with SAGEConv / HeteroConv / odeint replaced by
the oracle's restatements) against the oracle restatement (CPU, bit-exact) and the CUDA path (GPU, rel-L2 <= 1e-4)."""
import os

import numpy as np
import pytest
import torch

import swarm_ode_b200 as S
from oracle.gnode_ref import EDGE_TYPES, HeteroGraphODENetworkRef
from oracle.pyg_ref import RefHeteroData
from tests._util import FIXED_TOL, rel_l2

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hetero.npz")
DIMS = {"agv": 11, "picker": 8, "location": 5}
CASES = {"joint": dict(action_size=None, t=0.5), "typed": dict(action_size=5, t=1.0)}
KEYS = ["agv_q_values", "picker_q_values", "agv_embeddings", "picker_embeddings", "location_embeddings"]


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN)


def _data(gold, cls):
    d = cls()
    for k in ("agv", "picker", "location"):
        d[k].x = torch.from_numpy(gold[f"x/{k}"])
    for et in EDGE_TYPES:
        d[et].edge_index = torch.from_numpy(gold["edge/" + "__".join(et)])
    return d


def _model(gold, cls, tag):
    m = cls(DIMS, action_size=CASES[tag]["action_size"], hidden_dim=32, num_layers=2, ode_hidden_dim=16)
    sd = {k[len(tag) + 7:]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith(f"{tag}/param/")}
    m.load_state_dict(sd)     # the reference module's state_dict keys load unchanged
    return m


@pytest.mark.parametrize("tag", list(CASES))
def test_oracle_matches_reference_class(gold, tag):
    m = _model(gold, HeteroGraphODENetworkRef, tag)
    with torch.no_grad():
        out = m(_data(gold, RefHeteroData), integration_time=CASES[tag]["t"])
    assert list(out) == KEYS
    for k in KEYS:   # bit for bit on the generating machine; fp32 rounding elsewhere
        np.testing.assert_allclose(out[k].numpy(), gold[f"{tag}/{k}"], rtol=2e-5, atol=2e-6, err_msg=k)


def test_hetero_conv_skips_absent_relations_cpu_semantics():
    """A relation missing from edge_index_dict contributes nothing; a destination type without any relation is dropped."""
    from oracle.pyg_ref import HeteroConvRef, SAGEConvRef
    convs = {et: SAGEConvRef(4 if et[0] == et[2] else (4, 4), 4) for et in EDGE_TYPES}
    hc = HeteroConvRef(convs)
    x = {"agv": torch.randn(3, 4), "picker": torch.randn(2, 4), "location": torch.randn(5, 4)}
    out = hc(x, {("agv", "communicates", "agv"): torch.tensor([[0], [1]])})
    assert list(out) == ["agv"] and out["agv"].shape == (3, 4)


def test_cpu_tensors_raise_no_fallback(gold):
    m = _model(gold, S.HeteroGraphODENetwork, "typed")
    with pytest.raises(S.GnodeError):
        m(_data(gold, S.HeteroData))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", list(CASES))
def test_cuda_training_gradients_match_autograd_through_the_oracle(cuda, gold, tag):
    """The reference trains this Q-network with loss.backward() (scripts/gnode.py / scripts/run_gnode.py learners): every
    parameter gradient of a Q-value loss, through the action heads, the ODE solve (dopri5 for the joint form, euler for
    the typed form), both HeteroConv layers and the embeddings, against autograd through the oracle (which replays the
    GPU run's adaptive steps, see tests/test_gpu_integrate.py)."""
    m = _model(gold, S.HeteroGraphODENetwork, tag).to(cuda)
    d = _data(gold, S.HeteroData).to(cuda)
    out = m(d, integration_time=CASES[tag]["t"])
    loss = out["agv_q_values"].pow(2).mean() + 0.5 * out["picker_q_values"].pow(2).mean() + 0.1 * out["location_embeddings"].pow(2).mean()
    loss.backward()
    ref = _model(gold, HeteroGraphODENetworkRef, tag)
    if tag == "joint":
        ref.solver_options = {"imposed_dts": list(m.last_stats.dts)}
    o2 = ref(_data(gold, RefHeteroData), integration_time=CASES[tag]["t"])
    l2 = o2["agv_q_values"].pow(2).mean() + 0.5 * o2["picker_q_values"].pow(2).mean() + 0.1 * o2["location_embeddings"].pow(2).mean()
    l2.backward()
    assert abs(float(loss) - float(l2)) <= 1e-4 * abs(float(l2))
    rp = dict(ref.named_parameters())
    worst = 0.0
    for name, p in m.named_parameters():
        q = rp[name]
        if q.grad is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
            continue
        assert p.grad is not None, name
        e = rel_l2(p.grad, q.grad)
        worst = max(worst, e)
        assert e <= FIXED_TOL, (name, e)
    print(f"hetero {tag}: worst parameter-gradient rel-L2 {worst:.2e}")


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["simt", "auto"])
@pytest.mark.parametrize("tag", list(CASES))
def test_cuda_matches_reference_class(cuda, gold, tag, engine):
    prev = S.set_engine(engine)
    try:
        m = _model(gold, S.HeteroGraphODENetwork, tag).to(cuda)
        d = _data(gold, S.HeteroData).to(cuda)
        with torch.no_grad():
            out = m(d, integration_time=CASES[tag]["t"])
        assert list(out) == KEYS
        for k in KEYS:
            want = torch.from_numpy(gold[f"{tag}/{k}"])
            assert out[k].shape == want.shape
            assert rel_l2(out[k], want) <= FIXED_TOL, (k, rel_l2(out[k], want))
    finally:
        S.set_engine(prev)


@pytest.mark.gpu
def test_cuda_hetero_conv_without_some_relations(cuda):
    """Relations absent from edge_index_dict are skipped (PyG semantics): compare against the oracle."""
    from oracle.pyg_ref import HeteroConvRef, SAGEConvRef
    torch.manual_seed(0)
    H = 16
    convs = {et: S.SAGEConv(H if et[0] == et[2] else (H, H), H) for et in EDGE_TYPES}
    hc = S.HeteroConv(convs)
    ref = HeteroConvRef({et: SAGEConvRef(H if et[0] == et[2] else (H, H), H) for et in EDGE_TYPES})
    ref.load_state_dict(hc.state_dict())
    x = {"agv": torch.randn(7, H), "picker": torch.randn(3, H), "location": torch.randn(40, H)}
    eid = {("agv", "targets", "location"): torch.stack([torch.arange(7), torch.randint(0, 40, (7,))]),
           ("picker", "manages", "location"): torch.stack([torch.randint(0, 3, (40,)), torch.arange(40)]),
           ("agv", "communicates", "agv"): torch.tensor([[0, 1, 2, 2], [1, 0, 0, 6]])}
    with torch.no_grad():
        want = ref(x, eid)
        got = hc.to(cuda)({k: v.to(cuda) for k, v in x.items()}, {k: v.to(cuda) for k, v in eid.items()})
    assert set(got) == set(want) == {"location", "agv"}
    for k in want:
        assert rel_l2(got[k], want[k]) <= 1e-5, k


@pytest.mark.gpu
def test_cuda_batched_forward_over_many_hetero_graphs_from_the_converter(cuda):
    """The reference calls the network on ONE HeteroData per step (scripts/run_gnode.py:115-151); here many graphs from
    the (golden-checked) MultiAgentGraphConverter go through one forward as a disjoint union and every graph's Q-values and
    embeddings must equal its own single-graph forward through the oracle."""
    from swarm_ode_b200.hetero import MultiAgentGraphConverter
    g = np.load(os.path.join(os.path.dirname(GOLDEN), "multi_agent_converter.npz"))
    tags = ["medium/idle", "medium/one_agv_target", "medium/nothing_requested", "medium/agv_target_off_rack", "medium/reuse_first"]
    datas, refs = [], []
    for t in tags:
        racks = [tuple(int(v) for v in r) for r in g[f"{t}/racks"]]
        d = MultiAgentGraphConverter(19, 9)._build_graph_from_observation(g[f"{t}/obs"], racks)
        datas.append(d)
        r = RefHeteroData()
        for k in ("agv", "picker", "location"):
            r[k].x = d[k].x.clone()
        for et in EDGE_TYPES:
            r[et].edge_index = d[et].edge_index.clone()
        refs.append(r)
    dims = {"agv": 7, "picker": 4, "location": 2}
    torch.manual_seed(3)
    ref = HeteroGraphODENetworkRef(dims, action_size=5, hidden_dim=32, num_layers=2, ode_hidden_dim=16)
    m = S.HeteroGraphODENetwork(dims, action_size=5, hidden_dim=32, num_layers=2, ode_hidden_dim=16)
    m.load_state_dict(ref.state_dict())
    m = m.to(cuda)
    with torch.no_grad():
        out = m(S.HeteroData.from_data_list(datas).to(cuda), integration_time=1.0)
        wants = [ref(r, integration_time=1.0) for r in refs]
    for k in KEYS:
        want = torch.cat([w[k] for w in wants], dim=0)
        assert out[k].shape == want.shape, k
        assert rel_l2(out[k], want) <= FIXED_TOL, (k, rel_l2(out[k], want))
