"""Known-answer tests that pin the SAGEConv restatement (oracle/pyg_ref.py) without PyG."""
import torch

from oracle.pyg_ref import RefBatch, RefData, SAGEConvRef, sage_conv_ref, scatter_mean_ref


def test_scatter_mean_isolated_node_is_zero():
    rows = torch.tensor([[1.0, 2.0], [3.0, 6.0]])
    out = scatter_mean_ref(rows, torch.tensor([0, 0]), 3)
    assert torch.equal(out, torch.tensor([[2.0, 4.0], [0.0, 0.0], [0.0, 0.0]]))


def test_path_graph_hand_computed():
    # 0 -> 1 -> 2, plus 0 -> 2 ; x = one-hot-ish; identity-like weights make the answer readable
    x = torch.tensor([[1.0, 0.0], [0.0, 2.0], [4.0, 4.0]])
    ei = torch.tensor([[0, 1, 0], [1, 2, 2]])
    w_l = torch.tensor([[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]])
    w_r = torch.tensor([[2.0, 0.0], [0.0, 0.0], [0.0, -1.0]])
    b_l = torch.tensor([0.5, -0.5, 0.0])
    out = sage_conv_ref(x, ei, w_l, b_l, w_r)
    # agg: node0 = 0 (isolated as a destination), node1 = x0, node2 = (x1 + x0)/2 = [0.5, 1]
    agg = torch.tensor([[0.0, 0.0], [1.0, 0.0], [0.5, 1.0]])
    want = agg @ w_l.t() + b_l + x @ w_r.t()
    assert torch.allclose(out, want, atol=0, rtol=0)
    # isolated destination => bias_l + W_r x
    assert torch.equal(out[0], b_l + w_r @ x[0])


def test_star_graph_mean_and_multiedges():
    # leaves 1..4 -> hub 0 ; a duplicated edge counts twice in the mean (PyG scatter semantics)
    x = torch.arange(10.0).view(5, 2)
    ei = torch.tensor([[1, 2, 3, 4, 4], [0, 0, 0, 0, 0]])
    agg = scatter_mean_ref(x[ei[0]], ei[1], 5)
    assert torch.allclose(agg[0], (x[1] + x[2] + x[3] + 2 * x[4]) / 5)
    assert torch.count_nonzero(agg[1:]) == 0


def test_direction_is_source_to_target():
    x = torch.tensor([[1.0], [10.0]])
    ei = torch.tensor([[0], [1]])  # 0 -> 1 : node 1 aggregates node 0
    out = sage_conv_ref(x, ei, torch.ones(1, 1), torch.zeros(1), torch.zeros(1, 1))
    assert out[1].item() == 1.0 and out[0].item() == 0.0


def test_module_parameter_names_match_pyg():
    m = SAGEConvRef(7, 3)
    assert sorted(m.state_dict()) == ["lin_l.bias", "lin_l.weight", "lin_r.weight"]
    assert m.lin_l.weight.shape == (3, 7) and m.lin_r.weight.shape == (3, 7)


def test_batch_from_data_list_offsets():
    g0 = RefData(x=torch.zeros(3, 2), edge_index=torch.tensor([[0, 1], [1, 2]]), is_current_agent=torch.tensor([0, 0, 1], dtype=torch.bool))
    g1 = RefData(x=torch.ones(2, 2), edge_index=torch.tensor([[1], [0]]), is_current_agent=torch.tensor([1, 1], dtype=torch.bool))
    b = RefBatch.from_data_list([g0, g1])
    assert b.x.shape == (5, 2)
    assert torch.equal(b.edge_index, torch.tensor([[0, 1, 4], [1, 2, 3]]))
    assert torch.equal(b.batch, torch.tensor([0, 0, 0, 1, 1]))
    assert torch.equal(b.ptr, torch.tensor([0, 3, 5]))
    assert torch.equal(b.is_current_agent, torch.tensor([0, 0, 1, 1, 1], dtype=torch.bool))
