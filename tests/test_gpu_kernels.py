"""Parity of the individual CUDA operators against the CPU oracle (all through the C ABI)."""
import numpy as np
import pytest
import torch

import swarm_ode_b200 as S
from oracle.pyg_ref import sage_conv_ref, scatter_mean_ref
from oracle.train_gde_ref import GraphConverterRef, GraphODEFuncRef
from tests._util import FIXED_TOL, random_graph, rel_l2

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------- CSR
@pytest.mark.parametrize("n,e,seed", [(1, 0, 0), (7, 0, 1), (50, 200, 2), (1000, 5000, 3), (300, 30000, 4), (5000, 4, 5), (2000, 34000, 6), (40000, 90000, 7)])
def test_csr_matches_numpy(cuda, n, e, seed):
    ei = random_graph(n, e, seed)
    g = S.CSRGraph(ei.to(cuda), n)
    src, dst = ei[0].numpy(), ei[1].numpy()
    for key_arr, val_arr, rowptr, col in ((dst, src, g.rowptr, g.col), (src, dst, g.t_rowptr, g.t_col)):
        counts = np.bincount(key_arr, minlength=n)
        want_ptr = np.concatenate([[0], np.cumsum(counts)])
        assert np.array_equal(rowptr.cpu().numpy(), want_ptr)
        order = np.lexsort((val_arr, key_arr))            # by key, then ascending neighbour id
        assert np.array_equal(col.cpu().numpy()[:e], val_arr[order])


def test_csr_rejects_out_of_range_ids(cuda):
    ei = torch.tensor([[0, 1, 9], [1, 2, 0]], device=cuda)
    with pytest.raises(S.GnodeError, match="outside"):
        S.CSRGraph(ei, 5)
    ei = torch.tensor([[0, -1], [1, 0]], device=cuda)
    with pytest.raises(S.GnodeError):
        S.CSRGraph(ei, 5)


def test_cpu_tensors_are_refused_loudly():
    with pytest.raises(S.GnodeError):
        S.CSRGraph(torch.zeros(2, 3, dtype=torch.long), 4)


# ---------------------------------------------------------------- SAGE layer
@pytest.mark.parametrize("n,e,ci,co,relu", [(64, 200, 16, 8, False), (500, 1500, 399, 64, True), (500, 1500, 64, 399, False),
                                            (333, 0, 7, 5, True), (129, 4000, 64, 64, True), (40, 90, 3, 130, False)])
def test_sage_layer_forward_backward(cuda, n, e, ci, co, relu):
    torch.manual_seed(n + e)
    ei = random_graph(n, e, seed=n)
    x = torch.randn(n, ci)
    wl, wr, bl = torch.randn(co, ci) / ci ** 0.5, torch.randn(co, ci) / ci ** 0.5, torch.randn(co)
    ref_in = [t.clone().requires_grad_() for t in (x, wl, bl, wr)]
    ref = sage_conv_ref(ref_in[0], ei, ref_in[1], ref_in[2], ref_in[3])
    if relu:
        ref = torch.relu(ref)
    gout = torch.randn(n, co)
    ref.backward(gout)

    g = S.CSRGraph(ei.to(cuda), n)
    dev_in = [t.clone().to(cuda).requires_grad_() for t in (x, wl, bl, wr)]
    out = S.ops.sage_conv(dev_in[0], dev_in[1], dev_in[2], dev_in[3], g, relu)
    out.backward(gout.to(cuda))
    assert rel_l2(out, ref) <= 1e-5
    for d, r, name in zip(dev_in, ref_in, ("x", "wl", "bl", "wr")):
        assert rel_l2(d.grad, r.grad) <= 1e-5, name


def test_sage_isolated_nodes_known_answer(cuda):
    # no edges: out = b_l + W_r x exactly what the oracle's hand-checked case says
    x = torch.randn(10, 6)
    wl, wr, bl = torch.randn(4, 6), torch.randn(4, 6), torch.randn(4)
    g = S.CSRGraph(torch.empty(2, 0, dtype=torch.long, device=cuda), 10)
    out = S.ops.sage_conv(x.to(cuda), wl.to(cuda), bl.to(cuda), wr.to(cuda), g)
    assert rel_l2(out, bl + x @ wr.t()) <= 1e-6


def test_sage_module_state_dict_is_pyg_compatible(cuda):
    m = S.SAGEConv(12, 5)
    assert sorted(m.state_dict()) == ["lin_l.bias", "lin_l.weight", "lin_r.weight"]
    m = m.to(cuda)
    ei = random_graph(30, 100, 0).to(cuda)
    out = m(torch.randn(30, 12, device=cuda), ei)
    assert out.shape == (30, 5)


# ---------------------------------------------------------------- RHS (3 layers)
@pytest.mark.parametrize("n,e,D,H", [(95, 185, 399, 64), (140 * 3, 1200, 435, 64), (256, 5000, 128, 64), (17, 30, 10, 8)])
def test_rhs_forward_backward(cuda, n, e, D, H):
    torch.manual_seed(D)
    ei = random_graph(n, e, seed=D)
    ref = GraphODEFuncRef(D, H)
    S.synthetic.init_weights(ref, seed=1)
    x = torch.randn(n, D) * 3
    xr = x.clone().requires_grad_()
    out_ref = ref(None, xr, ei)
    gout = torch.randn(n, D)
    out_ref.backward(gout)

    f = S.GraphODEFunc(D, H)
    f.load_state_dict(ref.state_dict())
    f = f.to(cuda)
    xd = x.to(cuda).requires_grad_()
    out = f(torch.tensor(0.0), xd, ei.to(cuda))
    out.backward(gout.to(cuda))
    assert rel_l2(out, out_ref) <= 1e-5
    assert rel_l2(xd.grad, xr.grad) <= 1e-5
    rp = dict(ref.named_parameters())
    for name, p in f.named_parameters():
        assert rel_l2(p.grad, rp[name].grad) <= 2e-5, name


# ---------------------------------------------------------------- decoder
@pytest.mark.parametrize("shape,n_out", [((2, 190, 399), 2), ((1000, 64), 2), ((3, 7, 33), 5)])
def test_decoder(cuda, shape, n_out):
    torch.manual_seed(0)
    x = torch.randn(*shape)
    w, b = torch.randn(n_out, shape[-1]), torch.randn(n_out)
    ri = [t.clone().requires_grad_() for t in (x, w, b)]
    ref = torch.nn.functional.linear(*ri)
    gout = torch.randn_like(ref)
    gout[..., : shape[-2] // 2, :] = 0          # rows with zero cotangent take the skip path
    ref.backward(gout)
    di = [t.clone().to(cuda).requires_grad_() for t in (x, w, b)]
    out = S.ops.decode_positions(*di)
    out.backward(gout.to(cuda))
    assert rel_l2(out, ref) <= 1e-5
    for d, r in zip(di, ri):
        assert rel_l2(d.grad, r.grad) <= 1e-5


# ---------------------------------------------------------------- graph construction (bit-exact)
@pytest.mark.parametrize("n_snap,n", [(1, 2), (64, 19), (200, 28), (3, 70)])
def test_spatial_edges_bit_exact(cuda, n_snap, n):
    rng = np.random.default_rng(n)
    pos = rng.integers(0, 25, size=(n_snap, n, 2)).astype(np.float32)
    pos[0, 0] = (0, 0); pos[0, 1] = (3, 4)          # d == 5 exactly: strict '<' must exclude it
    if n > 2:
        pos[0, 2] = (0, 4)                           # d(0,2) = 4 < 5, d(1,2) = 3
    counts, edges = S.spatial_edges_cuda(torch.from_numpy(pos).to(cuda), 5.0)
    conv = GraphConverterRef(n, 0, distance_threshold=5.0)
    for s in range(n_snap):
        want = conv.spatial_edges(pos[s]).numpy()
        c = int(counts[s])
        assert c == want.shape[1]
        got = edges[s, :c].cpu().numpy().T
        assert np.array_equal(got, want), s
    got0 = edges[0, :int(counts[0])].cpu().numpy()
    assert not any((a, b) == (0, 1) for a, b in got0.tolist())


def test_spatial_edges_non_integer_positions_bit_exact(cuda):
    rng = np.random.default_rng(7)
    pos = (rng.random((50, 19, 2)) * 12).astype(np.float32)
    counts, edges = S.spatial_edges_cuda(torch.from_numpy(pos).to(cuda), 5.0)
    conv = GraphConverterRef(19, 0, distance_threshold=5.0)
    for s in range(50):
        want = conv.spatial_edges(pos[s]).numpy()
        assert int(counts[s]) == want.shape[1]
        assert np.array_equal(edges[s, :int(counts[s])].cpu().numpy().T, want)


def test_deferred_csr_validation_flags_bad_edges(cuda):
    """GraphODE builds its CSR without a host synchronisation; an out-of-range edge is skipped on the device and
    reported by validate() (or by the non-blocking poll of a later call)."""
    from swarm_ode_b200.graph import CSRGraph
    ei = torch.tensor([[0, 1, 7], [1, 2, 0]], device=cuda)      # node 7 does not exist (n = 3)
    with pytest.raises(S.GnodeError, match="outside"):
        CSRGraph(ei, 3)                                          # synchronous validation (default)
    g = CSRGraph(ei, 3, validate="deferred")
    assert g.rowptr.tolist() == [0, 0, 1, 2]                    # the two valid edges survive
    with pytest.raises(S.GnodeError, match="deferred"):
        g.validate()
    ok = CSRGraph(ei[:, :2], 3, validate="deferred")
    ok.validate()


def test_time_grid_cache_survives_address_reuse(cuda):
    """The host copy of a CUDA time grid is cached by storage address; a freed grid's address must not serve a new one."""
    from swarm_ode_b200 import ops
    seen = []
    for k in range(20):
        t = torch.tensor([0.0, 0.1 * (k + 1), 1.0 + k], device=cuda)
        seen.append(ops.time_grid_to_host(t))
        del t
    assert seen == [(0.0, float(torch.tensor(0.1 * (k + 1), dtype=torch.float32)), 1.0 + k) for k in range(20)]


@pytest.mark.parametrize("T,n_agv,n_pick,thr,W", [(1, 3, 2, 3.0, 5), (3, 12, 7, 5.0, 5), (23, 12, 7, 5.0, 5), (40, 19, 9, 3.0, 5),
                                                    (12, 4, 0, 100.0, 3), (9, 1, 0, 5.0, 5), (17, 6, 5, 0.5, 1)])
def test_build_episode_batch_equals_sequential_graph_converter(cuda, T, n_agv, n_pick, thr, W):
    """all window graphs of an episode built on the device == GraphConverter step by step (scripts/
    train_gde.py:116-184, itself golden-checked against the reference's code) + Batch.from_data_list, bit for bit."""
    g = torch.Generator().manual_seed(T * 131 + n_agv)
    n, D = n_agv + n_pick, 11
    obs = torch.randn(T, n, D, generator=g)
    obs[:, :, :5] = torch.randint(0, 12, (T, n, 5), generator=g).float()       # integer grid coordinates, many ties
    conv = S.GraphConverter(n_agv, n_pick, distance_threshold=thr, temporal_window=W)
    want = S.Batch.from_data_list([conv._build_graph_from_observation(obs[t].numpy()) for t in range(T)])
    got = S.build_episode_batch(obs.to(cuda), n_agv, n_pick, distance_threshold=thr, temporal_window=W)
    assert torch.equal(got.x.cpu(), want.x)
    assert got.edge_index.dtype == torch.int64 and torch.equal(got.edge_index.cpu(), want.edge_index)
    assert torch.equal(got.batch.cpu(), want.batch) and torch.equal(got.ptr.cpu(), want.ptr)
    assert torch.equal(got.is_current_agent.cpu(), want.is_current_agent)
    assert got.num_graphs == want.num_graphs and got.max_graph_nodes == want.max_graph_nodes
