"""Edge cases of the integrator against the oracle: batches without edges, a handful of nodes, graphs too large for
the graph-resident tiles, hidden widths the tcgen05 / chain kernels do not take (fallback engines), dense high-degree
graphs (SURVEY 8d config 4 shape, scaled down) and a geometric sweep graph (config 5 shape)."""
import pytest
import torch

import swarm_ode_b200 as S
from oracle.train_gde_ref import GraphODERef
from tests._util import FIXED_TOL, rel_l2, to_ref_batch

pytestmark = pytest.mark.gpu

GRAD_TOL = 1e-3   # see the comment in _run(): ReLU kinks on tiny batches; the solution itself is held to 1e-4


def _run(batch, D, H, solver, cuda, t, conv3_scale=0.05, seed=4):
    model = S.GraphODE(D, 3, 2, hidden_dim=H, ode_solver=solver)
    S.synthetic.init_weights(model, seed=seed, conv3_scale=conv3_scale)
    ref = GraphODERef(D, 3, 2, hidden_dim=H, ode_solver=solver)
    ref.load_state_dict(model.state_dict())
    model = model.to(cuda)
    rb = to_ref_batch(batch)
    want = ref(rb, t)
    (want["trajectories"][-1] ** 2).mean().backward()
    # float64 run of the oracle: tells a genuine mismatch from fp32 sensitivity of the oracle itself (a ReLU whose
    # pre-activation is within rounding of zero flips between two fp32 evaluation orders and moves the gradient)
    ref64 = GraphODERef(D, 3, 2, hidden_dim=H, ode_solver=solver).double()
    ref64.load_state_dict({k: v.double() for k, v in ref.state_dict().items()})
    rb64 = to_ref_batch(batch)
    rb64.x = rb64.x.double()
    (ref64(rb64, t.double())["trajectories"][-1] ** 2).mean().backward()
    r64 = dict(ref64.named_parameters())
    gb = batch.to(cuda)
    got = model(gb, t.to(cuda))
    (got["trajectories"][-1] ** 2).mean().backward()
    assert rel_l2(got["node_features"], want["node_features"]) <= FIXED_TOL
    assert rel_l2(got["trajectories"], want["trajectories"]) <= FIXED_TOL
    rp = dict(ref.named_parameters())
    worst = (0.0, "", 0.0)
    for name, p in model.named_parameters():
        denom = float(rp[name].grad.norm())
        if denom == 0.0:
            assert float(p.grad.norm()) <= 1e-6, name
        else:
            # Gradients of a ReLU network are discontinuous in the pre-activations: a unit whose pre-activation lies
            # within the 3xTF32 rounding (~3e-6) of zero may take the other branch than the oracle does.  Its forward
            # contribution is ~0 either way (the solution checks above stay at 1e-4) but its gradient differs, which
            # on these tiny batches (~100 nodes) shows up at the 1e-4..1e-3 level (scripts/dev/diag_grad.py: every
            # backward contraction switched to FFMA leaves the figure unchanged; FFMA forward removes it).
            e_ours, e_ref32 = rel_l2(p.grad, r64[name].grad), rel_l2(rp[name].grad, r64[name].grad)
            assert e_ours <= GRAD_TOL, (name, rel_l2(p.grad, rp[name].grad), e_ours, e_ref32)
            if e_ours > worst[0]:
                worst = (e_ours, name, e_ref32)
    # the margin under the relaxed tolerance, visible in the log (pytest -s / -rP)
    print(f"edge-case gradients: worst rel-L2 vs float64 oracle {worst[0]:.2e} ({worst[1]}; the fp32 oracle itself: {worst[2]:.2e}), "
          f"tolerance {GRAD_TOL:.0e}")
    from swarm_ode_b200 import _lib
    _lib.tc_check(cuda)


@pytest.mark.parametrize("fold", [True, False])
@pytest.mark.parametrize("n_nodes", [1, 3, 7, 130])
def test_no_edges_and_tiny_batches(cuda, fold, n_nodes):
    """No edges at all: every aggregate is 0 (PyG scatter-mean of an empty set); row counts below the tensor-core tile
    granularity go through the FFMA tails."""
    prev = S.set_fold(fold)
    try:
        g = torch.Generator().manual_seed(n_nodes)
        b = S.Batch(x=torch.randn(n_nodes, 37, generator=g), edge_index=torch.empty((2, 0), dtype=torch.long))
        b.batch = torch.zeros(n_nodes, dtype=torch.long)
        _run(b, 37, 64, "rk4", cuda, torch.tensor([0.0, 1.0]))
    finally:
        S.set_fold(prev)


@pytest.mark.parametrize("H", [32, 64, 128])
@pytest.mark.parametrize("solver", ["euler", "rk4"])
def test_hidden_widths(cuda, solver, H):
    """H = 32 / 128: outside the graph-resident kernel (specialised for H = 64); 2H = 256 also exceeds the tcgen05
    weight-gradient kernel's M operand, which falls back to FFMA."""
    batch, _ = S.synthetic.warehouse_batch(5, num_agvs=3, num_pickers=2, seed=9)
    _run(batch, batch.x.shape[1], H, solver, cuda, torch.tensor([0.0, 0.5, 1.0]))


def test_dense_200_node_graphs_run_as_two_block_tiles(cuda):
    """200-node dense graphs (config-4 shape, scaled): one graph per 256-row tile, processed by the chain kernels as two
    128-row blocks; in-degree 199 (every neighbour beyond the four kept in registers is walked in the CSR)."""
    batch = S.synthetic.dense_batch(3, num_agents=200, node_dim_=64, seed=2)
    assert batch.max_graph_nodes == 200
    _run(batch, 64, 64, "rk4", cuda, torch.tensor([0.0, 0.25]), conv3_scale=0.02)


def test_graphs_larger_than_256_nodes_use_the_per_op_path(cuda):
    """300-node dense graphs: no whole-graph tiling (max_graph_nodes > 256), kernel-per-op folded stages."""
    batch = S.synthetic.dense_batch(2, num_agents=300, node_dim_=64, seed=4)
    assert batch.max_graph_nodes == 300
    _run(batch, 64, 64, "rk4", cuda, torch.tensor([0.0, 0.25]), conv3_scale=0.02)


def test_geometric_sweep_graph(cuda):
    """Random geometric graphs (config-5 shape): ragged degrees, D = 128 (16-byte aligned rows -> plain TMA path)."""
    batch = S.synthetic.geometric_batch(6, num_agents=64, node_dim_=128, seed=3)
    _run(batch, 128, 64, "rk4", cuda, torch.tensor([0.0, 0.5]), conv3_scale=0.02)


def test_mixed_graph_sizes_tile_packing(cuda):
    """Graphs of different sizes (1 .. 128 nodes) packed greedily into tiles; a 128-node graph fills a tile alone."""
    g = torch.Generator().manual_seed(0)
    sizes = [1, 128, 5, 60, 70, 127, 2, 64, 64, 33]
    graphs = []
    for n in sizes:
        x = torch.randn(n, 24, generator=g)
        m = max(n * 2, 1)
        ei = torch.randint(0, n, (2, m), generator=g) if n > 1 else torch.empty((2, 0), dtype=torch.long)
        graphs.append(S.Data(x=x, edge_index=ei, is_current_agent=torch.ones(n, dtype=torch.bool)))
    batch = S.Batch.from_data_list(graphs)
    assert batch.max_graph_nodes == 128
    _run(batch, 24, 64, "rk4", cuda, torch.tensor([0.0, 1.0]))
    _run(batch, 24, 64, "midpoint", cuda, torch.tensor([0.0, 0.3, 0.6]))


def test_edge_leaving_its_graph_is_reported(cuda):
    """batch.ptr that does not match edge_index (an edge crosses two graphs) is caught by the tile check."""
    x = torch.randn(20, 16)
    ei = torch.tensor([[0, 12], [1, 3]])                   # 12 -> 3 crosses the boundary at 10
    b = S.Batch(x=x, edge_index=ei)
    b.batch = torch.cat([torch.zeros(10, dtype=torch.long), torch.ones(10, dtype=torch.long)])
    b.ptr = torch.tensor([0, 10, 20])
    b.num_graphs, b.max_graph_nodes = 2, 10
    model = S.GraphODE(16, 1, 1, hidden_dim=64, ode_solver="rk4").to(cuda)
    gb = b.to(cuda)
    # two 10-node graphs pack into ONE tile, where the edge is legal: make every graph its own tile by size
    gb.max_graph_nodes = 10
    with torch.no_grad():
        model(gb, torch.tensor([0.0, 1.0], device=cuda))
    S.graph.csr_for(gb.edge_index, 20, holder=gb).validate()      # same tile -> fine
    # now 70-node graphs: one per tile, the crossing edge must be flagged
    x = torch.randn(140, 16)
    b2 = S.Batch(x=x, edge_index=torch.tensor([[0, 100], [1, 3]]))
    b2.batch = torch.cat([torch.zeros(70, dtype=torch.long), torch.ones(70, dtype=torch.long)])
    b2.ptr = torch.tensor([0, 70, 140])
    b2.num_graphs, b2.max_graph_nodes = 2, 70
    gb2 = b2.to(cuda)
    with torch.no_grad():
        model(gb2, torch.tensor([0.0, 1.0], device=cuda))
    with pytest.raises(S.GnodeError, match="tile"):
        S.graph.csr_for(gb2.edge_index, 140, holder=gb2).validate()


def _greedy_tiles(ptr, tm=128):
    """Python statement of the packing: whole graphs, at most tm rows per tile, a new tile when the next graph does not fit."""
    starts, start = [int(ptr[0])], int(ptr[0])
    for g in range(len(ptr) - 1):
        b, e = int(ptr[g]), int(ptr[g + 1])
        if e - start > tm:
            starts.append(b)
            start = b
    return starts + [int(ptr[-1])]


@pytest.mark.gpu
@pytest.mark.parametrize("n_graphs,lo,hi", [(1, 95, 95), (7, 1, 128), (4096, 95, 95), (3000, 1, 128), (16384, 1, 40), (20000, 3, 90),
                                            (500, 129, 140), (400, 141, 144), (700, 1, 256), (20000, 100, 256)])
def test_tiles_build_matches_greedy_packing(cuda, n_graphs, lo, hi):
    """gnode_tiles_build (parallel pointer-doubling form up to 16384 graphs, sequential form above) == greedy packing."""
    g = torch.Generator().manual_seed(n_graphs)
    sizes = torch.randint(lo, hi + 1, (n_graphs,), generator=g)
    ptr = torch.cat([torch.zeros(1, dtype=torch.long), sizes.cumsum(0)])
    N = int(ptr[-1])
    ei = torch.zeros((2, 0), dtype=torch.long, device=cuda)
    csr = S.graph.CSRGraph(ei, N, graph_ptr=ptr.to(cuda), max_graph_nodes=int(sizes.max()))
    t = csr.tiles.cpu().tolist()
    m = int(sizes.max())
    assert csr.tile_rows == (128 if m <= 128 else (140 if m <= 140 else 256))
    want = _greedy_tiles(ptr.tolist(), tm=csr.tile_rows)
    assert t[0] == len(want) - 1
    assert t[1:1 + len(want)] == want


def test_understated_max_graph_nodes_is_reported(cuda):
    """A batch whose host-side ``max_graph_nodes`` is smaller than its largest graph (a stale or hand-made PyG-style
    batch): the tiles cannot be built (tiles[0] = -1), the chain kernels process nothing and flag it; validate() and
    the deferred check raise instead of handing back an uninitialised solution."""
    batch = S.synthetic.dense_batch(3, num_agents=140, node_dim_=32, seed=5)           # 140-node graphs
    model = S.GraphODE(32, 3, 2, hidden_dim=64, ode_solver="rk4").to(cuda)
    gb = batch.to(cuda)
    gb.max_graph_nodes = 100                                 # understated: 128-row tiles cannot hold a 140-node graph
    S.graph.poll_pending()
    with torch.no_grad():
        model(gb, torch.tensor([0.0, 1.0], device=cuda))
    with pytest.raises(S.GnodeError, match="max_graph_nodes"):
        S.graph.csr_for(gb.edge_index, gb.x.shape[0], holder=gb).validate()
    # the deferred path: the flag travels with the next schedule / poll
    gb2 = batch.to(cuda)
    gb2.max_graph_nodes = 100
    with torch.no_grad():
        model(gb2, torch.tensor([0.0, 1.0], device=cuda))
    torch.cuda.synchronize()
    with pytest.raises(S.GnodeError, match="max_graph_nodes"):
        S.graph.poll_pending()


def test_pending_checks_do_not_grow_when_a_batch_is_reused(cuda):
    """Loops that reuse one batch (evaluation, predict_trajectory roll-outs) keep at most one outstanding deferred check."""
    batch, _ = S.synthetic.warehouse_batch(3, num_agvs=3, num_pickers=2, seed=6)
    model = S.GraphODE(batch.x.shape[1], 3, 2, hidden_dim=64, ode_solver="euler").to(cuda)
    gb = batch.to(cuda)
    t = torch.tensor([0.0, 1.0], device=cuda)
    torch.cuda.synchronize()
    S.graph.poll_pending()
    base = len(S.graph._PENDING)
    with torch.no_grad():
        for _ in range(20):
            model(gb, t)
    assert len(S.graph._PENDING) <= base + 2
    torch.cuda.synchronize()
    S.graph.poll_pending()


def test_second_device_smoke():
    """Per-device state (max dynamic shared memory attribute, status word address) is keyed by device: the same
    process can run the tcgen05 path on cuda:1 after cuda:0."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    batch, _ = S.synthetic.warehouse_batch(4, num_agvs=12, num_pickers=7, seed=7)
    t = torch.tensor([0.0, 1.0])
    outs = []
    for d in (0, 1):
        dev = torch.device("cuda", d)
        model = S.GraphODE(batch.x.shape[1], 12, 7, hidden_dim=64, ode_solver="rk4")
        S.synthetic.init_weights(model, seed=1, conv3_scale=0.1)
        model = model.to(dev)
        out = model(batch.to(dev), t.to(dev))
        (out["trajectories"][-1] ** 2).mean().backward()
        from swarm_ode_b200 import _lib
        _lib.tc_check(dev)
        outs.append(out["node_features"].detach().cpu())
    assert torch.equal(outs[0], outs[1])      # deterministic kernels: bitwise the same on both devices
