"""Parity at the sizes bench.py measures: every other oracle comparison uses <= 16 graphs.

* BASELINE configs[1] -- the benched 4096-graph rk4 training step (4096 tiles, ~14 tiles per persistent CTA, 8 GB save
  area): `bench.parity_check` runs the full-size step on the GPU with a loss restricted to a random 64-graph subset and
  compares solution rows, decoded positions, loss and ALL parameter gradients with the oracle on that subset (graphs
  are independent ODE systems: scripts/train_gde.py:367).
* BASELINE configs[2] -- dopri5 over 16,384 graphs of 19 AGVs + 9 pickers (D = 435, 140 nodes, two-block tiles) at FULL
  batch size: identical accepted / attempted step lists.  The error norm of dopri5 is
  the RMS over the WHOLE batch tensor, so the oracle needs the whole batch -- ~50 TFLOP of fp32 CPU work and ~60 GB of
  host memory at 16,384 distinct graphs.  The batch is therefore 128 distinct graphs x 128 replicas: the RMS norm of a
  replicated batch equals the norm of one replica (sum of squares and count both scale by 128), so the oracle on the
  128 distinct graphs takes exactly the step decisions of the full batch, while the GPU integrates all 2,293,760 nodes
  (every tile, every CTA, the full workspace) and every replica is checked against the oracle's rows.
"""
import os
import sys

import pytest
import torch

import swarm_ode_b200 as S
from oracle.train_gde_ref import GraphODERef
from tests._util import FIXED_TOL, rel_l2, to_ref_batch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _replicate(batch: S.Batch, reps: int) -> S.Batch:
    """`reps` copies of a collated batch as one disjoint union (same graphs, shifted node ids)."""
    N, G = batch.x.shape[0], int(batch.ptr.numel() - 1)
    out = S.Batch(x=batch.x.repeat(reps, 1))
    offs = (torch.arange(reps, dtype=torch.long) * N).view(reps, 1, 1)
    out.edge_index = (batch.edge_index.unsqueeze(0) + offs).permute(1, 0, 2).reshape(2, -1)
    out.batch = (batch.batch.unsqueeze(0) + (torch.arange(reps, dtype=torch.long) * G).view(reps, 1)).reshape(-1)
    out.ptr = torch.cat([torch.zeros(1, dtype=torch.long),
                         (batch.ptr[1:].unsqueeze(0) + (torch.arange(reps, dtype=torch.long) * N).view(reps, 1)).reshape(-1)])
    out.is_current_agent = batch.is_current_agent.repeat(reps)
    out.num_graphs, out.max_graph_nodes = G * reps, batch.max_graph_nodes
    return out


def test_config1_benched_step_matches_oracle_on_a_random_subset(cuda):
    sys.path.insert(0, ROOT)
    import bench
    host, nxt = S.synthetic.warehouse_batch(4096, seed=0)
    D = host.x.shape[1]
    model = S.GraphODE(D, 12, 7, hidden_dim=64, ode_solver="rk4")
    S.synthetic.init_weights(model, seed=1, conv3_scale=0.1)
    model = model.to(cuda)
    resident = S.PackedBatch(host, nxt).to(cuda)[0]
    par = bench.parity_check(model, host, nxt, resident, None, torch.tensor([0.0, 1.0], device=cuda), cuda, n_sub=64, seed=3)
    print("config1 full-size parity:", par)
    assert par["n_graphs"] == 64 and par["of_graphs"] == 4096
    assert par["max_rel_l2"] <= FIXED_TOL, par
    assert par["loss_rel_err"] <= 1e-4, par
    assert par["grad_max_rel_l2"] <= 1e-3, par
    from swarm_ode_b200 import _lib
    _lib.tc_check(cuda)
    S.graph.poll_pending()


# The workload bench.py measures (dopri5_strong): 13 accepted steps.  At this batch size the global RMS norm averages over 10^9
# elements and the field is smooth, so the controller's error ratio stays below 1 (peak 0.77) and no attempt is rejected;
# rejected attempts are covered by the small-batch regimes of tests/test_gpu_integrate.py (identical accept / reject lists
# in 14 .. 46-step runs).  A ~50-step run into the exploding regime (conv3 x -128) does reject an attempt at full size on
# both sides, but there the two fp32 evaluation orders drift apart by one step (51/52 vs 52/53 attempts): sensitivity of
# the problem, not of the implementation, so it is not asserted here.
@pytest.mark.parametrize("conv3_scale,times,distinct,need_reject", [(None, None, 128, False)])
def test_config2_full_batch_dopri5_step_lists_equal_oracle(cuda, conv3_scale, times, distinct, need_reject):
    sys.path.insert(0, ROOT)
    import bench
    conv3_scale = bench.DOPRI5_CONV3_SCALE if conv3_scale is None else conv3_scale
    times = bench.DOPRI5_TIMES if times is None else times
    reps = 16384 // distinct                                # 16,384 graphs = 2,293,760 nodes on the GPU
    small, _ = S.synthetic.warehouse_batch(distinct, num_agvs=19, num_pickers=9, seed=1000)
    D = small.x.shape[1]
    assert D == 435 and small.max_graph_nodes == 140
    ref = GraphODERef(D, 19, 9, hidden_dim=64, ode_solver="dopri5")
    S.synthetic.init_weights(ref, seed=1, conv3_scale=conv3_scale)
    model = S.GraphODE(D, 19, 9, hidden_dim=64, ode_solver="dopri5")
    model.load_state_dict(ref.state_dict())
    model = model.to(cuda)
    t = torch.tensor(times)
    with torch.no_grad():
        want = ref(to_ref_batch(small), t)
    rst = ref.last_stats
    big = _replicate(small, reps)
    assert big.num_graphs == 16384 and big.x.shape[0] == 2293760
    gb = S.PackedBatch(big).to(cuda)
    with torch.no_grad():
        got = model(gb, t.to(cuda))
    st = model.last_stats
    torch.cuda.synchronize()
    print(f"config2 full batch: GPU accepted {st.n_accepted}/{st.n_attempted} nfe {st.nfe}; oracle {rst.n_accepted}/{rst.n_attempted} "
          f"nfe {rst.nfe}; min decision margin {st.min_margin:.3e}")
    assert st.n_accepted >= 10, "the workload must exercise the controller"
    if need_reject:
        assert st.n_attempted > st.n_accepted, "this case is meant to contain a rejected attempt"
    assert st.accepted == rst.accepted, (st.accepted, rst.accepted, st.error_ratios, rst.error_ratios)
    assert (st.n_accepted, st.n_attempted, st.nfe) == (rst.n_accepted, rst.n_attempted, rst.nfe)
    for a, b in zip(st.dts, rst.dts):
        assert abs(a - b) <= 2e-2 * abs(b), (st.dts, rst.dts)
    n = small.x.shape[0]
    sol = got["node_features"]                              # [4, 16384 * 140, 435]
    for r in (0, 1, reps // 2, reps - 1):
        assert rel_l2(sol[:, r * n:(r + 1) * n], want["node_features"]) <= FIXED_TOL, r
        assert rel_l2(got["trajectories"][:, r * n:(r + 1) * n], want["trajectories"]) <= FIXED_TOL, r
    # every replica integrates the same graphs: identical rows in every tile of the batch (deterministic kernels)
    first = sol[:, :n]
    worst = 0.0
    for r in range(1, reps):
        worst = max(worst, float((sol[:, r * n:(r + 1) * n] - first).abs().max()))
    print(f"config2 full batch: replicas bitwise identical: {worst == 0.0} (max abs difference {worst:.3e})")
    assert worst <= 1e-5 * float(first.abs().max()), worst
    from swarm_ode_b200 import _lib
    _lib.tc_check(cuda)
    S.graph.poll_pending()
