"""Trajectory parity of the integrators with the torch/torchdiffeq restatement (north_star gates):
fp32 relative L2 <= 1e-4 at fixed step, identical accepted-step counts for dopri5."""
import pytest
import torch

import swarm_ode_b200 as S
from oracle.train_gde_ref import GraphODERef, train_step_loss_ref
from tests._util import FIXED_TOL, rel_l2, to_ref_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["folded", "direct"])
def fold_mode(request, cuda):
    """Every test runs with the folded RK stages (default: D-wide contractions once per step, csrc/fold.cu) and with
    the direct evaluation of the field at every stage (the straightforward anchor)."""
    prev = S.set_fold(request.param == "folded")
    yield request.param
    S.set_fold(prev)


DOPRI5_GRAD_TOL = 1e-3   # against the float64 oracle when a ReLU branch flip separates the two fp32 runs (tiny batches)


def _models(D, solver, cuda, conv3_scale=0.1, seed=1):
    model = S.GraphODE(D, 12, 7, hidden_dim=64, ode_solver=solver)
    S.synthetic.init_weights(model, seed=seed, conv3_scale=conv3_scale)
    ref = GraphODERef(D, 12, 7, hidden_dim=64, ode_solver=solver)
    ref.load_state_dict(model.state_dict())
    return model.to(cuda), ref


def test_state_dict_keys_match_reference_layout():
    model = S.GraphODE(399, 3, 2)
    keys = sorted(model.state_dict())
    want = sorted([f"ode_func.conv{i}.lin_l.weight" for i in (1, 2, 3)] + [f"ode_func.conv{i}.lin_l.bias" for i in (1, 2, 3)] +
                  [f"ode_func.conv{i}.lin_r.weight" for i in (1, 2, 3)] + ["position_decoder.weight", "position_decoder.bias"])
    assert keys == want
    assert model.ode_func.conv1.lin_l.weight.shape == (64, 399)
    assert model.ode_func.conv3.lin_r.weight.shape == (399, 64)


@pytest.mark.parametrize("solver", ["euler", "midpoint", "rk4"])
def test_fixed_step_forward_dict(cuda, solver):
    batch, _ = S.synthetic.warehouse_batch(16, seed=0)
    D = batch.x.shape[1]
    model, ref = _models(D, solver, cuda)
    t = torch.tensor([0.0, 1.0])
    want = ref(to_ref_batch(batch), t)
    got = model(batch.to(cuda), t.to(cuda))
    assert set(got) == {"trajectories", "node_features", "batch"}
    assert got["node_features"].shape == want["node_features"].shape == (2, batch.x.shape[0], D)
    assert got["trajectories"].shape == (2, batch.x.shape[0], 2)
    assert torch.equal(got["node_features"][0].cpu(), batch.x.cpu())
    assert rel_l2(got["node_features"], want["node_features"]) <= FIXED_TOL
    assert rel_l2(got["trajectories"], want["trajectories"]) <= FIXED_TOL
    assert torch.equal(got["batch"].cpu(), batch.batch.cpu())


@pytest.mark.parametrize("solver,steps", [("euler", 5), ("rk4", 5), ("midpoint", 3)])
def test_predict_trajectory_multi_step(cuda, solver, steps):
    batch, _ = S.synthetic.warehouse_batch(4, seed=3)
    D = batch.x.shape[1]
    model, ref = _models(D, solver, cuda, conv3_scale=0.02)
    want = ref.predict_trajectory(to_ref_batch(batch), steps)
    got = model.predict_trajectory(batch.to(cuda), steps)
    assert got.shape == (steps + 1, batch.x.shape[0], 2)
    assert rel_l2(got, want) <= FIXED_TOL


@pytest.mark.parametrize("n_out", [1, 2, 3, 4, 6])
@pytest.mark.parametrize("fold", [True, False])
def test_fused_decoder_equals_decoder_over_the_solution(cuda, n_out, fold):
    """gnode_integrate_fixed_decoded (GraphODE.forward, scripts/train_gde.py:78-94): time point j + 1 decoded from time
    point j and the step's 2H-wide stage combination (folded integrator, n_out <= 4) must equal position_decoder applied
    to the D-wide solution; n_out = 6 and the direct integrator take the plain decoder inside the same call."""
    batch, _ = S.synthetic.warehouse_batch(6, seed=9)
    D = batch.x.shape[1]
    f = S.GraphODEFunc(D, 64)
    S.synthetic.init_weights(f, seed=2, conv3_scale=0.05)
    f = f.to(cuda)
    g = torch.Generator().manual_seed(5)
    dec_w = (torch.randn(n_out, D, generator=g) / D ** 0.5).to(cuda)
    dec_b = torch.randn(n_out, generator=g).to(cuda)
    gb = batch.to(cuda)
    graph = S.csr_for(gb.edge_index, gb.x.shape[0], graph_ptr=gb.ptr, max_graph_nodes=95)
    t = [0.0, 0.5, 1.0, 1.75, 2.0, 3.0]
    prev = S._lib.set_fold(fold)
    try:
        with torch.no_grad():
            sol, traj = S.ops.integrate_fixed_decode(gb.x, graph, f.param_list(), t, "rk4", dec_w, dec_b)
            sol2 = S.ops.integrate_fixed(gb.x, graph, f.param_list(), t, "rk4")
    finally:
        S._lib.set_fold(prev)
    S._lib.tc_check(cuda)
    assert torch.equal(sol[0], gb.x) and torch.equal(sol, sol2)
    want = torch.nn.functional.linear(sol.double(), dec_w.double(), dec_b.double())
    err = rel_l2(traj, want)
    print(f"fused decoder n_out={n_out} fold={fold}: rel-L2 vs float64 decoder over the solution {err:.2e}")
    assert traj.shape == (len(t), gb.x.shape[0], n_out)
    assert err <= 2e-6


@pytest.mark.parametrize("solver,T", [("euler", 2), ("rk4", 2), ("midpoint", 2), ("rk4", 4)])
def test_train_step_gradients(cuda, solver, T):
    batch, nxt = S.synthetic.warehouse_batch(8, seed=1)
    D = batch.x.shape[1]
    model, ref = _models(D, solver, cuda, conv3_scale=0.05)
    t = torch.linspace(0, 1, T)
    rb = to_ref_batch(batch)
    loss_ref = train_step_loss_ref(ref, rb, nxt, t)
    loss_ref.backward()
    gb = batch.to(cuda)
    pred = model(gb, t.to(cuda))["trajectories"][1]
    loss = torch.nn.functional.mse_loss(pred[gb.is_current_agent], nxt.to(cuda).view(-1, 2))
    loss.backward()
    assert abs(float(loss) - float(loss_ref)) <= 1e-4 * abs(float(loss_ref))
    rp = dict(ref.named_parameters())
    for name, p in model.named_parameters():
        assert p.grad is not None, name
        assert rel_l2(p.grad, rp[name].grad) <= FIXED_TOL, name


@pytest.mark.parametrize("solver,T", [("rk4", 2), ("midpoint", 4), ("euler", 3)])
def test_saved_intermediates_equal_recompute(cuda, solver, T):
    """The backward that reads the forward's save area must give the gradients of the recomputing backward."""
    from swarm_ode_b200 import ops
    batch, nxt = S.synthetic.warehouse_batch(6, seed=8)
    D = batch.x.shape[1]
    model, _ = _models(D, solver, cuda, conv3_scale=0.05)
    gb = batch.to(cuda)
    t = torch.linspace(0, 1, T).to(cuda)
    grads = {}
    for mode, frac in (("save", 0.5), ("recompute", 0.0)):
        old = ops.SAVE_FRACTION
        ops.SAVE_FRACTION = frac
        try:
            model.zero_grad(set_to_none=True)
            gb.x.grad = None
            gb.x.requires_grad_()
            out = model(gb, t)
            loss = (out["trajectories"][-1] ** 2).mean() + out["node_features"][1].sum() * 1e-3
            loss.backward()
            grads[mode] = [p.grad.clone() for p in model.parameters()] + [gb.x.grad.clone()]
        finally:
            ops.SAVE_FRACTION = old
    for a, b in zip(grads["save"], grads["recompute"]):
        assert torch.equal(a, b)


def test_backward_wrt_initial_state_and_all_time_points(cuda):
    batch, _ = S.synthetic.warehouse_batch(3, seed=5)
    D = batch.x.shape[1]
    model, ref = _models(D, "rk4", cuda, conv3_scale=0.05)
    t = torch.tensor([0.0, 0.4, 1.0])
    rb = to_ref_batch(batch)
    rb.x.requires_grad_()
    out_ref = ref(rb, t)
    w = torch.randn_like(out_ref["node_features"])
    (out_ref["node_features"] * w).sum().backward()
    gb = batch.to(cuda)
    gb.x.requires_grad_()
    out = model(gb, t.to(cuda))
    (out["node_features"] * w.to(cuda)).sum().backward()
    assert rel_l2(gb.x.grad, rb.x.grad) <= FIXED_TOL


def test_masked_mse_train_step_matches_reference_loop(cuda):
    """Two optimiser steps of the reference loop (Adam, clip 1.0) stay on the oracle's trajectory."""
    batch, nxt = S.synthetic.warehouse_batch(8, seed=2)
    D = batch.x.shape[1]
    model, ref = _models(D, "euler", cuda, conv3_scale=0.05)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)
    opt_ref = torch.optim.Adam(ref.parameters(), lr=1e-3, weight_decay=1e-4)
    gb, rb = batch.to(cuda), None
    from swarm_ode_b200.dist import masked_mse_train_step
    b2, _ = S.synthetic.warehouse_batch(8, seed=2)
    rb = to_ref_batch(b2)
    for _ in range(2):
        loss = masked_mse_train_step(model, opt, gb, nxt.to(cuda))
        opt_ref.zero_grad()
        loss_ref = train_step_loss_ref(ref, rb, nxt)
        loss_ref.backward()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm=1.0)
        opt_ref.step()
        assert abs(float(loss) - float(loss_ref)) <= 1e-4 * abs(float(loss_ref))
    rp = dict(ref.named_parameters())
    for name, p in model.named_parameters():
        assert rel_l2(p, rp[name]) <= 1e-4, name


# ---------------------------------------------------------------- dopri5
@pytest.mark.parametrize("B,t_points,conv3_scale", [(8, [0.0, 1.0], 0.1), (4, [0.0, 0.25, 0.5, 1.0, 2.0], 0.1), (6, [0.0, 1.0], 0.3)])
def test_dopri5_counts_and_values(cuda, B, t_points, conv3_scale):
    batch, _ = S.synthetic.warehouse_batch(B, num_agvs=19, num_pickers=9, seed=4)
    D = batch.x.shape[1]
    assert D == 435
    model, ref = _models(D, "dopri5", cuda, conv3_scale=conv3_scale)
    t = torch.tensor(t_points)
    with torch.no_grad():
        want = ref(to_ref_batch(batch), t)
        got = model(batch.to(cuda), t.to(cuda))
    st, rst = model.last_stats, ref.last_stats
    print(f"dopri5: accepted {st.n_accepted}/{rst.n_accepted} attempted {st.n_attempted}/{rst.n_attempted} "
          f"first_step {st.first_step:.6g}/{rst.first_step:.6g} min|ratio-1| {st.min_margin:.3g}")
    assert st.n_accepted == rst.n_accepted
    assert st.n_attempted == rst.n_attempted
    assert st.accepted == rst.accepted
    assert st.nfe == rst.nfe
    assert abs(st.first_step - rst.first_step) <= 1e-5 * rst.first_step
    for a, b in zip(st.error_ratios, rst.error_ratios):
        # tiny ratios are dominated by fp32 cancellation in the error estimate; decisions live near 1
        assert abs(a - b) <= 2e-2 * abs(b) + 1e-6
    assert rel_l2(got["node_features"], want["node_features"]) <= FIXED_TOL
    assert rel_l2(got["trajectories"], want["trajectories"]) <= FIXED_TOL


@pytest.mark.parametrize("scale,rtol,atol,t_points", [(3.0, 1e-5, 1e-6, [0.0, 5.0]), (6.0, 1e-4, 1e-5, [0.0, 1.0, 3.0]),
                                                       (10.0, 1e-3, 1e-4, [0.0, 0.5, 1.0, 1.5, 2.0, 3.0]),
                                                       (3.0, 1e-6, 1e-7, [0.0, 2.0])])
def test_dopri5_many_steps_identical_decisions(cuda, scale, rtol, atol, t_points):
    """Harder regimes (14-46 steps): every accept/reject decision, the step sizes and the dense outputs
    must follow the torchdiffeq restatement."""
    from oracle.torchdiffeq_ref import SolverStats, odeint_ref
    from oracle.train_gde_ref import GraphODEFuncRef

    batch, _ = S.synthetic.warehouse_batch(2, num_agvs=19, num_pickers=9, seed=4)
    D = batch.x.shape[1]
    fref = GraphODEFuncRef(D, 64)
    S.synthetic.init_weights(fref, seed=1, conv3_scale=scale)
    f = S.GraphODEFunc(D, 64)
    f.load_state_dict(fref.state_dict())
    f = f.to(cuda)
    rst = SolverStats()
    t = torch.tensor(t_points)
    with torch.no_grad():
        want = odeint_ref(lambda tt, x: fref(tt, x, batch.edge_index), batch.x, t, rtol=rtol, atol=atol, method="dopri5",
                          stats=rst)
        gb = batch.to(cuda)
        g = S.csr_for(gb.edge_index, gb.x.shape[0])
        got, st = S.ops.integrate_dopri5(gb.x, g, f.param_list(), t_points, rtol, atol)
    print(f"dopri5[{scale},{rtol}]: accepted {st.n_accepted}/{rst.n_accepted} attempted {st.n_attempted}/{rst.n_attempted} "
          f"min|ratio-1| {st.min_margin:.3g}")
    assert rst.n_accepted >= 10
    assert st.accepted == rst.accepted and st.n_attempted == rst.n_attempted and st.nfe == rst.nfe
    for a, b in zip(st.dts, rst.dts):
        assert abs(a - b) <= 1e-2 * abs(b)
    assert rel_l2(got, want) <= FIXED_TOL


@pytest.mark.parametrize("tiled", [True, False])
@pytest.mark.parametrize("scale,rtol,atol,t_points", [(6.0, 1e-4, 1e-5, [0.0, 1.0, 3.0]),
                                                       (10.0, 1e-3, 1e-4, [0.0, 0.5, 1.0, 1.5, 2.0, 3.0])])
def test_dopri5_folded_fsal_matches_the_step_by_step_form(cuda, tiled, scale, rtol, atol, t_points):
    """gnode_set_dopri5_fsal: Z_0 handed from stage 6 of an accepted step to the next attempt and dense output as one
    projection (default) against Z_0 contracted from y on every attempt and torchdiffeq's quartic evaluated D-wide.
    Same decisions, same values to fp32 rounding, and both against the oracle (scripts/train_gde.py:78-85 with
    ode_solver='dopri5').  tiled = graph-resident stage kernel (whole-graph tiles from `ptr`), else kernel per op."""
    from oracle.torchdiffeq_ref import SolverStats, odeint_ref
    from oracle.train_gde_ref import GraphODEFuncRef

    batch, _ = S.synthetic.warehouse_batch(3, num_agvs=19, num_pickers=9, seed=4)
    D = batch.x.shape[1]
    fref = GraphODEFuncRef(D, 64)
    S.synthetic.init_weights(fref, seed=1, conv3_scale=scale)
    f = S.GraphODEFunc(D, 64)
    f.load_state_dict(fref.state_dict())
    f = f.to(cuda)
    rst = SolverStats()
    t = torch.tensor(t_points)
    with torch.no_grad():
        want = odeint_ref(lambda tt, x: fref(tt, x, batch.edge_index), batch.x, t, rtol=rtol, atol=atol, method="dopri5",
                          stats=rst)
        gb = batch.to(cuda)
        if tiled:
            g = S.csr_for(gb.edge_index, gb.x.shape[0], graph_ptr=gb.ptr, max_graph_nodes=140)
        else:
            g = S.csr_for(gb.edge_index, gb.x.shape[0])
        prev = S._lib.set_dopri5_fsal(True)
        try:
            assert prev is True                      # the default
            got_on, st_on = S.ops.integrate_dopri5(gb.x, g, f.param_list(), t_points, rtol, atol)
            S._lib.set_dopri5_fsal(False)
            got_off, st_off = S.ops.integrate_dopri5(gb.x, g, f.param_list(), t_points, rtol, atol)
        finally:
            S._lib.set_dopri5_fsal(prev)
    S._lib.tc_check(cuda)
    print(f"dopri5 fsal[{'tiles' if tiled else 'per-op'},{scale},{rtol}]: accepted {st_on.n_accepted}/{st_off.n_accepted}/{rst.n_accepted} "
          f"attempted {st_on.n_attempted}/{st_off.n_attempted}/{rst.n_attempted} on-vs-off {rel_l2(got_on, got_off):.2e} "
          f"on-vs-oracle {rel_l2(got_on, want):.2e} off-vs-oracle {rel_l2(got_off, want):.2e}")
    assert rst.n_accepted >= 10
    for st in (st_on, st_off):
        assert st.accepted == rst.accepted and st.n_attempted == rst.n_attempted and st.nfe == rst.nfe
        for a, b in zip(st.dts, rst.dts):
            assert abs(a - b) <= 1e-2 * abs(b)
    assert rel_l2(got_on, got_off) <= 5e-5      # the D-wide quartic of the off form cancels 32 |y| terms in fp32
    assert rel_l2(got_on, want) <= FIXED_TOL
    assert rel_l2(got_off, want) <= FIXED_TOL


def test_odeint_entry_point_signature(cuda):
    """`odeint(func, y0, t, rtol=, atol=, method=)` with the reference's closure replaced by .bind()."""
    from oracle.torchdiffeq_ref import odeint_ref
    from oracle.train_gde_ref import GraphODEFuncRef

    batch, _ = S.synthetic.warehouse_batch(2, seed=6)
    D = batch.x.shape[1]
    fref = GraphODEFuncRef(D, 64)
    S.synthetic.init_weights(fref, seed=1, conv3_scale=0.1)
    f = S.GraphODEFunc(D, 64)
    f.load_state_dict(fref.state_dict())
    f = f.to(cuda)
    t = torch.tensor([0.0, 0.3, 1.0])
    gb, _ = S.synthetic.warehouse_batch(2, seed=6)
    gb = gb.to(cuda)
    for method in ("euler", "midpoint", "rk4", "dopri5", None):
        with torch.no_grad():
            want = odeint_ref(lambda tt, x: fref(tt, x, batch.edge_index), batch.x, t, rtol=1e-3, atol=1e-4, method=method)
            got = S.odeint(f.bind(gb.edge_index), gb.x, t.to(cuda), rtol=1e-3, atol=1e-4, method=method)
        assert got.shape == (3,) + tuple(batch.x.shape)
        assert rel_l2(got, want) <= FIXED_TOL, method


@pytest.mark.parametrize("t_points,conv3_scale", [((0.0, 1.0), 0.1), ((0.0, 0.4, 1.0), 0.1), ((0.0, 0.25, 0.5, 2.0), 0.3)])
def test_dopri5_backward_matches_autograd_through_the_oracle(cuda, t_points, conv3_scale):
    """loss.backward() through the adaptive solve (scripts/train_gde.py:493 with ode_solver='dopri5'): gradients of the
    parameters and of the initial state against autograd through the restated torchdiffeq solver.  Outputs fall inside
    accepted steps (dense output), several outputs can share a step, and later steps feed on earlier ones.

    The oracle replays the GPU run's attempted step sizes: the first, tiny step has a rounding-level error estimate
    (ratio ~1e-5), so the size of the second step is only reproducible to ~0.5 % between implementations, and the
    gradient is more sensitive to the discretisation than the solution (identical free-running decisions are the subject
    of test_dopri5_counts_and_values / test_dopri5_many_steps_identical_decisions)."""
    batch, nxt = S.synthetic.warehouse_batch(6, seed=5)
    D = batch.x.shape[1]
    model, ref = _models(D, "dopri5", cuda, conv3_scale=conv3_scale)
    t = torch.tensor(t_points)
    w_out = torch.linspace(1.0, 2.0, len(t_points)).view(-1, 1, 1)
    gb = batch.to(cuda)
    gb.x = gb.x.clone().requires_grad_(True)
    out = model(gb, t.to(cuda))
    st = model.last_stats
    loss = (out["trajectories"] * w_out.to(cuda)).pow(2).mean() + 1e-3 * (out["node_features"] * w_out.to(cuda)).pow(2).mean()
    loss.backward()
    rb = to_ref_batch(batch)
    rb.x = rb.x.clone().requires_grad_(True)
    ref.solver_options = {"imposed_dts": list(st.dts)}
    out_ref = ref(rb, t)
    rst = ref.last_stats
    assert rst.accepted == st.accepted, (rst.accepted, st.accepted, rst.error_ratios, st.error_ratios)
    loss_ref = (out_ref["trajectories"] * w_out).pow(2).mean() + 1e-3 * (out_ref["node_features"] * w_out).pow(2).mean()
    loss_ref.backward()
    # float64 run of the oracle on the same steps: tells a genuine mismatch from a ReLU whose pre-activation is within
    # fp32 rounding of zero and takes different branches in two fp32 evaluation orders (tests/test_gpu_edge_cases.py)
    ref64 = GraphODERef(D, 12, 7, hidden_dim=64, ode_solver="dopri5").double()
    ref64.load_state_dict({k: v.double() for k, v in ref.state_dict().items()})
    ref64.solver_options = {"imposed_dts": list(st.dts)}
    rb64 = to_ref_batch(batch)
    rb64.x = rb64.x.double().requires_grad_(True)
    o64 = ref64(rb64, t.double())
    ((o64["trajectories"] * w_out.double()).pow(2).mean() + 1e-3 * (o64["node_features"] * w_out.double()).pow(2).mean()).backward()
    print(f"dopri5 bwd {t_points}: accepted {st.n_accepted}/{st.n_attempted} loss {float(loss):.6g} vs {float(loss_ref):.6g}")
    assert rel_l2(out["node_features"], out_ref["node_features"]) <= FIXED_TOL
    assert abs(float(loss) - float(loss_ref)) <= 1e-4 * abs(float(loss_ref))
    rp, r64 = dict(ref.named_parameters()), dict(ref64.named_parameters())
    worst = 0.0
    for name, p in model.named_parameters():
        assert p.grad is not None, name
        e32 = rel_l2(p.grad, rp[name].grad)
        worst = max(worst, e32)
        if e32 > FIXED_TOL:
            e_ours, e_ref32 = rel_l2(p.grad, r64[name].grad), rel_l2(rp[name].grad, r64[name].grad)
            assert e_ours <= DOPRI5_GRAD_TOL, (name, e32, e_ours, e_ref32)
    print(f"dopri5 backward: worst parameter-gradient rel-L2 vs the fp32 oracle {worst:.2e} (gate {FIXED_TOL:.0e}, "
          f"{DOPRI5_GRAD_TOL:.0e} against float64 where a ReLU branch flip separates the two fp32 runs)")
    e32 = rel_l2(gb.x.grad, rb.x.grad)
    assert e32 <= FIXED_TOL or rel_l2(gb.x.grad, rb64.x.grad) <= DOPRI5_GRAD_TOL, (e32, rel_l2(gb.x.grad, rb64.x.grad))


def test_dopri5_train_step(cuda):
    """The reference's training step with the adaptive solver: masked MSE on the decoded positions at t = 1."""
    batch, nxt = S.synthetic.warehouse_batch(8, seed=2)
    D = batch.x.shape[1]
    model, ref = _models(D, "dopri5", cuda, conv3_scale=0.1)
    t = torch.tensor([0.0, 1.0])
    gb = batch.to(cuda)
    pred = model(gb, t.to(cuda))["trajectories"][1]
    loss = torch.nn.functional.mse_loss(pred[gb.is_current_agent], nxt.to(cuda).view(-1, 2))
    loss.backward()
    ref.solver_options = {"imposed_dts": list(model.last_stats.dts)}     # same discretisation (see the test above)
    loss_ref = train_step_loss_ref(ref, to_ref_batch(batch), nxt, t)
    loss_ref.backward()
    assert abs(float(loss) - float(loss_ref)) <= 1e-4 * abs(float(loss_ref))
    rp = dict(ref.named_parameters())
    for name, p in model.named_parameters():
        assert rel_l2(p.grad, rp[name].grad) <= DOPRI5_GRAD_TOL, (name, rel_l2(p.grad, rp[name].grad))


@pytest.mark.parametrize("device_hook", [False, True])
def test_dopri5_two_rank_lockstep_reproduces_unsharded_decisions(cuda, device_hook):
    """Two data-parallel 'ranks' run as two host threads (own CUDA stream each) on one GPU; their
    error-norm hooks rendezvous on the host and exchange (sum of squares, count) exactly as the NCCL /
    gloo all-reduce does.  Every rank must then take the decisions of the unsharded batch.
    device_hook: the per-attempt exchange runs on the device copy of the sum (gnode_set_dopri5_device_allreduce, what
    dist.dopri5_norm_allreduce does with NCCL); the host hook then only sums the element count, once per solve."""
    import threading

    batch, _ = S.synthetic.warehouse_batch(6, num_agvs=19, num_pickers=9, seed=9)
    D = batch.x.shape[1]
    model, _ = _models(D, "dopri5", cuda)
    params = [p.detach() for p in model.ode_func.param_list()]
    t = [0.0, 0.5, 1.0]
    full_batch = batch.to(cuda)
    g_full = S.csr_for(full_batch.edge_index, full_batch.x.shape[0])
    full, full_stats = S.ops.integrate_dopri5(full_batch.x, g_full, params, t, 1e-3, 1e-4)

    b2, _ = S.synthetic.warehouse_batch(6, num_agvs=19, num_pickers=9, seed=9)
    shards = [b2.shard(r, 2).to(cuda) for r in range(2)]
    barrier = threading.Barrier(2)
    slots = [None, None]
    results, errors = [None, None], []

    dev_slots = [None, None]
    calls = {"host": [0, 0], "dev": [0, 0]}

    def hook_for(rank):
        def hook(s, c):
            calls["host"][rank] += 1
            slots[rank] = (s, c)
            barrier.wait(timeout=60)
            tot = (slots[0][0] + slots[1][0], slots[0][1] + slots[1][1])
            barrier.wait(timeout=60)
            return tot

        def device_allreduce(tt):
            assert tt.is_cuda and tt.dtype == torch.float64 and tt.numel() == 1
            calls["dev"][rank] += 1
            dev_slots[rank] = tt.clone()
            torch.cuda.current_stream().synchronize()
            barrier.wait(timeout=60)
            tt.copy_(dev_slots[0] + dev_slots[1])
            torch.cuda.current_stream().synchronize()
            barrier.wait(timeout=60)

        if device_hook:
            hook.device_allreduce = device_allreduce
        return hook

    def run(rank):
        try:
            stream = torch.cuda.Stream(device=cuda)
            with torch.cuda.stream(stream):
                sh = shards[rank]
                g = S.CSRGraph(sh.edge_index, sh.x.shape[0])
                results[rank] = S.ops.integrate_dopri5(sh.x, g, params, t, 1e-3, 1e-4, allreduce=hook_for(rank))
                stream.synchronize()
        except Exception as exc:  # pragma: no cover
            errors.append(exc)
            barrier.abort()

    torch.cuda.synchronize()
    threads = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    [th.start() for th in threads]
    [th.join(timeout=120) for th in threads]
    assert not errors, errors
    torch.cuda.synchronize()
    for r in range(2):
        st = results[r][1]
        assert st.accepted == full_stats.accepted
        assert st.n_attempted == full_stats.n_attempted
        for a, b in zip(st.error_ratios, full_stats.error_ratios):
            assert abs(a - b) <= 1e-5 * max(abs(b), 1e-9)
    got = torch.cat([results[0][0], results[1][0]], dim=1)
    assert rel_l2(got, full) <= 1e-5
    n_norms = full_stats.n_attempted + 3             # three norms select the initial step, one per attempted step
    if device_hook:
        assert calls["dev"] == [n_norms, n_norms] and calls["host"] == [1, 1]
    else:
        assert calls["dev"] == [0, 0] and calls["host"] == [n_norms, n_norms]


@pytest.mark.parametrize("n_agv,n_pick,graphs", [(12, 7, 9), (4, 3, 23), (2, 1, 40), (19, 6, 5), (19, 9, 7), (19, 9, 300), (30, 11, 4),
                                                 (40, 11, 3), (26, 0, 5)])
@pytest.mark.parametrize("solver", ["rk4", "euler"])
def test_graph_resident_chain_equals_kernel_per_op_path(cuda, fold_mode, solver, n_agv, n_pick, graphs):
    """With batch.ptr / max_graph_nodes present the folded stages run graph-resident (csrc/chain_fwd.cu): one CTA keeps
    a tile of whole graphs on chip for all stages.  Same arithmetic as the kernel-per-op folded path up to fp32
    summation order; graphs of 35 / 15 nodes pack several per tile, 125 nodes fill one; graphs of 140 (19 AGVs + 9 pickers:
    BASELINE configs[2]), 205 and 255 nodes run as tiles of two 128-row blocks (140-row and 256-row variants), 130-node
    graphs (26 agents) likewise."""
    if fold_mode != "folded":
        pytest.skip("the chain kernel belongs to the folded integrator")
    batch, nxt = S.synthetic.warehouse_batch(graphs, num_agvs=n_agv, num_pickers=n_pick, seed=11)
    D = batch.x.shape[1]
    model = S.GraphODE(D, n_agv, n_pick, hidden_dim=64, ode_solver=solver)
    S.synthetic.init_weights(model, seed=2, conv3_scale=0.05)
    model = model.to(cuda)
    t = torch.tensor([0.0, 0.5, 1.0], device=cuda)
    outs = {}
    for mode in ("chain", "per_op"):
        # (Data.to moves in place and csr_for caches by the edge tensor's storage: the per-op run needs its own copy of the
        # edge list, or it would silently get the tiled graph of the chain run)
        S.graph.clear_cache()
        gb = batch.to(cuda) if mode == "chain" else S.Batch(x=batch.x.to(cuda), edge_index=batch.edge_index.to(cuda).clone())
        if mode == "per_op":
            gb.batch, gb.is_current_agent = batch.batch.to(cuda), batch.is_current_agent.to(cuda)
        model.zero_grad(set_to_none=True)
        out = model(gb, t)
        loss = (out["trajectories"][-1] ** 2).mean()
        loss.backward()
        outs[mode] = (out["node_features"].detach().clone(), [p.grad.clone() for p in model.parameters()])
        S.graph.csr_for(gb.edge_index, gb.x.shape[0], holder=gb).validate()
    e_sol = rel_l2(outs["chain"][0], outs["per_op"][0])
    assert 0.0 < e_sol <= 5e-6, e_sol        # > 0: the two runs really took different kernels
    for a, b in zip(outs["chain"][1], outs["per_op"][1]):
        assert rel_l2(a, b) <= 1e-4
    from swarm_ode_b200 import _lib
    _lib.tc_check(cuda)
