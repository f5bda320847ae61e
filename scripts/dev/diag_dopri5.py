import sys, torch
sys.path.insert(0, ".")
import swarm_ode_b200 as S
from oracle.train_gde_ref import GraphODERef
from tests._util import to_ref_batch, rel_l2
cuda = torch.device("cuda:0")
batch, nxt = S.synthetic.warehouse_batch(6, seed=5)
D = batch.x.shape[1]
for fold, tend in ((True, 0.95), (True, 0.7), (False, 0.7)):
    t = torch.tensor([0.0, tend])
    S.set_fold(fold)
    model = S.GraphODE(D, 12, 7, hidden_dim=64, ode_solver="dopri5")
    S.synthetic.init_weights(model, seed=1, conv3_scale=0.1)
    ref = GraphODERef(D, 12, 7, hidden_dim=64, ode_solver="dopri5")
    ref.load_state_dict(model.state_dict())
    model = model.to(cuda)
    gb = batch.to(cuda)
    out = model(gb, t.to(cuda)); st = model.last_stats
    loss = out["trajectories"].pow(2).mean() + 1e-3 * out["node_features"].pow(2).mean()
    loss.backward()
    rb = to_ref_batch(batch)
    ref.solver_options = {"imposed_dts": list(st.dts)}
    o2 = ref(rb, t)
    l2 = o2["trajectories"].pow(2).mean() + 1e-3 * o2["node_features"].pow(2).mean()
    l2.backward()
    rp = dict(ref.named_parameters())
    errs = {n.replace("ode_func.", "").replace("weight", "w").replace("bias", "b"): f"{rel_l2(p.grad, rp[n].grad):.1e}" for n, p in model.named_parameters()}
    print("fold", fold, "tend", tend, "dts", [round(d, 4) for d in st.dts], "sol", f"{rel_l2(out['node_features'], o2['node_features']):.1e}", errs)
