import sys, torch
sys.path.insert(0, ".")
import swarm_ode_b200 as S
cuda = torch.device("cuda:0")
batch, nxt = S.synthetic.warehouse_batch(6, seed=5)
D = batch.x.shape[1]
t = torch.tensor([0.0, 1.0])
for fold in (True, False, True):
    S.set_fold(fold)
    model = S.GraphODE(D, 12, 7, hidden_dim=64, ode_solver="dopri5")
    S.synthetic.init_weights(model, seed=1, conv3_scale=0.1)
    model = model.to(cuda)
    gb = batch.to(cuda); gb.x = gb.x.clone().requires_grad_(True)
    out = model(gb, t.to(cuda))
    print("fold", fold, "x.requires_grad", gb.x.requires_grad, "is_leaf", gb.x.is_leaf, "grad_fn", out["node_features"].grad_fn)
    loss = out["trajectories"].pow(2).mean() + 1e-3 * out["node_features"].pow(2).mean()
    loss.backward()
    print("   x.grad", None if gb.x.grad is None else float(gb.x.grad.norm()), "w grad", float(model.ode_func.conv1.lin_l.weight.grad.norm()))
