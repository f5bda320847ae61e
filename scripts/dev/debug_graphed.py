import copy, sys, torch
sys.path.insert(0, ".")
import swarm_ode_b200 as S
from swarm_ode_b200.dist import masked_mse_train_step
from swarm_ode_b200.graphed import GraphedTrainStep
from swarm_ode_b200 import graph as G
dev = torch.device("cuda:0")
batch, nxt = S.synthetic.warehouse_batch(32, seed=3)
model_e = S.GraphODE(batch.x.shape[1], 12, 7, hidden_dim=64, ode_solver="rk4")
S.synthetic.init_weights(model_e, seed=1, conv3_scale=0.1)
model_e = model_e.to(dev)
model_g = copy.deepcopy(model_e)
def wsum(m): return float(sum(p.double().abs().sum() for p in m.parameters()))
print("w", wsum(model_e), wsum(model_g))
opt_e = torch.optim.Adam(model_e.parameters(), lr=1e-3, weight_decay=1e-4, capturable=True)
opt_g = torch.optim.Adam(model_g.parameters(), lr=1e-3, weight_decay=1e-4, capturable=True)
t = torch.tensor([0.0, 1.0], device=dev)
b, nx = batch.to(dev), nxt.to(dev)
gs = GraphedTrainStep(model_g, opt_g, b, nx, t, edge_capacity=int(b.edge_index.size(1)) + 17)
print("after capture w", wsum(model_e), wsum(model_g))
model_g.load_state_dict(model_e.state_dict())
print("after reload w", wsum(model_e), wsum(model_g))
le = masked_mse_train_step(model_e, opt_e, b, nx, t); print("le", float(le))
lg = gs.step(b, nx); print("lg", float(lg))
print("after step w", wsum(model_e), wsum(model_g))
