import sys, torch
sys.path.insert(0, ".")
import swarm_ode_b200 as S
from oracle.train_gde_ref import GraphODERef
from tests._util import rel_l2, to_ref_batch
cuda = torch.device("cuda:0")
batch, _ = S.synthetic.warehouse_batch(5, num_agvs=3, num_pickers=2, seed=9)
D = batch.x.shape[1]; H = 64; solver = "rk4"; t = torch.tensor([0.0, 0.5, 1.0])
model = S.GraphODE(D, 3, 2, hidden_dim=H, ode_solver=solver)
S.synthetic.init_weights(model, seed=4, conv3_scale=0.05)
ref64 = GraphODERef(D, 3, 2, hidden_dim=H, ode_solver=solver).double()
ref64.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
rb = to_ref_batch(batch); rb.x = rb.x.double()
o64 = ref64(rb, t.double()); (o64["trajectories"][-1] ** 2).mean().backward()
r64 = dict(ref64.named_parameters())
model = model.to(cuda)
for engine in ("simt", "auto"):
    for fold in (False, True):
        for tiles in (False, True):
            if tiles and not fold: continue
            S.set_engine(engine); S.set_fold(fold)
            gb = batch.to(cuda)
            if not tiles: gb.__dict__.pop("ptr", None); gb.max_graph_nodes = None
            model.zero_grad(set_to_none=True)
            out = model(gb, t.to(cuda)); (out["trajectories"][-1] ** 2).mean().backward()
            errs = {n.split("ode_func.")[-1]: rel_l2(p.grad, r64[n].grad) for n, p in model.named_parameters()}
            print(engine, "fold" if fold else "direct", "tiles" if tiles else "per-op", "sol", f"{rel_l2(out['node_features'], o64['node_features']):.2e}",
                  " ".join(f"{k}:{v:.1e}" for k, v in errs.items()), flush=True)
