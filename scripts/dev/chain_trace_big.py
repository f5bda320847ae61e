"""Developer probe: phase timestamps of one two-block tile (140-node graphs, dopri5 stages) of the forward chain; needs -DCHAIN_TRACE."""
import ctypes as C, sys
import torch
sys.path.insert(0, ".")
import swarm_ode_b200 as S
from swarm_ode_b200 import _lib
dev = torch.device("cuda:0")
batch, _ = S.synthetic.warehouse_batch(4096, num_agvs=19, num_pickers=9, seed=0)
model = S.GraphODE(batch.x.shape[1], 19, 9, hidden_dim=64, ode_solver="dopri5")
S.synthetic.init_weights(model, seed=1, conv3_scale=0.1)
model = model.to(dev); b = batch.to(dev); t = torch.tensor([0.0, 1.0], device=dev)
with torch.no_grad():
    for _ in range(2):
        model(b, t)
torch.cuda.synchronize()
lib = C.CDLL(_lib.LIB_PATH)
buf = (C.c_longlong * 128)()
lib.gnode_chain_trace(buf)
v = list(buf)
base0 = min(x for x in v if x)
names = ["in", "sync", "hoff1", "acc1", "epi1", "agg1", "agg2", "hoff2", "store1", "acc2", "epi2", "agg3", "store2", "end"]
for st in range(7):
    row = v[16 * st:16 * st + 14]
    d = [f"{(row[i + 1] - row[i]) / 1.965e3:6.2f}" if row[i] and row[i + 1] else "   -  " for i in range(13)]
    print("stage", st, f"start {(row[0] - base0) / 1.965e3:7.1f} us; deltas:", " ".join(d))
