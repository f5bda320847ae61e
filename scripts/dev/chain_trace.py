"""Developer probe: phase timestamps (clock64) of one tile of the forward chain kernel; needs a CHAIN_TRACE build."""
import ctypes as C, sys
import torch
sys.path.insert(0, ".")
import swarm_ode_b200 as S
from swarm_ode_b200 import _lib, graph as G
from swarm_ode_b200.dist import masked_mse_train_step
dev = torch.device("cuda:0")
host, nxt = S.synthetic.warehouse_batch(4096, seed=0)
D = host.x.shape[1]
model = S.GraphODE(D, 12, 7, hidden_dim=64, ode_solver="rk4")
S.synthetic.init_weights(model, seed=1, conv3_scale=0.1)
model = model.to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
b = host.to(dev); nx = nxt.to(dev)
t = torch.tensor([0.0, 1.0], device=dev)
for _ in range(3):
    masked_mse_train_step(model, opt, b, nx, t)
torch.cuda.synchronize()
lib = C.CDLL(_lib.LIB_PATH)
for fn in ("gnode_chain_trace", "gnode_chain_trace_b"):
  buf = (C.c_longlong * 128)()
  rc = getattr(lib, fn)(buf)
  v = list(buf)
  print(fn, "rc", rc)
  base0 = min(x for x in v if x)
  for st in (range(4) if fn == "gnode_chain_trace" else range(3, -1, -1)):
    row = v[16 * st:16 * st + 16]
    base = base0
    print("stage", st, " ".join(f"{(x - base) / 1.965e3:7.2f}" if x else "    -  " for x in row[:14]))
    d = [f"{(row[i + 1] - row[i]) / 1.965e3:6.2f}" if row[i] and row[i + 1] else "   -  " for i in range(13)]
    print("   delta us:", " ".join(d))
