"""Developer probe: per-kernel-class device time of the configs[3] dense train step (256-agent complete graphs)."""
import sys, torch
sys.path.insert(0, ".")
import swarm_ode_b200 as S
from swarm_ode_b200 import _lib, graph as G
from swarm_ode_b200.dist import masked_mse_train_step
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
D = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda:0")
b = S.synthetic.dense_batch(B, num_agents=256, node_dim_=D, seed=1).to(dev)
nxt = torch.randn(B, 256, 2, device=dev)
model = S.GraphODE(D, 256, 0, hidden_dim=64, ode_solver="rk4")
S.synthetic.init_weights(model, seed=1, conv3_scale=0.02)
model = model.to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)
t = torch.tensor([0.0, 1.0], device=dev)
def step():
    G.clear_cache(); b.__dict__.pop("_gnode_csr", None)
    return masked_mse_train_step(model, opt, b, nxt, t)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step()
e1.record(); torch.cuda.synchronize()
print(f"dense B={B} D={D}: {e0.elapsed_time(e1)/5:.3f} ms per rk4 train step")
_lib.prof_enable(True)
for _ in range(3): step()
prof = _lib.prof_read(); _lib.prof_enable(False)
prof.sort(key=lambda p: -p["ms"])
for p in prof[:10]:
    if p["launches"]:
        print(f"  {p['ms']/3:8.3f} ms/step {p['launches']/3:5.1f} x  {p['name']}")
