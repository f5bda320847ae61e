"""Developer probe: per-kernel-class device time of the bench training step (library profiler) + torch-side gaps."""
import sys, time
import torch
sys.path.insert(0, ".")
import swarm_ode_b200 as S
from swarm_ode_b200 import _lib, graph as G
from swarm_ode_b200.dist import masked_mse_train_step

graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
solver = sys.argv[2] if len(sys.argv) > 2 else "rk4"
steps = 5
dev = torch.device("cuda:0")
host, nxt = S.synthetic.warehouse_batch(graphs, seed=0)
D = host.x.shape[1]
model = S.GraphODE(D, 12, 7, hidden_dim=64, ode_solver=solver)
S.synthetic.init_weights(model, seed=1, conv3_scale=0.1)
model = model.to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)
b = host.to(dev); nx = nxt.to(dev)
t = torch.tensor([0.0, 1.0], device=dev)

def step():
    G.clear_cache(); b.__dict__.pop("_gnode_csr", None)
    return masked_mse_train_step(model, opt, b, nx, t)

for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(steps): step()
e1.record(); torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / steps * 1e3
print(f"step: {e0.elapsed_time(e1)/steps:.3f} ms device, {wall:.3f} ms wall")
evs = [torch.cuda.Event(enable_timing=True) for _ in range(13)]
hs = []
evs[0].record()
for i in range(12):
    h0 = time.perf_counter(); step(); hs.append((time.perf_counter() - h0) * 1e3); evs[i + 1].record()
torch.cuda.synchronize()
print("per-step device ms:", " ".join(f"{evs[i].elapsed_time(evs[i+1]):.1f}" for i in range(12)))
print("per-step host enqueue ms:", " ".join(f"{h:.1f}" for h in hs))
_lib.prof_enable(True)
for _ in range(steps): step()
prof = _lib.prof_read(); _lib.prof_enable(False)
prof.sort(key=lambda p: -p["ms"])
tot = sum(p["ms"] for p in prof)
print(f"library kernels: {tot/steps:.3f} ms per step in {sum(p['launches'] for p in prof)/steps:.0f} profiled scopes")
for p in prof:
    if p["launches"]:
        print(f"  {p['ms']/steps:8.3f} ms/step  {p['launches']/steps:5.1f} x {p['ms']/p['launches']*1e3:8.1f} us  "
              f"{p['bytes']/max(p['ms'],1e-9)/1e6:8.0f} GB/s {p['flops']/max(p['ms'],1e-9)/1e9:8.1f} TF/s  {p['name']}")
_lib.tc_check(dev)
