"""Developer probe: does the H2D upload of the next batch overlap with the training step?"""
import sys, time
import torch
sys.path.insert(0, ".")
import swarm_ode_b200 as S
from swarm_ode_b200 import graph as G
from swarm_ode_b200.dist import masked_mse_train_step

dev = torch.device("cuda:0")
host, nxt = S.synthetic.warehouse_batch(4096, seed=0)
host.pin_memory(); nxt = nxt.pin_memory()
print("pinned:", host.x.is_pinned(), host.edge_index.is_pinned(), nxt.is_pinned())
D = host.x.shape[1]
model = S.GraphODE(D, 12, 7, hidden_dim=64, ode_solver="rk4")
S.synthetic.init_weights(model, seed=1, conv3_scale=0.1)
model = model.to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)
t = torch.tensor([0.0, 1.0], device=dev)
cs = torch.cuda.Stream(device=dev)

def upload(stream):
    with torch.cuda.stream(stream):
        b = S.Batch(x=host.x.to(dev, non_blocking=True), edge_index=host.edge_index.to(dev, non_blocking=True))
        b.batch = host.batch.to(dev, non_blocking=True)
        b.is_current_agent = host.is_current_agent.to(dev, non_blocking=True)
        nx = nxt.to(dev, non_blocking=True)
    return b, nx

b, nx = upload(torch.cuda.current_stream()); torch.cuda.synchronize()
def step(b, nx):
    G.clear_cache(); b.__dict__.pop("_gnode_csr", None)
    return masked_mse_train_step(model, opt, b, nx, t)
for _ in range(5): step(b, nx)
torch.cuda.synchronize()

def timeit(fn, reps=6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3

print("compute only      %.2f ms" % timeit(lambda: step(b, nx)))
keep = []
def up_only():
    keep.clear(); keep.append(upload(cs))
print("upload only       %.2f ms" % timeit(up_only))
def both():
    keep.clear(); keep.append(upload(cs)); step(b, nx)
print("upload || compute %.2f ms" % timeit(both))
def both_sync():
    keep.clear(); keep.append(upload(cs)); float(step(b, nx))
print("upload || compute + float(loss) %.2f ms" % timeit(both_sync))
