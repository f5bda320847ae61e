"""Developer probe: bench.py's e2e loop, bisected."""
import sys, time
import torch
sys.path.insert(0, ".")
import swarm_ode_b200 as S
from swarm_ode_b200 import graph as G
from swarm_ode_b200.dist import masked_mse_train_step

dev = torch.device("cuda:0")
host, nxt_host = S.synthetic.warehouse_batch(4096, seed=0)
host.pin_memory(); nxt_host = nxt_host.pin_memory()
D = host.x.shape[1]
model = S.GraphODE(D, 12, 7, hidden_dim=64, ode_solver="rk4")
S.synthetic.init_weights(model, seed=1, conv3_scale=0.1)
model = model.to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)
t_dev = torch.tensor([0.0, 1.0], device=dev)
copy_stream = torch.cuda.Stream(device=dev)
pending = {}

def to_device(nb=True):
    b = S.Batch(x=host.x.to(dev, non_blocking=nb), edge_index=host.edge_index.to(dev, non_blocking=nb))
    b.batch = host.batch.to(dev, non_blocking=nb)
    b.is_current_agent = host.is_current_agent.to(dev, non_blocking=nb)
    return b, nxt_host.to(dev, non_blocking=nb)

def upload():
    with torch.cuda.stream(copy_stream):
        b, nx = to_device(True)
        ev = torch.cuda.Event(); ev.record(copy_stream)
    pending["next"] = (b, nx, ev)

def step_e2e(record=True, sync=True):
    G.clear_cache()
    b, nx, ev = pending.pop("next")
    torch.cuda.current_stream(dev).wait_event(ev)
    upload()
    loss = masked_mse_train_step(model, opt, b, nx, t_dev)
    if record:
        for t in (b.x, b.edge_index, b.batch, b.is_current_agent, nx):
            t.record_stream(torch.cuda.current_stream(dev))
    return float(loss) if sync else loss

def run(label, **kw):
    pending.clear(); upload()
    for _ in range(4): step_e2e(**kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(8):
        t0 = time.perf_counter(); step_e2e(**kw); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print(label, " ".join(f"{x:.1f}" for x in ts), flush=True)

resident, nres = to_device(False)
def step_res():
    G.clear_cache(); resident.__dict__.pop("_gnode_csr", None)
    return masked_mse_train_step(model, opt, resident, nres, t_dev)
for _ in range(5): step_res()
torch.cuda.synchronize()
ts = []
for _ in range(8):
    t0 = time.perf_counter(); float(step_res()); ts.append((time.perf_counter() - t0) * 1e3)
print("resident+sync", " ".join(f"{x:.1f}" for x in ts))
run("e2e record+sync", record=True, sync=True)
run("e2e norecord   ", record=False, sync=True)
run("e2e nosync     ", record=True, sync=False)
print(torch.cuda.memory_summary(abbreviated=True)[:1500])
