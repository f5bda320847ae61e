"""Developer probe: per-kernel-class device time of one dopri5 solve (configs[2] shape)."""
import sys
import torch
sys.path.insert(0, ".")
import swarm_ode_b200 as S
from swarm_ode_b200 import _lib
graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1           # bench.py's dopri5_strong leg: 16384 8.0 0,1,2,3
times = [float(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0.0, 1.0]
dev = torch.device("cuda:0")
batch, _ = S.synthetic.warehouse_batch(graphs, num_agvs=19, num_pickers=9, seed=0)
D = batch.x.shape[1]
model = S.GraphODE(D, 19, 9, hidden_dim=64, ode_solver="dopri5")
S.synthetic.init_weights(model, seed=1, conv3_scale=scale)
model = model.to(dev)
b = batch.to(dev)
t = torch.tensor(times, device=dev)
with torch.no_grad():
    for _ in range(2):
        model(b, t)
    torch.cuda.synchronize()
    _lib.prof_enable(True)
    model(b, t)
    prof = _lib.prof_read(); _lib.prof_enable(False)
prof.sort(key=lambda p: -p["ms"])
tot = sum(p["ms"] for p in prof)
print(f"N={b.x.shape[0]} D={D} nfe={model.last_stats.nfe} accepted={model.last_stats.n_accepted}/{model.last_stats.n_attempted}: library kernels {tot:.2f} ms")
for p in prof[:14]:
    if p["launches"]:
        print(f"  {p['ms']:8.3f} ms {p['launches']:4d} x {p['ms']/p['launches']*1e3:8.1f} us {p['bytes']/max(p['ms'],1e-9)/1e6:8.0f} GB/s  {p['name']}")
