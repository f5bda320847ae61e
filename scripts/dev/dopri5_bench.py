"""Developer probe: dopri5 forward (BASELINE configs[2] shape: 19 AGV + 9 pickers, D=435) folded vs direct."""
import sys, time
import torch
sys.path.insert(0, ".")
import swarm_ode_b200 as S

graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dev = torch.device("cuda:0")
batch, _ = S.synthetic.warehouse_batch(graphs, num_agvs=19, num_pickers=9, seed=0)
D = batch.x.shape[1]
model = S.GraphODE(D, 19, 9, hidden_dim=64, ode_solver="dopri5")
S.synthetic.init_weights(model, seed=1, conv3_scale=0.1)
model = model.to(dev)
b = batch.to(dev)
t = torch.tensor([0.0, 1.0], device=dev)
for fold in (True, False):
    S.set_fold(fold)
    with torch.no_grad():
        for _ in range(2):
            out = model(b, t)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            out = model(b, t)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / reps * 1e3
    st = model.last_stats
    units = b.x.shape[0] * st.nfe
    print(f"fold={fold}: N={b.x.shape[0]} D={D} {ms:.2f} ms  nfe={st.nfe} accepted={st.n_accepted}/{st.n_attempted} "
          f"{units / ms / 1e3:.1f} M agent-state-steps/s  min_margin={st.min_margin:.3e}", flush=True)
