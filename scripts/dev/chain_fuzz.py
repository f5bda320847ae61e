"""Developer probe: randomised chain-vs-per-op equivalence (graph sizes 1..256 mixed, random edges, fwd + bwd)."""
import sys
import torch
sys.path.insert(0, ".")
import swarm_ode_b200 as S
from tests._util import rel_l2
cuda = torch.device("cuda:0")
g = torch.Generator().manual_seed(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
worst = (0.0, 0.0)
for trial in range(int(sys.argv[2]) if len(sys.argv) > 2 else 24):
    hi = [16, 96, 128, 144, 200, 256][trial % 6]
    lo = 1 if trial % 2 else max(1, hi // 2)
    n_graphs = int(torch.randint(1, 40, (1,), generator=g))
    sizes = torch.randint(lo, hi + 1, (n_graphs,), generator=g).tolist()
    D = [24, 37, 64][trial % 3]
    graphs = []
    for n in sizes:
        x = torch.randn(n, D, generator=g)
        m = int(torch.randint(0, 4 * n + 1, (1,), generator=g))
        ei = torch.randint(0, n, (2, m), generator=g) if n > 1 and m > 0 else torch.empty((2, 0), dtype=torch.long)
        graphs.append(S.Data(x=x, edge_index=ei, is_current_agent=torch.ones(n, dtype=torch.bool)))
    batch = S.Batch.from_data_list(graphs)
    solver = ["rk4", "euler", "midpoint"][trial % 3]
    model = S.GraphODE(D, 1, 1, hidden_dim=64, ode_solver=solver)
    S.synthetic.init_weights(model, seed=trial, conv3_scale=0.05)
    model = model.to(cuda)
    t = torch.tensor([0.0, 0.4, 1.0], device=cuda)
    outs = {}
    for mode in ("chain", "per_op"):
        # (Data.to moves in place and csr_for caches by the edge tensor's storage: the per-op run needs its own copy of the
        # edge list, or it would silently get the tiled graph of the chain run)
        S.graph.clear_cache()
        gb = batch.to(cuda) if mode == "chain" else S.Batch(x=batch.x.to(cuda), edge_index=batch.edge_index.to(cuda).clone())
        if mode == "per_op":
            gb.batch, gb.is_current_agent = batch.batch.to(cuda), batch.is_current_agent.to(cuda)
        model.zero_grad(set_to_none=True)
        out = model(gb, t)
        (out["trajectories"][-1] ** 2).mean().backward()
        outs[mode] = (out["node_features"].detach().clone(), [p.grad.clone() for p in model.parameters()])
        S.graph.csr_for(gb.edge_index, gb.x.shape[0], holder=gb).validate()
    e_sol = rel_l2(outs["chain"][0], outs["per_op"][0])
    e_grad = max(rel_l2(a, b) for a, b in zip(outs["chain"][1], outs["per_op"][1]))
    worst = (max(worst[0], e_sol), max(worst[1], e_grad))
    flag = "" if (e_sol <= 5e-6 and e_grad <= 1e-4) else "   <-- CHECK"
    print(f"trial {trial:2d}: {n_graphs:2d} graphs of {lo}..{hi} nodes, D={D}, {solver:8s} sol {e_sol:.1e} grad {e_grad:.1e}{flag}")
from swarm_ode_b200 import _lib
_lib.tc_check(cuda)
print("worst", worst)
