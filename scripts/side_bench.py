#!/usr/bin/env python
"""Side benches for the BASELINE.json configs that bench.py's contract line does not carry (one GPU, CUDA events):

    python scripts/side_bench.py sweep      # configs[4]: agents 16-1024 x trajectories 1k-64k, rk4 10 steps, fused path
                                            #             vs an eager-torch restatement on the SAME GPU (and on the host cores)
    python scripts/side_bench.py dense      # configs[3]: 256-agent complete graphs, train_gde.py-style loop (euler + rk4)
    python scripts/side_bench.py baselines  # the reference's GRU / LSTM predictors (cuDNN through torch, no custom
                                            #            kernels) on the same windows, next to GraphODE
    python scripts/side_bench.py secondary  # rows a11 / a12: ODEFunction solves and the batched HeteroGraphODENetwork
    python scripts/side_bench.py all [--out gpurun_out/r2_side.json]

Every mode prints JSON rows (one per line) and `all` also writes a markdown summary next to the JSON.  The "eager"
columns are a plain-PyTorch statement of the same algorithm written here (SAGEConv by index_add, torchdiffeq's 3/8-rule
rk4; scripts/train_gde.py:20-45,78-85) -- what a user gets from the reference's modules on a GPU, launch for launch."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import swarm_ode_b200 as S  # noqa: E402
from swarm_ode_b200.dist import masked_mse_train_step  # noqa: E402
from swarm_ode_b200 import graph as G  # noqa: E402

DEV = torch.device("cuda", 0)


# ------------------------------------------------------------------------------------------------ eager restatement
class EagerSage(nn.Module):
    def __init__(self, ci, co):
        super().__init__()
        self.lin_l, self.lin_r = nn.Linear(ci, co, bias=True), nn.Linear(ci, co, bias=False)

    def forward(self, x, ei, inv_deg):
        agg = torch.zeros_like(x).index_add_(0, ei[1], x.index_select(0, ei[0])) * inv_deg
        return self.lin_l(agg) + self.lin_r(x)


class EagerGraphODE(nn.Module):
    def __init__(self, D, H=64, solver="rk4"):
        super().__init__()
        self.c1, self.c2, self.c3 = EagerSage(D, H), EagerSage(H, H), EagerSage(H, D)
        self.dec, self.solver = nn.Linear(D, 2), solver

    def load_from(self, m: "S.GraphODE"):
        f = m.ode_func
        for mine, theirs in ((self.c1, f.conv1), (self.c2, f.conv2), (self.c3, f.conv3)):
            mine.lin_l.weight.data.copy_(theirs.lin_l.weight.data); mine.lin_l.bias.data.copy_(theirs.lin_l.bias.data)
            mine.lin_r.weight.data.copy_(theirs.lin_r.weight.data)
        self.dec.weight.data.copy_(m.position_decoder.weight.data); self.dec.bias.data.copy_(m.position_decoder.bias.data)
        return self

    def field(self, x, ei, inv_deg):
        h = F.relu(self.c1(x, ei, inv_deg))
        h = F.relu(self.c2(h, ei, inv_deg))
        return self.c3(h, ei, inv_deg)

    def forward(self, x, ei, t):
        deg = torch.zeros(x.shape[0], device=x.device).index_add_(0, ei[1], torch.ones(ei.shape[1], device=x.device))
        inv_deg = (1.0 / deg.clamp_min(1.0)).unsqueeze(1)
        ys, y = [x], x
        for j in range(len(t) - 1):
            dt = float(t[j + 1] - t[j])
            f = lambda v: self.field(v, ei, inv_deg)  # noqa: E731
            if self.solver == "euler":
                y = y + dt * f(y)
            else:       # torchdiffeq rk4 = 3/8 rule
                k1 = f(y); k2 = f(y + dt * k1 / 3); k3 = f(y + dt * (k2 - k1 / 3)); k4 = f(y + dt * (k1 - k2 + k3))
                y = y + dt * (k1 + 3 * (k2 + k3) + k4) / 8
            ys.append(y)
        sol = torch.stack(ys)
        return sol, self.dec(sol)


def timed(fn, iters, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


# ------------------------------------------------------------------------------------------------ configs[4]: sweep
def sweep(rows, quick=False):
    D, H, steps = 128, 64, 10
    t = torch.linspace(0.0, 1.0, steps + 1)
    budget = 0.55 * torch.cuda.get_device_properties(DEV).total_memory
    ns = [16, 32, 64, 128, 256, 512, 1024] if not quick else [16, 128]
    Bs = [1024, 4096, 16384, 65536] if not quick else [1024]
    for n in ns:
        proto = S.synthetic.geometric_batch(min(256, Bs[0]), num_agents=n, node_dim_=D, seed=n)     # 256 distinct graphs, tiled to B
        for B in Bs:
            N = B * n
            need = 4.0 * N * D * (steps + 1) * 2.2 + 16.0 * proto.edge_index.shape[1] * (B / proto.num_graphs) * 3
            if need > budget or N > 40_000_000:
                rows.append({"bench": "sweep", "agents": n, "trajectories": B, "skipped": "exceeds 55 % of HBM"})
                continue
            reps = B // proto.num_graphs
            b = _tile(proto, reps).to(DEV)
            model = S.GraphODE(D, n, 0, hidden_dim=H, ode_solver="rk4")
            S.synthetic.init_weights(model, seed=1, conv3_scale=0.02)
            model = model.to(DEV)
            eager = EagerGraphODE(D, H, "rk4").to(DEV).load_from(model)
            td = t.to(DEV)
            with torch.no_grad():
                def ours():
                    G.clear_cache(); b.__dict__.pop("_gnode_csr", None)
                    return model(b, td)
                ms = timed(ours, 3 if N > 4e6 else 6)
                out = ours()
                units = N * 4 * steps
                row = {"bench": "sweep", "agents": n, "trajectories": B, "nodes": N, "edges": int(b.edge_index.shape[1]), "D": D,
                       "rk4_steps": steps, "path": "chain" if n <= 256 else "per-op", "ours_ms": ms, "ours_units_per_s": units / (ms * 1e-3)}
                eager_need = 4.0 * (b.edge_index.shape[1] * D * 2 + N * D * (steps + 8))
                if eager_need < budget:
                    try:
                        ems = timed(lambda: eager(b.x, b.edge_index, t), 2 if N > 1e6 else 4, warmup=1)
                        sol_e, _ = eager(b.x, b.edge_index, t)
                        row.update({"eager_gpu_ms": ems, "eager_gpu_units_per_s": units / (ems * 1e-3), "speedup_vs_eager_gpu": ems / ms,
                                    "rel_l2_vs_eager": rel_l2(out["node_features"], sol_e)})
                        del sol_e
                    except torch.OutOfMemoryError:
                        row["eager_gpu_ms"] = None
                if B == Bs[0]:        # host cores: one point per agent count
                    cb = _tile(proto, max(1, 64 // proto.num_graphs) if proto.num_graphs < 64 else 1)
                    ce = EagerGraphODE(D, H, "rk4").load_from(model.cpu()); model.to(DEV)
                    torch.set_num_threads(os.cpu_count() or 1)
                    w0 = time.perf_counter(); ce(cb.x, cb.edge_index, t); cpu_s = time.perf_counter() - w0
                    row.update({"cpu_units_per_s": cb.x.shape[0] * 4 * steps / cpu_s, "cpu_cores": os.cpu_count(), "cpu_graphs": cb.num_graphs})
            rows.append(row)
            print(json.dumps(row), flush=True)
            del b, model, eager, out
            torch.cuda.empty_cache()


def _tile(batch, reps):
    if reps <= 1:
        return batch
    N, Gn = batch.x.shape[0], batch.num_graphs
    out = S.Batch(x=batch.x.repeat(reps, 1))
    offs = (torch.arange(reps, dtype=torch.long) * N).view(reps, 1, 1)
    out.edge_index = (batch.edge_index.unsqueeze(0) + offs).permute(1, 0, 2).reshape(2, -1)
    out.batch = (batch.batch.unsqueeze(0) + (torch.arange(reps, dtype=torch.long) * Gn).view(reps, 1)).reshape(-1)
    out.ptr = torch.cat([torch.zeros(1, dtype=torch.long), (batch.ptr[1:].unsqueeze(0) + (torch.arange(reps, dtype=torch.long) * N).view(reps, 1)).reshape(-1)])
    out.is_current_agent = batch.is_current_agent.repeat(reps)
    out.num_graphs, out.max_graph_nodes = Gn * reps, batch.max_graph_nodes
    return out


# ------------------------------------------------------------------------------------------------ configs[3]: dense train loop
def dense(rows, quick=False):
    n, iters = 256, (100 if not quick else 5)
    for D, B in ((64, 1024), (435, 1024)) if not quick else ((64, 64),):
        proto = S.synthetic.dense_batch(64, num_agents=n, node_dim_=D, seed=D)
        b = _tile(proto, B // 64).to(DEV)
        nxt = torch.randn(B, n, 2, device=DEV)
        for solver in ("euler", "rk4"):
            model = S.GraphODE(D, n, 0, hidden_dim=64, ode_solver=solver)
            S.synthetic.init_weights(model, seed=1, conv3_scale=0.02)
            model = model.to(DEV)
            opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=True)
            t = torch.tensor([0.0, 1.0], device=DEV)
            losses = []

            def step():
                G.clear_cache(); b.__dict__.pop("_gnode_csr", None)
                losses.append(masked_mse_train_step(model, opt, b, nxt, t))
            ms = timed(step, iters, warmup=3)
            stages = {"euler": 1, "rk4": 4}[solver]
            units = b.x.shape[0] * stages
            row = {"bench": "dense", "agents": n, "graphs_per_gpu": B, "graphs_total_config": 8192, "gpus_config": 8, "D": D, "solver": solver,
                   "nodes": int(b.x.shape[0]), "edges": int(b.edge_index.shape[1]), "train_step_ms": ms, "units_per_s": units / (ms * 1e-3),
                   "iters": iters, "loss_first": float(losses[3]), "loss_last": float(losses[-1]),
                   "hbm_frac_8d_step": units * 3.0 * (8.0 * D + 4.0 * (n - 1) + 4.0) / (ms * 1e-3) / 6521.4e9}
            rows.append(row)
            print(json.dumps(row), flush=True)
        del b
        torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------ f4: sequence baselines
class SeqPredictor(nn.Module):
    """GRU/LSTMTrajectoryPredictor of scripts/train_baselines.py:128-243 (encoder Linear, 2-layer cuDNN RNN with dropout 0.1,
    decoder Linear-ReLU-Linear on the last hidden state), stated with torch modules -- comparison rows only."""

    def __init__(self, kind, obs_dim, hidden_dim=128, num_layers=2):
        super().__init__()
        self.encoder = nn.Linear(obs_dim, hidden_dim)
        self.rnn = (nn.GRU if kind == "gru" else nn.LSTM)(hidden_dim, hidden_dim, num_layers=num_layers, batch_first=True, dropout=0.1)
        self.decoder = nn.Sequential(nn.Linear(hidden_dim, hidden_dim // 2), nn.ReLU(), nn.Linear(hidden_dim // 2, 2))

    def forward(self, obs):                      # [B, W, n, D] -> [B, n, 2]
        Bn, W, n, D = obs.shape
        h, _ = self.rnn(self.encoder(obs.permute(0, 2, 1, 3).reshape(Bn * n, W, D)))
        return self.decoder(h[:, -1]).view(Bn, n, 2)


def baselines(rows, quick=False):
    B = 4096 if not quick else 64
    batch, nxt = S.synthetic.warehouse_batch(B, seed=0)
    D, W, n = batch.x.shape[1], 5, 19
    obs = batch.x.view(B, W, n, D).to(DEV)
    tgt = nxt.to(DEV)
    for kind in ("gru", "lstm"):
        torch.manual_seed(0)
        m = SeqPredictor(kind, D).to(DEV)
        opt = torch.optim.Adam(m.parameters(), lr=1e-3, fused=True)

        def step():
            opt.zero_grad(set_to_none=True)
            loss = F.mse_loss(m(obs), tgt)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
            opt.step()
        ms = timed(step, 10, warmup=3)
        row = {"bench": "baselines", "model": f"{kind.upper()}TrajectoryPredictor (cuDNN via torch)", "trajectories": B, "window": W, "agents": n, "obs_dim": D,
               "train_step_ms": ms, "samples_per_s": B / (ms * 1e-3), "agent_windows_per_s": B * n / (ms * 1e-3)}
        rows.append(row)
        print(json.dumps(row), flush=True)
    gb = batch.to(DEV)
    for solver in ("euler", "rk4"):
        model = S.GraphODE(D, 12, 7, hidden_dim=64, ode_solver=solver)
        S.synthetic.init_weights(model, seed=1, conv3_scale=0.1)
        model = model.to(DEV)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=True)
        t = torch.tensor([0.0, 1.0], device=DEV)

        def gstep():
            G.clear_cache(); gb.__dict__.pop("_gnode_csr", None)
            masked_mse_train_step(model, opt, gb, tgt, t)
        ms = timed(gstep, 10, warmup=3)
        row = {"bench": "baselines", "model": f"GraphODE {solver} (this library)", "trajectories": B, "window": W, "agents": n, "obs_dim": D,
               "train_step_ms": ms, "samples_per_s": B / (ms * 1e-3), "agent_windows_per_s": B * n / (ms * 1e-3)}
        rows.append(row)
        print(json.dumps(row), flush=True)


# ------------------------------------------------------------------------------------------------ a11 / a12
def secondary(rows, quick=False):
    from swarm_ode_b200.hetero import MultiAgentGraphConverter
    # a11: ODEFunction (Linear-tanh-Linear-tanh-Linear, H = 128, h = 32) integrated over M rows, euler 1 step / dopri5 defaults
    for M in (268, 65536, 1048576) if not quick else (268,):
        f = S.ODEFunction(128, 32).to(DEV)
        y0 = torch.randn(M, 128, device=DEV) * 0.5
        for method, kw in (("euler", {}), ("rk4", {}), ("dopri5", dict(rtol=1e-3, atol=1e-4))):
            t = torch.tensor([0.0, 1.0], device=DEV)
            with torch.no_grad():
                ms = timed(lambda: S.odeint(f, y0, t, method=method, **kw), 10, warmup=2)
            def train():
                f.zero_grad(set_to_none=True)
                S.odeint(f, y0, t, method=method, **kw)[-1].pow(2).mean().backward()
            tms = timed(train, 10, warmup=2)
            row = {"bench": "secondary", "what": "ODEFunction odeint (a11)", "rows": M, "method": method, "forward_ms": ms, "fwd_bwd_ms": tms,
                   "rows_per_s_forward": M / (ms * 1e-3)}
            rows.append(row)
            print(json.dumps(row), flush=True)
    # a12: HeteroGraphODENetwork (run_gnode.py form) over many converter graphs in ONE forward
    rng = np.random.default_rng(0)
    na, npk, nl = 19, 9, 160
    racks = [(int(c % 25) + 1, int(c // 25) + 1, int(g)) for c, g in zip(rng.permutation(25 * 22)[:nl], rng.integers(0, 4, nl))]
    Dobs = 7 + 4 * (na + npk - 1) + 2 * nl

    def one():
        obs = np.zeros((na + npk, Dobs), dtype=np.float32)
        obs[:na, :3] = rng.integers(0, 2, (na, 3)); obs[:na, 3:5] = rng.integers(0, 22, (na, 2)); obs[na:, 0:2] = rng.integers(0, 22, (npk, 2))
        sh = np.zeros(2 * nl, dtype=np.float32); sh[0::2] = rng.random(nl) < 0.9; sh[1::2] = rng.random(nl) < 0.12
        obs[0, 7 + 4 * (na + npk - 1):] = sh
        return MultiAgentGraphConverter(na, npk)._build_graph_from_observation(obs, racks)
    w0 = time.perf_counter()
    graphs = [one() for _ in range(64 if not quick else 4)]
    conv_ms = (time.perf_counter() - w0) / len(graphs) * 1e3
    net = S.HeteroGraphODENetwork({"agv": 7, "picker": 4, "location": 2}, action_size=5, hidden_dim=128, num_layers=2, ode_hidden_dim=32).to(DEV)
    for nb in (1, 64, 1024) if not quick else (1, 4):
        data = S.HeteroData.from_data_list([graphs[i % len(graphs)] for i in range(nb)]).to(DEV)
        with torch.no_grad():
            ms = timed(lambda: net(data, integration_time=1.0), 10, warmup=3)
        def train():
            net.zero_grad(set_to_none=True)
            o = net(data, integration_time=1.0)
            (o["agv_q_values"].pow(2).mean() + o["picker_q_values"].pow(2).mean()).backward()
        tms = timed(train, 10, warmup=3)
        row = {"bench": "secondary", "what": "HeteroGraphODENetwork forward (a12, run_gnode.py form)", "graphs_per_call": nb, "forward_ms": ms,
               "fwd_bwd_ms": tms, "graphs_per_s_forward": nb / (ms * 1e-3), "converter_ms_per_graph_host": conv_ms}
        rows.append(row)
        print(json.dumps(row), flush=True)


def markdown(rows):
    out = ["# Side benches (scripts/side_bench.py, one B200)\n"]
    sw = [r for r in rows if r.get("bench") == "sweep" and "ours_ms" in r]
    if sw:
        out.append("## configs[4] sweep: rk4, 10 steps, D = 128, H = 64, random geometric graphs (forward solve, all 11 time points kept)\n")
        out.append("| agents | trajectories | nodes | path | ours ms | ours M units/s | eager-torch GPU ms | speed-up | rel-L2 vs eager | CPU M units/s (cores) |\n|---:|---:|---:|---|---:|---:|---:|---:|---:|---:|")
        for r in sw:
            out.append(f"| {r['agents']} | {r['trajectories']} | {r['nodes']} | {r['path']} | {r['ours_ms']:.2f} | {r['ours_units_per_s'] / 1e6:.0f} | "
                       + (f"{r['eager_gpu_ms']:.1f} | {r['speedup_vs_eager_gpu']:.1f}x | {r['rel_l2_vs_eager']:.1e}" if r.get("eager_gpu_ms") else "- | - | -")
                       + (f" | {r['cpu_units_per_s'] / 1e6:.2f} ({r['cpu_cores']})" if "cpu_units_per_s" in r else " | ") + " |")
        out.append("")
    dn = [r for r in rows if r.get("bench") == "dense"]
    if dn:
        out.append("## configs[3]: 256 agents, complete interaction graph, train_gde.py-style step (per-GPU share of the 8k batch on 8 GPUs)\n")
        out.append("| D | solver | graphs/GPU | nodes | edges | train step ms | M units/s | 8(d) step frac | loss first -> last |\n|---:|---|---:|---:|---:|---:|---:|---:|---|")
        for r in dn:
            out.append(f"| {r['D']} | {r['solver']} | {r['graphs_per_gpu']} | {r['nodes']} | {r['edges']} | {r['train_step_ms']:.2f} | {r['units_per_s'] / 1e6:.0f} | "
                       f"{r['hbm_frac_8d_step']:.3f} | {r['loss_first']:.4g} -> {r['loss_last']:.4g} |")
        out.append("")
    bl = [r for r in rows if r.get("bench") == "baselines"]
    if bl:
        out.append("## f4: sequence baselines of scripts/train_baselines.py on the same windows (full train step, 4096 trajectories)\n")
        out.append("| model | train step ms | trajectories/s |\n|---|---:|---:|")
        for r in bl:
            out.append(f"| {r['model']} | {r['train_step_ms']:.2f} | {r['samples_per_s']:.0f} |")
        out.append("")
    sc = [r for r in rows if r.get("bench") == "secondary"]
    if sc:
        out.append("## a11 / a12: secondary variants\n")
        out.append("| what | size | method | forward ms | fwd+bwd ms |\n|---|---:|---|---:|---:|")
        for r in sc:
            out.append(f"| {r['what']} | {r.get('rows', r.get('graphs_per_call'))} | {r.get('method', 'euler')} | {r['forward_ms']:.3f} | {r['fwd_bwd_ms']:.3f} |")
        out.append("")
    return "\n".join(out)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["sweep", "dense", "baselines", "secondary", "all"])
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    assert torch.cuda.is_available(), "side benches need a CUDA device"
    torch.cuda.set_device(DEV)
    rows = []
    for name, fn in (("sweep", sweep), ("dense", dense), ("baselines", baselines), ("secondary", secondary)):
        if a.mode in (name, "all"):
            fn(rows, a.quick)
    S._lib.tc_check(DEV)
    if a.out:
        with open(a.out, "w") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")
        with open(os.path.splitext(a.out)[0] + ".md", "w") as f:
            f.write(markdown(rows))
