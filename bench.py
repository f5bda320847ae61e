#!/usr/bin/env python
"""bench.py -- GNODE agent-state-steps/s on B200 (the reference's headline metric, BASELINE.json).

    python bench.py --gpus N --steps K --warmup W              # this framework (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU path (oracle port)

Workload (N=1): BASELINE.json configs[1] -- "GNODE, RK4 forward+backward training step, 4096
trajectories on 1 B200 (fp32)", medium warehouse, 12 AGVs + 7 pickers, D=399, 95 nodes/graph.
A step = one full training step of scripts/train_gde.py:478-495 (forward through the rk4 solver,
masked MSE, backward through the solver, grad clip, Adam) over one synthetic batch.
Metric unit = one node x one vector-field (RK stage) evaluation = "agent-state-step".
With N GPUs every rank runs its own 4096-graph shard (weak scaling) and gradients are all-reduced.

One JSON line is printed by rank 0; see DESIGN.md "Measurement" for every key.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

RK_STAGES = {"euler": 1, "midpoint": 2, "rk4": 4}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--graphs", type=int, default=4096, help="trajectories (graphs) per GPU")
    ap.add_argument("--solver", default="rk4", choices=list(RK_STAGES))
    ap.add_argument("--engine", default="auto", choices=["auto", "simt", "tc"])
    ap.add_argument("--cpu-graphs", type=int, default=0, help="graphs per CPU-baseline step (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-device-dataset", action="store_true", help="skip the extra leg fed by swarm_ode_b200.dataset")
    ap.add_argument("--no-dopri5", action="store_true", help="skip the dopri5 strong-scaling leg (BASELINE configs[2])")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the benched step")
    ap.add_argument("--no-side", action="store_true", help="skip the small side legs (fp32-upload e2e, small-batch launch count)")
    ap.add_argument("--dopri5-graphs", type=int, default=16384, help="trajectories IN TOTAL of the dopri5 leg (strong scaling)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "bf16_tflops_burst": float(p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "bf16_tflops_burst": 1590.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_traffic(label: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel class, from the committed `ncu --set full`
    capture (profiles/traffic.json, written by scripts/prof_summary.py); None when that class was not captured."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    try:
        return json.load(open(path)).get(label)
    except Exception:
        return None


def kernel_roofline(dom, pk, steps, ms_total):
    """Roofline object of the dominant kernel class.  Which roof binds is decided from the ALGORITHMIC work of the
    launch: HBM time = bytes / measured copy bandwidth; tensor time = 6 x flops / measured bf16 rate for the
    fp32-grade 3xTF32 contraction (three tf32 MMAs per MAC, tf32 issues at half the bf16 rate).  `achieved` is the
    plain algorithmic figure (no 3x / 6x inflation) over the CUDA-event duration measured live in this run."""
    t = dom["ms"] * 1e-3
    hbm_time = dom["bytes"] / (pk["hbm_gbs"] * 1e9)
    # tensor time of the fp32-grade product in bf16-rate units: 3xTF32 = three tf32 MMAs (tf32 issues at half the bf16
    # rate) = 6x; the chain / row-major kernels use two tf32 MMAs + one bf16 MMA per MAC = 5x
    tens_factor = 5.0 if dom["name"].startswith(("chain", "gemm_nt[k128")) else 6.0
    tens_time = tens_factor * dom["flops"] / (pk["bf16_tflops"] * 1e12)
    if hbm_time >= tens_time:
        ach = dom["bytes"] / t / 1e9
        roof = {"kernel": dom["name"], "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": ach / pk["hbm_gbs"], "frac_kernel_bytes": ach / pk["hbm_gbs"], "traffic": measured_traffic(dom["name"]),
                "peak_note": f"{pk['source']} copy bandwidth (MEASURED_PEAKS.json hbm_gbs); `frac` = this kernel's own "
                             "algorithmic bytes (DESIGN.md 4.4) over its measured time; frac_8d_* = SURVEY 8(d)'s Q per unit"}
    else:
        ach = dom["flops"] / t / 1e12
        roof = {"kernel": dom["name"], "bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                "frac": ach / pk["bf16_tflops"], "frac_kernel_bytes": dom["bytes"] / t / 1e9 / pk["hbm_gbs"],
                "traffic": measured_traffic(dom["name"]),
                "peak_note": f"{pk['source']} cuBLAS bf16 sustained; algorithmic fp32 flops of a 3xTF32 contraction "
                             "(a perfect kernel reaches 1/6 of this peak)"}
    roof["hbm_time_ms"], roof["tensor_time_ms"] = hbm_time * 1e3, tens_time * 1e3
    roof["launches_per_step"] = dom["launches"] / steps
    roof["ms_per_launch"] = dom["ms"] / dom["launches"]
    roof["share_of_step"] = dom["ms"] / (ms_total if ms_total > 0 else 1)
    return roof


def algorithmic_bytes_per_unit(D: int, mean_in_degree: float) -> float:
    """Q = 4D (read stage input) + 4D (write stage derivative) + 4*dbar (CSR cols) + 4 (rowptr)."""
    return 8.0 * D + 4.0 * mean_in_degree + 4.0


def algorithmic_flops_per_unit(D: int, H: int) -> float:
    return 8.0 * D * H + 4.0 * H * H


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank: int, world: int):
    """The reference's CPU path (pure-torch restatement of torch_geometric + torchdiffeq, oracle/) timed on
    the box's host cores with all threads.  Rank 0 only."""
    if rank != 0:
        return
    from oracle.train_gde_ref import GraphODERef, train_step_loss_ref
    from oracle.pyg_ref import RefBatch
    import swarm_ode_b200.synthetic as syn

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = args.cpu_graphs or 128
    batch, nxt = syn.warehouse_batch(B, seed=0)
    D = batch.x.shape[1]
    model = GraphODERef(D, 12, 7, hidden_dim=64, ode_solver=args.solver)
    syn.init_weights(model, seed=1, conv3_scale=0.1)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)
    rb = RefBatch(x=batch.x, edge_index=batch.edge_index)
    rb.batch, rb.is_current_agent = batch.batch, batch.is_current_agent
    t = torch.tensor([0.0, 1.0])

    def step():
        opt.zero_grad()
        loss = train_step_loss_ref(model, rb, nxt, t)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
        opt.step()
        return float(loss.detach())

    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    units = batch.x.shape[0] * RK_STAGES[args.solver]
    value = units / dt
    sample = f"{B} graphs ({batch.x.shape[0]} nodes) per step, full train step, {args.steps} steps"
    line = {
        "impl": "reference", "metric": "GNODE agent-state-steps/sec (RK stages)", "value": value,
        "unit": "agent-state-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"GNODE {args.solver} fwd+bwd train step, medium warehouse 12 AGV + 7 pickers, D={D}, "
                               f"{B} trajectories/step (bounded CPU sample of the 4096-trajectory config)",
                   "graphs_per_step": B, "nodes_per_step": batch.x.shape[0], "node_dim": D, "hidden_dim": 64,
                   "solver": args.solver},
        "cpu_baseline": {"value": value, "unit": "agent-state-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "agent-state-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def cpu_baseline(args, D: int):
    """Bounded oracle run (about 10-30 s) on rank 0 at N=1 for the 'cpu_baseline' object."""
    from oracle.train_gde_ref import GraphODERef, train_step_loss_ref
    from oracle.pyg_ref import RefBatch
    import swarm_ode_b200.synthetic as syn

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = args.cpu_graphs or 128
    batch, nxt = syn.warehouse_batch(B, seed=0)
    model = GraphODERef(D, 12, 7, hidden_dim=64, ode_solver=args.solver)
    syn.init_weights(model, seed=1, conv3_scale=0.1)
    rb = RefBatch(x=batch.x, edge_index=batch.edge_index)
    rb.batch, rb.is_current_agent = batch.batch, batch.is_current_agent
    t = torch.tensor([0.0, 1.0])
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)

    def step():
        opt.zero_grad()
        loss = train_step_loss_ref(model, rb, nxt, t)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
        opt.step()

    step()
    t0 = time.perf_counter(); step(); one = time.perf_counter() - t0
    reps = max(2, min(50, int(12.0 / max(one, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    dt = (time.perf_counter() - t0) / reps
    units = batch.x.shape[0] * RK_STAGES[args.solver]
    return {"value": units / dt, "unit": "agent-state-steps/s", "cores": cores, "kind": "port",
            "sample": f"{reps} train steps of {B} graphs ({batch.x.shape[0]} nodes), same solver/weights; "
                      f"pure-torch restatement of the PyG+torchdiffeq reference path",
            "ms_per_step": dt * 1e3}



# ------------------------------------------------------------------------------------------------
def parity_check(model, host, nxt_host, resident, nxt_res, t_dev, dev, n_sub=64, seed=0):
    """Oracle check of the BENCHED configuration: the full 4096-graph step (forward through rk4, decoder, loss,
    backward) runs on the GPU with a loss that only sees `n_sub` randomly chosen graphs; graphs are independent ODE
    systems (disjoint union, scripts/train_gde.py:367), so the solution rows of those graphs and ALL parameter gradients
    must equal the CPU oracle's on the 64-graph sub-batch.  Returns the `parity` object of the JSON line."""
    import torch.nn.functional as F
    from oracle.train_gde_ref import GraphODERef
    from oracle.pyg_ref import RefBatch, RefData
    import swarm_ode_b200 as S

    rng = np.random.default_rng(seed)
    G = int(host.ptr.numel() - 1)
    sub = np.sort(rng.choice(G, size=min(n_sub, G), replace=False))
    ptr = host.ptr.tolist()
    n_agents = int(nxt_host.shape[1])
    # ---- GPU: full batch, loss over the masked rows of the chosen graphs only
    rows = torch.cat([torch.arange(ptr[g], ptr[g + 1]) for g in sub])
    mask_rows = rows[host.is_current_agent[rows]]
    tgt = nxt_host[torch.as_tensor(sub)].reshape(-1, 2)
    model.zero_grad(set_to_none=True)
    resident.__dict__.pop("_gnode_csr", None)
    out = model(resident, t_dev)
    pred = out["trajectories"][1].index_select(0, mask_rows.to(dev))
    loss = F.mse_loss(pred, tgt.to(dev))
    loss.backward()
    torch.cuda.synchronize()
    sol_gpu = out["node_features"].detach()[:, rows.to(dev)].cpu()
    traj_gpu = out["trajectories"].detach()[:, rows.to(dev)].cpu()
    grads_gpu = {n: p.grad.detach().cpu() for n, p in model.named_parameters()}
    # ---- CPU oracle on the sub-batch (fp32, and fp64 to tell arithmetic noise of the oracle itself from a mismatch)
    ei = host.edge_index
    datas = []
    for g in sub:
        n0, n1 = ptr[g], ptr[g + 1]
        sel = (ei[1] >= n0) & (ei[1] < n1)
        datas.append(RefData(x=host.x[n0:n1].clone(), edge_index=ei[:, sel] - n0, is_current_agent=host.is_current_agent[n0:n1].clone()))
    rb = RefBatch.from_data_list(datas)
    D = host.x.shape[1]
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    res = {}
    for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
        ref = GraphODERef(D, 12, 7, hidden_dim=64, ode_solver=model.ode_solver).to(dt)
        ref.load_state_dict({k: v.to(dt) for k, v in sd.items()})
        rbb = RefBatch(x=rb.x.to(dt), edge_index=rb.edge_index)
        rbb.batch, rbb.is_current_agent, rbb.ptr = rb.batch, rb.is_current_agent, rb.ptr
        o = ref(rbb, t_dev.detach().cpu().to(dt))
        l = F.mse_loss(o["trajectories"][1][rbb.is_current_agent], tgt.to(dt))
        l.backward()
        res[name] = (o, float(l), {n: p.grad.detach() for n, p in ref.named_parameters()})

    def rel(a, b):
        a, b = a.double(), b.double()
        d = float(b.norm())
        return float((a - b).norm()) / d if d > 0 else float((a - b).norm())

    o32, l32, g32 = res["f32"]
    _o64, _l64, g64 = res["f64"]
    e_sol = rel(sol_gpu, o32["node_features"].detach())
    e_traj = rel(traj_gpu, o32["trajectories"].detach())
    e_loss = abs(float(loss) - l32) / max(abs(l32), 1e-30)
    e_grad = {n: rel(grads_gpu[n], g64[n]) for n in grads_gpu}
    e_grad_oracle32 = {n: rel(g32[n], g64[n]) for n in grads_gpu}
    worst_grad = max(e_grad, key=e_grad.get)
    par = {"n_graphs": int(len(sub)), "of_graphs": G, "max_rel_l2": max(e_sol, e_traj), "node_features_rel_l2": e_sol,
           "trajectories_rel_l2": e_traj, "loss_rel_err": e_loss, "grad_max_rel_l2": e_grad[worst_grad], "grad_worst": worst_grad,
           "grad_max_rel_l2_fp32_oracle_vs_fp64": max(e_grad_oracle32.values()),
           "tolerance": {"solution": 1e-4, "gradients": 1e-3},
           "how": "full-size GPU step with a loss restricted to a random subset of graphs vs the CPU oracle on that subset "
                  "(fp32 for the solution, fp64 for the gradients)"}
    par["ok"] = bool(par["max_rel_l2"] <= 1e-4 and par["grad_max_rel_l2"] <= 1e-3 and e_loss <= 1e-4)
    model.zero_grad(set_to_none=True)
    return par


def small_batch_leg(dev, solver, sizes=(32, 64, 256), iters=30):
    """The reference's own default regime (`batch_size: 32`, scripts/train_gde.py:437-445): device time and kernel
    launches of one full train step at small batches, where launch count, not bandwidth, sets the pace."""
    import swarm_ode_b200 as S
    from swarm_ode_b200.dist import masked_mse_train_step
    from swarm_ode_b200 import graph as G
    rows = []
    for B in sizes:
        batch, nxt = S.synthetic.warehouse_batch(B, seed=7)
        D = batch.x.shape[1]
        model = S.GraphODE(D, 12, 7, hidden_dim=64, ode_solver=solver)
        S.synthetic.init_weights(model, seed=1, conv3_scale=0.1)
        model = model.to(dev)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=True)
        b, nx = batch.to(dev), nxt.to(dev)
        t = torch.tensor([0.0, 1.0], device=dev)

        def step():
            G.clear_cache()
            b.__dict__.pop("_gnode_csr", None)
            return masked_mse_train_step(model, opt, b, nx, t, distributed=False)   # rank-0-only leg: no collective
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        l0 = S.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for _ in range(iters):
            step()
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - w0) / iters * 1e3
        ms = e0.elapsed_time(e1) / iters
        units = b.x.shape[0] * RK_STAGES[solver]
        row = {"graphs": B, "nodes": int(b.x.shape[0]), "ms_per_step": ms, "wall_ms_per_step": wall,
               "library_launches_per_step": (S.launch_count() - l0) / iters, "value": units / (ms * 1e-3)}
        # the same step captured once into a CUDA graph (swarm_ode_b200/graphed.py): ONE launch per batch, the batch
        # copied into the graph's static buffers inside the timed loop (a fresh batch per step, as in training)
        try:
            from swarm_ode_b200.graphed import GraphedTrainStep
            model_g = S.GraphODE(D, 12, 7, hidden_dim=64, ode_solver=solver)
            S.synthetic.init_weights(model_g, seed=1, conv3_scale=0.1)
            model_g = model_g.to(dev)
            opt_g = torch.optim.Adam(model_g.parameters(), lr=1e-3, weight_decay=1e-4, fused=True, capturable=True)
            gs = GraphedTrainStep(model_g, opt_g, b, nx, t)
            for _ in range(3):
                gs.step(b, nx)
            torch.cuda.synchronize()
            w0 = time.perf_counter()
            e0.record()
            for _ in range(iters):
                gs.step(b, nx)
            e1.record()
            torch.cuda.synchronize()
            gs.check()
            gms = e0.elapsed_time(e1) / iters
            row.update({"graphed_ms_per_step": gms, "graphed_wall_ms_per_step": (time.perf_counter() - w0) / iters * 1e3,
                        "graphed_launches_per_step": 1, "graphed_value": units / (gms * 1e-3)})
        except Exception as e:   # reported, never fatal for the headline
            row["graphed_error"] = f"{type(e).__name__}: {e}"[:300]
        rows.append(row)
    return {"unit": "agent-state-steps/s", "solver": solver, "rows": rows,
            "note": "full train step (CSR build, forward, loss, backward, clip, Adam) per batch: eager launches, and the "
                    "same step replayed as one CUDA graph (graphed_*: swarm_ode_b200.graphed.GraphedTrainStep)"}


DOPRI5_CONV3_SCALE = 8.0
DOPRI5_TIMES = (0.0, 1.0, 2.0, 3.0)


def dopri5_strong(args, rank, world, dev, pk, steps=3, warmup=1):
    """BASELINE configs[2] under the driver: adaptive dopri5 (rtol 1e-3, atol 1e-4) over partial-observation graphs of
    19 AGVs + 9 pickers (D = 435, 140 nodes per graph), `--dopri5-graphs` trajectories IN TOTAL sharded over the ranks
    (strong scaling); the error norm is global over the whole batch, so all ranks take the unsharded batch's step
    decisions.  conv3 weights x 8 and t = 0..3 with three output points: 13 accepted steps (with the SURVEY's 0.1 scale the
    solve needs two steps and the controller is not exercised).  No attempt is rejected at this batch size -- the global RMS
    norm averages over 10^9 elements of a smooth field (peak error ratio 0.77); rejected attempts are covered by the
    small-batch regimes of tests/test_gpu_integrate.py, with accept / reject lists identical to the oracle's."""
    import swarm_ode_b200 as S
    from swarm_ode_b200 import dist as Dm
    lo, hi = Dm.shard_range(args.dopri5_graphs, rank, world)
    batch, _ = S.synthetic.warehouse_batch(hi - lo, num_agvs=19, num_pickers=9, seed=1000 + rank)
    Dn = batch.x.shape[1]
    model = S.GraphODE(Dn, 19, 9, hidden_dim=64, ode_solver="dopri5")
    S.synthetic.init_weights(model, seed=1, conv3_scale=DOPRI5_CONV3_SCALE)
    model = model.to(dev)
    model.dopri5_allreduce = Dm.dopri5_norm_allreduce(device=dev)
    b = batch.to(dev)
    E = int(batch.edge_index.shape[1])
    del batch
    t = torch.tensor(DOPRI5_TIMES, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(warmup):
            model(b, t)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            model(b, t)
        e1.record()
        barrier()
    st = model.last_stats
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    cnt = torch.tensor([float(b.x.shape[0]), float(E)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    from swarm_ode_b200 import _lib
    _lib.tc_check(dev)
    nodes, edges = float(cnt[0]), float(cnt[1])
    units = nodes * st.nfe
    Q = algorithmic_bytes_per_unit(Dn, edges / nodes)
    value = units / (float(ms) * 1e-3)
    del b, model
    torch.cuda.empty_cache()
    return {"value": value, "unit": "agent-state-steps/s", "ms_per_solve": float(ms), "scaling": "strong", "n_gpus": world,
            "graphs_total": args.dopri5_graphs, "nodes_total": int(nodes), "node_dim": Dn, "nodes_per_graph": 140,
            "rtol": 1e-3, "atol": 1e-4, "t": list(DOPRI5_TIMES), "conv3_scale": DOPRI5_CONV3_SCALE,
            "nfe": st.nfe, "accepted": st.n_accepted, "attempted": st.n_attempted,
            "bytes_per_unit": Q, "hbm_frac_8d": value * Q / world / 1e9 / pk["hbm_gbs"],
            "steps": steps, "warmup": warmup,
            "workload": "GNODE dopri5 forward (BASELINE configs[2]): 19 AGV + 9 pickers partial-obs graphs, "
                        f"{args.dopri5_graphs} trajectories in total over {world} GPU(s), global error norm (two doubles "
                        "all-reduced per attempted step)"}


# ------------------------------------------------------------------------------------------------
def run_ours(args, rank: int, world: int, local_rank: int):
    import swarm_ode_b200 as S
    from swarm_ode_b200 import _lib
    from swarm_ode_b200.dist import masked_mse_train_step
    from swarm_ode_b200 import graph as G

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path)"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    S.set_engine(args.engine)

    # ---- synthetic batch: pinned on the host (e2e leg copies it every step), resident copy for `value`
    host, nxt_host = S.synthetic.warehouse_batch(args.graphs, seed=rank)
    host.pin_memory()
    nxt_host = nxt_host.pin_memory()
    # transport format of the same batch (lossless: u8 node features verified element by element at pack time, int32
    # edges, `batch` derived from `ptr` on the device) -- what the e2e leg uploads every step
    packed = S.PackedBatch(host, nxt_host).pin_memory()
    D = host.x.shape[1]
    H = 64
    N_nodes, E = host.x.shape[0], host.edge_index.shape[1]
    stages = RK_STAGES[args.solver]
    units_per_step_rank = N_nodes * stages

    model = S.GraphODE(D, 12, 7, hidden_dim=H, ode_solver=args.solver)
    S.synthetic.init_weights(model, seed=1, conv3_scale=0.1)
    model = model.to(dev)
    # same update rule as the reference's torch.optim.Adam(lr, weight_decay) (scripts/train_gde.py:459); fused=True runs it as
    # one multi-tensor kernel instead of ~10 (host glue, not part of the library)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=True)
    t_dev = torch.tensor([0.0, 1.0], device=dev)

    def to_device(non_blocking=True):
        b = S.Batch(x=host.x.to(dev, non_blocking=non_blocking), edge_index=host.edge_index.to(dev, non_blocking=non_blocking))
        b.batch = host.batch.to(dev, non_blocking=non_blocking)
        b.is_current_agent = host.is_current_agent.to(dev, non_blocking=non_blocking)
        b.ptr = host.ptr.to(dev, non_blocking=non_blocking)          # graph offsets (PyG Batch.ptr)
        b.num_graphs, b.max_graph_nodes = host.num_graphs, host.max_graph_nodes
        return b, nxt_host.to(dev, non_blocking=non_blocking)

    resident, nxt_res = to_device(False)

    def step_resident():
        G.clear_cache()
        resident.__dict__.pop("_gnode_csr", None)      # a new batch every step: CSR is rebuilt inside the step
        return masked_mse_train_step(model, opt, resident, nxt_res, t_dev)

    copy_stream = torch.cuda.Stream(device=dev)
    pending = {}

    use_packed = [True]

    def upload():
        """Enqueue the host -> device copy of one batch (pinned memory) on the copy stream: the packed transport format
        (widened to the reference's fp32 / int64 tensors by the library's unpack kernels, same stream) or, for the
        comparison leg, the plain fp32 / int64 tensors."""
        with torch.cuda.stream(copy_stream):
            b, nx = packed.to(dev, non_blocking=True) if use_packed[0] else to_device(True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        pending["next"] = (b, nx, ev)

    def step_e2e():
        """One step through the public API with HOST inputs: this step's batch was uploaded while the previous step
        computed (a prefetching loader); the next step's upload is enqueued before this step's compute so the two
        overlap.  Every step still pays one full H2D copy of its inputs and one D2H read of the loss."""
        G.clear_cache()
        b, nx, ev = pending.pop("next")
        torch.cuda.current_stream(dev).wait_event(ev)
        upload()
        loss = masked_mse_train_step(model, opt, b, nx, t_dev)
        for t in (b.x, b.edge_index, b.batch, b.is_current_agent, b.ptr, nx):
            t.record_stream(torch.cuda.current_stream(dev))
        # device -> host read of the step's result: an asynchronous copy into pinned memory every step; the value is
        # consumed one step later (like a training loop that logs the previous step's loss), so the host never
        # drains the GPU queue and a host-side hiccup does not stall the device
        slot = loss_ring[step_e2e.i % 2]
        slot[0].copy_(loss.detach().reshape(1), non_blocking=True)
        slot[1].record(torch.cuda.current_stream(dev))
        step_e2e.i += 1
        prev = loss_ring[step_e2e.i % 2]
        if step_e2e.i >= 2:
            prev[1].synchronize()
            return float(prev[0][0])
        return None
    step_e2e.i = 0
    loss_ring = [(torch.empty(1, dtype=torch.float32).pin_memory(), torch.cuda.Event()) for _ in range(2)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM ----
    n_warm = max(args.warmup, 5)   # >= 3 required; two more let the caching allocator settle (no cudaMalloc when timed)
    # clocks / throttle reasons: nvidia-smi needs ~0.2 s to deliver its first sample and the timed region is ~60 ms, so the
    # sampler runs from the warm-up steps through the timed region and a tail of identical untimed steps (same load)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(n_warm):
        step_resident()
    barrier()
    launches0 = S.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_resident()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = S.launch_count() - launches0
    for _ in range(80):            # untimed tail under the same load (every rank: the step holds a collective at N > 1)
        step_resident()
    barrier()
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["window"] = "warm-up + timed region + 80 identical untimed steps (nvidia-smi -lms 100)"
    # per-kernel-class device time (CUDA events on the launching stream around every launch): a SEPARATE pass over
    # the same steps, so that the event records do not sit inside the timed region above
    prof_steps = min(args.steps, 5)
    _lib.prof_enable(True)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for _ in range(prof_steps):
        step_resident()
    pe1.record()
    barrier()
    prof_ms_total = pe0.elapsed_time(pe1)
    prof = _lib.prof_read()
    # forward kernels only (solver + decoder, no loss / backward): the denominator of SURVEY 8(d)'s per-stage figure
    _lib.prof_enable(True)
    with torch.no_grad():
        for _ in range(prof_steps):
            G.clear_cache()
            resident.__dict__.pop("_gnode_csr", None)
            model(resident, t_dev)
    barrier()
    prof_fwd = _lib.prof_read()
    _lib.prof_enable(False)
    _lib.tc_check(dev)          # a barrier timeout inside the timed region invalidates the numbers: raise, do not print
    G.poll_pending()
    tmax = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = float(tmax) / args.steps
    value = units_per_step_rank * world / (ms_step * 1e-3)

    # ---- e2e: host buffers in, loss out, every step ----
    def run_e2e(packed_transport: bool):
        use_packed[0] = packed_transport
        step_e2e.i = 0
        upload()
        for _ in range(4):
            step_e2e()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_wall = []
        for _ in range(args.steps):
            w0 = time.perf_counter()
            step_e2e()
            e2e_wall.append((time.perf_counter() - w0) * 1e3)
        for slot in loss_ring:          # the last two losses are read before the clock stops
            slot[1].synchronize()
            float(slot[0][0])
        e1.record()
        barrier()
        pending.clear()
        if rank == 0:
            print(f"e2e ({'packed' if packed_transport else 'fp32'}) per-step wall ms: " + " ".join(f"{w:.1f}" for w in e2e_wall), file=sys.stderr)
        te = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return float(te) / args.steps

    e2e_ms = run_e2e(True)
    e2e_value = units_per_step_rank * world / (e2e_ms * 1e-3)
    h2d = packed.nbytes
    h2d_fp32 = sum(t.numel() * t.element_size() for t in (host.x, host.edge_index, host.batch, host.is_current_agent, host.ptr, nxt_host))
    e2e_fp32 = None
    if not args.no_side:
        ms32 = run_e2e(False)
        e2e_fp32 = {"value": units_per_step_rank * world / (ms32 * 1e-3), "unit": "agent-state-steps/s", "ms_per_step": ms32,
                    "h2d_bytes_per_step": h2d_fp32, "note": "same step, the batch uploaded as plain fp32 / int64 tensors (round-1 e2e)"}
    _lib.tc_check(dev)
    G.poll_pending()

    # ---- the same step fed by the repository's own loader (swarm_ode_b200/dataset.py): the episodes' window graphs were
    # built on the GPU and live there, a step uploads only its sample indices and gathers whole graphs on the device.
    # Reported next to `e2e` (which uploads the full fp32 batch every step, as the reference's DataLoader does), not
    # instead of it.
    dd = None
    if not args.no_device_dataset:
        from swarm_ode_b200.dataset import WarehouseDataset, synthetic_episode
        n_ep, ep_len = 12, 401
        ds = WarehouseDataset([synthetic_episode(ep_len, 12, 7, seed=100 * rank + e) for e in range(n_ep)], dev)
        rng = np.random.default_rng(rank)
        full = np.arange(len(ds))
        full = full[ds._ptr[ds._samples[full] + 1] - ds._ptr[ds._samples[full]] == 95]      # full 5-snapshot windows only
        def step_ds():
            G.clear_cache()
            idx = rng.choice(full, size=args.graphs, replace=len(full) < args.graphs)
            tb = ds.collate(idx)
            return masked_mse_train_step(model, opt, tb.graphs, tb.next_positions, t_dev)
        for _ in range(3):
            step_ds()
        barrier()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        last = None
        for _ in range(args.steps):
            last = step_ds()
        float(last)                        # device -> host read of the final loss inside the timed region
        d1.record()
        barrier()
        td = torch.tensor([d0.elapsed_time(d1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
        dd_ms = float(td) / args.steps
        dd = {"value": units_per_step_rank * world / (dd_ms * 1e-3), "unit": "agent-state-steps/s", "ms_per_step": dd_ms,
              "h2d_bytes_per_step": int(args.graphs * 8 * 6), "note": "device-resident dataset (window graphs built on the GPU once), "
              "per step: sample indices up, device-side collation of 4096 graphs, train step, loss down"}
        del ds

    # ---- oracle check of the benched configuration (rank 0; the other ranks wait at the next barrier) ----
    parity = None
    if rank == 0 and not args.no_parity:
        parity = parity_check(model, host, nxt_host, resident, nxt_res, t_dev, dev)
        print("parity: " + json.dumps(parity), file=sys.stderr)

    # ---- small-batch regime: kernel launches of one train step at the reference's default batch sizes ----
    small = None
    if rank == 0 and not args.no_side:
        small = small_batch_leg(dev, args.solver)

    # free the headline workload before the dopri5 leg allocates its own
    del resident, nxt_res
    pending.clear()
    torch.cuda.empty_cache()
    pk = peaks()
    d5 = None
    if not args.no_dopri5:
        barrier()
        d5 = dopri5_strong(args, rank, world, dev, pk)

    if rank != 0:
        return
    # ---- roofline of the dominant kernel class (device time measured live with CUDA events) ----
    prof = [p for p in prof if p["launches"] > 0 and not p["name"].startswith(("csr_build", "decoder_bwd"))]
    prof.sort(key=lambda p: -p["ms"])
    dom = prof[0] if prof else None
    roof = kernel_roofline(dom, pk, prof_steps, prof_ms_total) if dom else None
    dbar = E / N_nodes
    Q = algorithmic_bytes_per_unit(D, dbar)
    F = algorithmic_flops_per_unit(D, H)
    # units x Q over the summed device time of the FORWARD kernels (CSR / tile build excluded: once per
    # batch, not per stage), and 3Q per unit over the whole step
    fwd_k = [p for p in prof_fwd if p["launches"] > 0 and not p["name"].startswith(("csr_build", "tiles"))]
    fwd_ms = sum(p["ms"] for p in fwd_k) / prof_steps if fwd_k else None
    if roof is not None and fwd_ms:
        roof["frac_8d_forward"] = units_per_step_rank * Q / (fwd_ms * 1e-3) / 1e9 / pk["hbm_gbs"]
        roof["forward_ms"] = fwd_ms
        roof["forward_kernels"] = {p["name"]: round(p["ms"] / prof_steps, 4) for p in sorted(fwd_k, key=lambda p: -p["ms"])[:8]}
        roof["frac_8d_step"] = units_per_step_rank * Q * 3.0 / (ms_step * 1e-3) / 1e9 / pk["hbm_gbs"]
        roof["bytes_per_unit_8d"] = Q
    # forward units + backward (F_bwd = 2F, Q_bwd = 2Q: SURVEY 8-d4)
    step_bytes = units_per_step_rank * Q * 3.0
    step_flops = units_per_step_rank * F * 3.0
    step_roof = {"hbm_GBps_algorithmic": step_bytes / (ms_step * 1e-3) / 1e9,
                 "hbm_frac": step_bytes / (ms_step * 1e-3) / 1e9 / pk["hbm_gbs"],
                 "tflops_algorithmic": step_flops / (ms_step * 1e-3) / 1e12,
                 "bytes_per_unit": Q, "flops_per_unit": F, "fwd_plus_bwd_factor": 3.0, "peak_source": pk["source"]}
    kernels = [{"name": p["name"], "launches": p["launches"], "ms": round(p["ms"], 4),
                "share": round(p["ms"] / prof_ms_total, 4) if prof_ms_total else None,
                "TFLOPs": round(p["flops"] / (p["ms"] * 1e-3) / 1e12, 3) if p["ms"] > 0 else None,
                "GBps": round(p["bytes"] / (p["ms"] * 1e-3) / 1e9, 1) if p["ms"] > 0 else None} for p in prof[:12]]

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args, D)

    line = {
        "metric": "GNODE agent-state-steps/sec (RK stages)", "value": value, "unit": "agent-state-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": n_warm, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"GNODE {args.solver} fwd+bwd train step (BASELINE configs[1]): medium warehouse, 12 AGV + 7 pickers, "
                               f"D={D}, 95 nodes/graph, {args.graphs} trajectories per GPU",
                   "graphs_per_gpu": args.graphs, "nodes_per_gpu": N_nodes, "edges_per_gpu": E, "node_dim": D,
                   "hidden_dim": H, "solver": args.solver, "rk_stages": stages, "engine": args.engine,
                   "parallelism": f"dp{world} (graphs sharded, gradient all-reduce only)",
                   "l2": "inputs (x = %.0f MB per GPU) exceed the 126 MB L2; no explicit flush" % (host.x.numel() * 4 / 1e6),
                   "csr": "rebuilt every step (new batch each step)"},
        "e2e": {"value": e2e_value, "unit": "agent-state-steps/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4,
                "transport": f"PackedBatch: x as {packed.kind_name} (exact round trip verified at pack time), edge_index int32, "
                             "batch vector derived from ptr on the device; widened to the reference's fp32 / int64 tensors "
                             "by gnode_unpack_* inside the timed region",
                "h2d_bytes_per_step_fp32": h2d_fp32},
        "e2e_fp32_upload": e2e_fp32,
        "parity": parity,
        "dopri5_strong": d5,
        "small_batch": small,
        "e2e_device_dataset": dd,
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roof,
        "step_roofline": step_roof,
        "kernels": kernels,
        "cpu_baseline": cpu,
    }
    emit(line)


_JSON_FD = None


def emit(line: dict):
    """The ONE JSON line of the contract, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    args = parse()
    # Libraries (NCCL prints its version banner on stdout) must not add lines to stdout: everything but the JSON line
    # goes to stderr.
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1 and dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
