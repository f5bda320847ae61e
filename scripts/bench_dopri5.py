#!/usr/bin/env python
"""BASELINE.json configs[2]: adaptive dopri5 integration (rtol 1e-3, atol 1e-4, t = [0, 1]) of the GNODE over partial-
observation graphs of 19 AGVs + 9 pickers (D = 435, 140 nodes per graph), 16k trajectories sharded over the ranks.

    python scripts/bench_dopri5.py [--graphs-total 16384] [--steps 5] [--warmup 2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_dopri5.py

The reference's error norm is global over the whole batch tensor, so every attempted step all-reduces (sum of squares,
count) across ranks and all ranks take the unsharded batch's step decisions (swarm_ode_b200/dist.py).  Strong scaling:
the total is fixed.  Prints one JSON line (rank 0): units = nodes x vector-field evaluations ("agent-state-steps").
This is a side bench; `bench.py` at the repo root carries the contract line for configs[1]."""
import argparse, json, os, sys, time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import swarm_ode_b200 as S  # noqa: E402
from swarm_ode_b200 import dist as D  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--graphs-total", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    out_fd = os.dup(1)
    os.dup2(2, 1)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lo, hi = D.shard_range(args.graphs_total, rank, world)
    batch, _ = S.synthetic.warehouse_batch(hi - lo, num_agvs=19, num_pickers=9, seed=1000 + rank)
    Dn = batch.x.shape[1]
    model = S.GraphODE(Dn, 19, 9, hidden_dim=64, ode_solver="dopri5")
    S.synthetic.init_weights(model, seed=1, conv3_scale=0.1)
    model = model.to(dev)
    model.dopri5_allreduce = D.dopri5_norm_allreduce(device=dev)
    b = batch.to(dev)
    t = torch.tensor([0.0, 1.0], device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(args.warmup):
            model(b, t)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            model(b, t)
        e1.record()
        barrier()
        wall = (time.perf_counter() - w0) / args.steps * 1e3
    st = model.last_stats
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
    nodes = torch.tensor([float(b.x.shape[0])], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(nodes, op=dist.ReduceOp.SUM)
    if rank == 0:
        units = float(nodes) * st.nfe
        line = {"metric": "GNODE agent-state-steps/sec (RK stages)", "value": units / (float(ms) * 1e-3), "unit": "agent-state-steps/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(ms), "wall_ms_per_step": wall,
                "higher_is_better": True, "scaling": "strong", "dtype": "f32", "data": "synthetic",
                "config": {"workload": "GNODE dopri5 forward (BASELINE configs[2]): 19 AGV + 9 pickers partial-obs graphs, "
                                       f"D={Dn}, 140 nodes/graph, {args.graphs_total} trajectories in total, rtol 1e-3 atol 1e-4, t=[0,1]",
                           "nodes_total": int(float(nodes)), "nfe": st.nfe, "accepted": st.n_accepted, "attempted": st.n_attempted,
                           "parallelism": f"dp{world} (graphs sharded; two doubles all-reduced per error norm)",
                           "path": "folded stages in the graph-resident chain kernel, one 140-node graph per 144-row tile (two 128-row blocks)"}}
        os.write(out_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
