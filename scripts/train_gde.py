#!/usr/bin/env python
"""Entry point of the reference's GNODE training (scripts/train_gde.py:430-534) on the B200-native path.

    python scripts/train_gde.py --data warehouse_data_*_seed0.npz [...]          # shards from scripts/h5_to_npz.py
    python scripts/train_gde.py --synthetic 8 --steps-per-episode 200             # no data at hand: synthetic episodes
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/train_gde.py ...

Same procedure as the reference: (current window graph, next positions) pairs, 80/20 random split, batches of
``--batch-size`` graphs, ``GraphODE(node_dim, hidden_dim=64, ode_solver='euler')``, Adam(lr, weight_decay), masked MSE on the
decoded positions at t = 1, gradient clipping at 1.0, validation every epoch, ``best_model.pth`` (a state_dict with the
reference's keys) and periodic checkpoints under ``--save-dir``.  What differs is where things run: the episodes' window
graphs are built on the GPU and stay there, a batch is a device-side gather, and with several ranks every rank takes a
shard of each batch (gradients all-reduced once per step).  Logging goes to stdout (wandb only with ``--wandb``)."""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import swarm_ode_b200 as S  # noqa: E402
from swarm_ode_b200.dataset import WarehouseDataset, split_indices, synthetic_episode  # noqa: E402
from swarm_ode_b200.dist import masked_mse_train_step  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", nargs="*", default=[], help=".npz shards (scripts/h5_to_npz.py) or .h5 files of collect_data.py")
    ap.add_argument("--synthetic", type=int, default=0, help="number of synthetic episodes when no --data is given")
    ap.add_argument("--steps-per-episode", type=int, default=200)
    ap.add_argument("--num-agvs", type=int, default=12)
    ap.add_argument("--num-pickers", type=int, default=7)
    ap.add_argument("--num-epochs", type=int, default=200)
    ap.add_argument("--batch-size", type=int, default=32)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--weight-decay", type=float, default=1e-4)
    ap.add_argument("--hidden-dim", type=int, default=64)
    ap.add_argument("--ode-solver", default="euler", choices=["euler", "midpoint", "rk4", "dopri5"])
    ap.add_argument("--save-dir", default="./trained_models")
    ap.add_argument("--checkpoint-every", type=int, default=50)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--wandb", action="store_true")
    ap.add_argument("--cuda-graph", action="store_true",
                    help="single GPU, fixed-grid solver: replay the training step as ONE CUDA graph (swarm_ode_b200.GraphedTrainStep); "
                         "at the reference's batch size of 32 the eager step is bound by ~100 launches, not by the GPU")
    return ap.parse_args()


def main():
    args = parse()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    if not torch.cuda.is_available():
        raise SystemExit("scripts/train_gde.py needs a CUDA device: libgnode_b200 has no CPU path")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(args.seed)

    t0 = time.perf_counter()
    if args.data:
        ds = WarehouseDataset.from_files(args.data, dev, distance_threshold=5.0, temporal_window=5)
    elif args.synthetic > 0:
        eps = [synthetic_episode(args.steps_per_episode, args.num_agvs, args.num_pickers, seed=args.seed + e) for e in range(args.synthetic)]
        ds = WarehouseDataset(eps, dev, distance_threshold=5.0, temporal_window=5)
    else:
        raise SystemExit("pass --data files or --synthetic N")
    train_idx, val_idx = split_indices(len(ds), 0.8, seed=args.seed)
    if rank == 0:
        print(f"Loaded {len(ds)} step pairs ({ds.x.shape[0]} window-graph nodes resident on {dev}) in {time.perf_counter() - t0:.1f} s")
        print(f"Node dimension: {ds.node_dim}\nAgents: {ds.num_agvs} AGVs, {ds.num_pickers} Pickers")
        print(f"Dataset sizes - Train: {len(train_idx)}, Val: {len(val_idx)}")

    model = S.GraphODE(node_dim=ds.node_dim, num_agvs=ds.num_agvs, num_pickers=ds.num_pickers, hidden_dim=args.hidden_dim,
                       ode_solver=args.ode_solver).to(dev)
    if world > 1:   # identical initial weights on every rank
        for p in model.parameters():
            dist.broadcast(p.data, src=0)
    use_graph = args.cuda_graph and world == 1 and args.ode_solver != "dopri5"
    opt = torch.optim.Adam(model.parameters(), lr=args.lr, weight_decay=args.weight_decay, fused=True, capturable=use_graph)
    graphed = None
    save_dir = os.path.join(args.save_dir, time.strftime("%Y%m%d_%H%M%S"))
    if rank == 0:
        os.makedirs(save_dir, exist_ok=True)
    run = None
    if args.wandb and rank == 0:
        import wandb
        run = wandb.init(project="graph-ode-warehouse", config=vars(args))
    time_span = torch.tensor([0.0, 1.0], device=dev)
    rng = np.random.default_rng(args.seed + 1)
    best = float("inf")

    def shard(idx):          # every rank takes a contiguous part of the batch's graphs
        lo, hi = (len(idx) * rank) // world, (len(idx) * (rank + 1)) // world
        return idx[lo:hi]

    for epoch in range(args.num_epochs):
        model.train()
        perm = rng.permutation(train_idx)
        tot, nb = torch.zeros((), device=dev), 0
        for b0 in range(0, len(perm), args.batch_size):
            mine = shard(perm[b0:b0 + args.batch_size])
            if world > 1 and len(perm[b0:b0 + args.batch_size]) < world:
                continue
            batch = ds.collate(mine)
            if use_graph and graphed is None and len(mine) == args.batch_size:
                # captured on the first full batch; later batches are copied into its static buffers (a batch of another size,
                # e.g. the last one of an epoch, takes the eager step)
                graphed = S.GraphedTrainStep(model, opt, batch.graphs, batch.next_positions, time_span)
            if graphed is not None:
                tot += graphed.step(batch.graphs, batch.next_positions)
            else:
                tot += masked_mse_train_step(model, opt, batch.graphs, batch.next_positions, time_span)
            nb += 1
        train_loss = float(tot) / max(nb, 1)
        model.eval()
        vtot, vn = torch.zeros((), device=dev, dtype=torch.float64), torch.zeros((), device=dev, dtype=torch.float64)
        with torch.no_grad():
            for b0 in range(0, len(val_idx), args.batch_size):
                mine = shard(val_idx[b0:b0 + args.batch_size])
                if len(mine) == 0:
                    continue
                batch = ds.collate(mine)
                pred = model(batch.graphs, time_span)["trajectories"][1]
                tgt = batch.next_positions.view(-1, 2)
                err = (pred[batch.graphs.is_current_agent] - tgt).pow(2).sum()
                vtot += err.double(); vn += tgt.numel()
        if world > 1:
            dist.all_reduce(vtot); dist.all_reduce(vn)
        val_loss = float(vtot / vn.clamp_min(1.0))
        # once per epoch: the deferred device-side verdicts (edge-list / tile checks) and the tcgen05 barrier status word;
        # raises GnodeError instead of training on with invalid results
        S.graph.poll_pending()
        S._lib.tc_check(dev)
        if graphed is not None:
            graphed.check()
        if rank == 0:
            if val_loss < best:
                best = val_loss
                torch.save(model.state_dict(), os.path.join(save_dir, "best_model.pth"))
                print(f"Saved best model at epoch {epoch} with val loss {best:.6f}")
            if epoch % args.checkpoint_every == 0:
                torch.save(model.state_dict(), os.path.join(save_dir, f"checkpoint_epoch{epoch}.pth"))
            print(f"Epoch {epoch:3d} | Train Loss: {train_loss:.6f} | Val Loss: {val_loss:.6f}", flush=True)
            if run is not None:
                run.log({"epoch": epoch, "train_loss": train_loss, "val_loss": val_loss})
    if world > 1:
        dist.destroy_process_group()
    return best


if __name__ == "__main__":
    main()
