#!/usr/bin/env python
"""Runnable stand-in for the reference's heterogeneous GNODE entry point (scripts/run_gnode.py:1328-1530) on the native path.

The reference script cannot run as shipped: its module level needs wandb, a gymnasium TA-RWARE environment and the QMIX
agent, and its driver loop fails before the model is used.  What belongs to the hot path is kept and runs here
without the simulator: `MultiAgentGraphConverter` turns joint observations of the 19 AGV + 9 picker warehouse into
`HeteroData`, `HeteroGraphODENetwork(node_dims, action_size, hidden_dim=128)` (pre-GNN, per-type MLP ODEs with 'euler',
Q heads) scores them, and a TD(0) update with a target network trains it -- the per-agent part of what the reference's
agents do with the network's `agv_q_values` / `picker_q_values`; the QMIX mixer, replay buffer and environment are out of
scope.  Observations are synthetic (`swarm_ode_b200.synthetic.multi_agent_observation`).

    python scripts/run_gnode.py --steps 50 --batch-size 128
"""
import argparse
import copy
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import swarm_ode_b200 as S  # noqa: E402
from swarm_ode_b200.hetero import HeteroData, HeteroGraphODENetwork, MultiAgentGraphConverter  # noqa: E402
from swarm_ode_b200.synthetic import multi_agent_observation, rack_locations  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-agvs", type=int, default=19)
    ap.add_argument("--num-pickers", type=int, default=9)
    ap.add_argument("--locations", type=int, default=160)
    ap.add_argument("--action-size", type=int, default=161, help="env.unwrapped.action_size of the reference (racks + 1)")
    ap.add_argument("--hidden-dim", type=int, default=128)
    ap.add_argument("--batch-size", type=int, default=128)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--gamma", type=float, default=0.999)
    ap.add_argument("--target-every", type=int, default=20)
    ap.add_argument("--seed", type=int, default=0)
    return ap.parse_args()


def joint_batch(rng, conv, racks, args, dev):
    """`batch_size` joint observations -> one batched HeteroData (disjoint union, HeteroData.from_data_list)."""
    graphs = [conv._build_graph_from_observation(multi_agent_observation(rng, args.num_agvs, args.num_pickers, racks,
                                                                         agv_target_prob=0.3), racks)
              for _ in range(args.batch_size)]
    return HeteroData.from_data_list(graphs).to(dev)


def main():
    args = parse()
    if not torch.cuda.is_available():
        raise SystemExit("scripts/run_gnode.py needs a CUDA device: libgnode_b200 has no CPU path")
    dev = torch.device("cuda:0")
    torch.manual_seed(args.seed)
    rng = np.random.default_rng(args.seed)
    node_dims = {"agv": 7, "picker": 4, "location": 2}            # scripts/run_gnode.py:1339-1343
    net = HeteroGraphODENetwork(node_dims, args.action_size, hidden_dim=args.hidden_dim).to(dev)
    target = copy.deepcopy(net)
    opt = torch.optim.Adam(net.parameters(), lr=args.lr)
    conv = MultiAgentGraphConverter(num_agvs=args.num_agvs, num_pickers=args.num_pickers, fresh=True)
    racks = rack_locations(rng, args.locations)
    n_agents = args.num_agvs + args.num_pickers
    t_build = t_step = 0.0
    for step in range(args.steps):
        h0 = time.perf_counter()
        cur, nxt = joint_batch(rng, conv, racks, args, dev), joint_batch(rng, conv, racks, args, dev)
        t_build += time.perf_counter() - h0
        act_a = torch.randint(0, args.action_size, (args.batch_size * args.num_agvs, 1), device=dev)
        act_p = torch.randint(0, args.action_size, (args.batch_size * args.num_pickers, 1), device=dev)
        rew = torch.rand(args.batch_size, device=dev)
        torch.cuda.synchronize()
        h0 = time.perf_counter()
        out = net(cur)
        with torch.no_grad():
            tout = target(nxt)
            ya = rew.repeat_interleave(args.num_agvs) + args.gamma * tout["agv_q_values"].max(dim=1).values
            yp = rew.repeat_interleave(args.num_pickers) + args.gamma * tout["picker_q_values"].max(dim=1).values
        qa = out["agv_q_values"].gather(1, act_a).squeeze(1)
        qp = out["picker_q_values"].gather(1, act_p).squeeze(1)
        loss = torch.nn.functional.smooth_l1_loss(qa, ya) + torch.nn.functional.smooth_l1_loss(qp, yp)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 10.0)
        opt.step()
        torch.cuda.synchronize()
        t_step += time.perf_counter() - h0
        if (step + 1) % args.target_every == 0:
            target.load_state_dict(net.state_dict())
        if step % 10 == 0 or step == args.steps - 1:
            print(f"step {step:4d}  td loss {float(loss.detach()):.5f}  agv_q {tuple(out['agv_q_values'].shape)}  "
                  f"picker_q {tuple(out['picker_q_values'].shape)}")
    per = t_step / args.steps
    print(f"{args.steps} updates of {args.batch_size} joint observations ({n_agents} agents each): {per * 1e3:.2f} ms per update "
          f"= {args.batch_size * n_agents / per:,.0f} agent decisions/s; observation -> HeteroData on the host "
          f"{t_build / args.steps * 1e3:.1f} ms per update (two batches)")
    S.ops  # noqa: B018  (the native library is loaded: HeteroGraphODENetwork has no eager path)


if __name__ == "__main__":
    main()
