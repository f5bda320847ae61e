"""Host-side data-parallel logic on CPU: two ``gloo`` ranks (no GPU needed).

What shards: independent graphs (contiguous ranges).  What is exchanged: one flat gradient all-reduce per
training step, weighted so that the masked-MEAN loss of scripts/train_gde.py:490 reproduces the unsharded
gradient, and (dopri5) two doubles per norm.  The CUDA kernels are not involved here: a CPU ``nn.Linear``
stands in for the model so that the collectives' arithmetic can be checked exactly.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from swarm_ode_b200 import dist as D
from swarm_ode_b200.data import Batch, Data


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _toy_graphs(num_graphs=7, seed=0):
    g = torch.Generator().manual_seed(seed)
    graphs = []
    for i in range(num_graphs):
        n = 3 + (i % 4)
        x = torch.randn(n, 5, generator=g)
        ei = torch.stack([torch.arange(n - 1), torch.arange(1, n)])
        mask = torch.zeros(n, dtype=torch.bool)
        mask[: 1 + (i % 3)] = True          # ragged number of masked ("current agent") nodes per graph
        graphs.append(Data(x=x, edge_index=ei, is_current_agent=mask))
    return graphs


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        graphs = _toy_graphs()
        full = Batch.from_data_list(graphs)
        lo, hi = D.shard_range(len(graphs), rank, world)
        mine = full.shard(rank, world)
        assert int(mine.batch.max()) + 1 == hi - lo
        # every rank sees a disjoint, contiguous node range; the union is the whole batch
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([mine.x.shape[0]]))
        assert int(sum(sizes)) == full.x.shape[0]

        torch.manual_seed(1)
        model = torch.nn.Linear(5, 2)
        target_full = torch.randn(int(full.is_current_agent.sum()), 2, generator=torch.Generator().manual_seed(2))
        # unsharded reference gradient of the masked-mean loss
        ref = torch.nn.Linear(5, 2)
        ref.load_state_dict(model.state_dict())
        loss_ref = torch.nn.functional.mse_loss(ref(full.x)[full.is_current_agent], target_full)
        loss_ref.backward()
        # this rank's shard: its slice of the targets follows the masked-node order
        before = int(full.is_current_agent[: int((full.batch < lo).sum())].sum())
        local_n = int(mine.is_current_agent.sum())
        tgt = target_full[before: before + local_n]
        loss = torch.nn.functional.mse_loss(model(mine.x)[mine.is_current_agent], tgt)
        loss.backward()
        total_n = D.global_count(local_n)
        assert total_n == target_full.shape[0]
        D.allreduce_gradients(model.parameters(), local_n / total_n)
        for p, q in zip(model.parameters(), ref.parameters()):
            assert torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-6), (p.grad, q.grad)
        # the whole training step (one sync-free collective): same gradients and the global masked-mean loss

        class _Toy(torch.nn.Module):
            def __init__(self, lin):
                super().__init__()
                self.lin = lin

            def forward(self, graphs, time_span):
                y = self.lin(graphs.x)
                return {"trajectories": torch.stack([torch.zeros_like(y), y])}

        toy = _Toy(torch.nn.Linear(5, 2))
        toy.lin.load_state_dict(ref.state_dict())
        opt = torch.optim.SGD(toy.parameters(), lr=0.0)
        gl = D.masked_mse_train_step(toy, opt, mine, tgt, torch.tensor([0.0, 1.0]), max_norm=1e9)
        assert torch.allclose(gl, loss_ref.detach(), rtol=1e-5, atol=1e-7), (gl, loss_ref)
        for p, q in zip(toy.lin.parameters(), ref.parameters()):
            assert torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-6), (p.grad, q.grad)
        # dopri5 norm hook: (sum of squares, element count) summed over ranks
        hook = D.dopri5_norm_allreduce()
        s, c = hook(float(rank + 1), 10.0 * (rank + 1))
        assert s == sum(r + 1 for r in range(world)) and c == 10.0 * sum(r + 1 for r in range(world))
        # the device-side exchange is only offered for CUDA devices (gnode_set_dopri5_device_allreduce); its arithmetic is
        # the same in-place SUM, checked here on the host tensor the hook would be handed on a GPU
        assert not hasattr(hook, "device_allreduce")
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        assert float(t) == s
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gradient_allreduce_and_norm_hook(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_zeros_like_many_returns_independent_aligned_views():
    """ops._zeros_like_many: one flat zero buffer behind the parameter-gradient tensors of a backward call."""
    from swarm_ode_b200 import ops
    like = [torch.ones(3, 5), torch.ones(7), torch.ones(64, 2), torch.ones(1)]
    out = ops._zeros_like_many(like)
    assert [o.shape for o in out] == [t.shape for t in like]
    assert all(o.is_contiguous() and float(o.abs().sum()) == 0.0 for o in out)
    base = out[0].data_ptr()
    assert all((o.data_ptr() - base) % 256 == 0 for o in out)          # every view starts on a 256-byte boundary
    out[1].fill_(3.0)                                                   # views do not overlap
    assert float(out[0].abs().sum()) == 0.0 and float(out[2].abs().sum()) == 0.0 and float(out[1].sum()) == 21.0
    assert ops._zeros_like_many([]) == []


def test_single_process_helpers_are_identity():
    assert D.world() == (0, 1) and not D.is_dist()
    assert D.global_count(5) == 5 and D.dopri5_norm_allreduce() is None
    assert D.shard_range(10, 1, 4) == (2, 5)
