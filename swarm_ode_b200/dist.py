"""Data parallelism over independent graphs (one process per GPU, ``torch.distributed``).

The reference is single-process; the only parallel axis the path offers is the batch:
``Batch.from_data_list`` is a disjoint union (scripts/train_gde.py:367), no edge crosses graphs, so
contiguous graph ranges shard across ranks with **no data-path collective** for fixed-step
integration.  Two exchanges exist:

* training: ONE all-reduce (sum) of a flat fp32 gradient buffer per step; the loss is a mean over
  masked nodes (scripts/train_gde.py:490), so each rank's gradient is weighted by
  ``local_masked / global_masked`` to reproduce the single-process numerics; gradient clipping then
  uses the global norm (scripts/train_gde.py:494);
* dopri5: the reference's error norm and step size are global over the whole batch tensor, so each
  attempted step all-reduces (sum of squares, element count) -- two doubles -- to make every rank
  take the decisions of the unsharded batch.
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def is_dist() -> bool:
    return dist.is_available() and dist.is_initialized()


def world() -> Tuple[int, int]:
    return (dist.get_rank(), dist.get_world_size()) if is_dist() else (0, 1)


def shard_range(num_graphs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced range of graphs of one rank (same rule as ``Batch.shard``)."""
    return (num_graphs * rank) // world_size, (num_graphs * (rank + 1)) // world_size


def allreduce_gradients(params: Iterable[torch.nn.Parameter], local_weight: float = 1.0,
                        group=None) -> Optional[torch.Tensor]:
    """Sum ``local_weight * grad`` over ranks through one flat buffer and write the result back.

    ``local_weight`` = ``local_masked / global_masked`` for the masked-mean loss.  Returns the flat
    reduced buffer (for norm computation) or None when there is nothing to reduce.
    """
    params = [p for p in params if p.grad is not None]
    if not params:
        return None
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    if local_weight != 1.0:
        flat.mul_(local_weight)
    if is_dist():
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return flat


def global_count(local: int, device=None, group=None) -> int:
    """Sum of an integer over ranks (e.g. number of masked nodes)."""
    if not is_dist():
        return int(local)
    t = torch.tensor([float(local)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(round(t.item()))


def dopri5_norm_allreduce(device=None, group=None) -> Optional[Callable[[float, float], Tuple[float, float]]]:
    """Hook for ``GraphODE.dopri5_allreduce`` / ``ops.integrate_dopri5(allreduce=...)``."""
    if not is_dist():
        return None

    def hook(sumsq: float, count: float) -> Tuple[float, float]:
        t = torch.tensor([sumsq, count], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        s, c = t.tolist()
        return s, c

    def device_allreduce(t: torch.Tensor) -> None:
        """In-place SUM over the ranks of the local sum of squares where the library left it (device memory), enqueued
        on the current stream: the per-attempt exchange of the GNODE solve; ``hook`` itself then runs once per solve
        (for the global element count)."""
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)

    if device is not None and torch.device(device).type == "cuda":
        hook.device_allreduce = device_allreduce
    return hook


def masked_mse_train_step(model, optimizer, graphs, next_positions: torch.Tensor,
                          time_span: Optional[torch.Tensor] = None, max_norm: float = 1.0,
                          group=None, distributed: Optional[bool] = None) -> torch.Tensor:
    """One training step of scripts/train_gde.py:478-495 on this rank's shard.

    forward -> masked MSE -> backward -> (all-reduce) -> clip_grad_norm_(1.0) -> optimizer.step().
    Returns the *global* loss (detached).  With one rank this is exactly the reference step
    (with the reference's device-mismatch bug at :476/:490 corrected).  ``distributed=False`` keeps the step local to this
    process even when a process group exists (a rank-0-only side computation must not enter a collective).
    """
    use_dist = is_dist() if distributed is None else (bool(distributed) and is_dist())
    if time_span is None:
        time_span = torch.tensor([0.0, 1.0], device=graphs.x.device)
    optimizer.zero_grad(set_to_none=True)
    pred = model(graphs, time_span)["trajectories"][1]
    mask = graphs.is_current_agent
    target = next_positions.view(-1, 2)
    local_n = int(target.shape[0])
    # pred[mask] of the reference (scripts/train_gde.py:490) synchronises with the host to size its result; the number
    # of masked nodes is known here (one target row each), so the same rows are gathered without a synchronisation
    if hasattr(torch, "nonzero_static") and mask.is_cuda:
        idx = torch.nonzero_static(mask, size=local_n).view(-1)
        loss = torch.nn.functional.mse_loss(pred.index_select(0, idx), target)
    else:
        loss = torch.nn.functional.mse_loss(pred[mask], target)
    loss.backward()
    if use_dist:
        # ONE collective per step and no host synchronisation: every rank contributes local_n * [grads, loss] and local_n
        # itself; dividing the sum by the global count reproduces the masked mean of the unsharded batch
        params = [p for p in model.parameters() if p.grad is not None]
        tail = torch.stack([loss.detach().reshape(()), torch.ones((), device=pred.device, dtype=loss.dtype)])
        flat = torch.cat([p.grad.reshape(-1) for p in params] + [tail]).mul_(float(local_n))
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(flat[-1].clamp_min(1.0))
        views, off = [], 0
        for p in params:
            n = p.grad.numel()
            views.append(flat[off:off + n].view_as(p.grad))
            off += n
        torch._foreach_copy_([p.grad for p in params], views)      # one launch, not one per parameter
        gl = flat[-2].clone()
    else:
        gl = loss.detach()
    torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=max_norm)
    optimizer.step()
    return gl
