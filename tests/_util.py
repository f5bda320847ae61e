"""Shared helpers for the parity tests (oracle on CPU vs the CUDA path)."""
import torch

from oracle.pyg_ref import RefBatch


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().cpu().double()
    b = b.detach().cpu().double()
    denom = float(b.norm())
    if denom == 0.0:
        return float((a - b).norm())
    return float((a - b).norm() / denom)


def to_ref_batch(batch) -> RefBatch:
    rb = RefBatch(x=batch.x.detach().cpu().clone(), edge_index=batch.edge_index.detach().cpu().clone())
    rb.batch = batch.batch.detach().cpu().clone()
    if getattr(batch, "is_current_agent", None) is not None:
        rb.is_current_agent = batch.is_current_agent.detach().cpu().clone()
    if getattr(batch, "ptr", None) is not None:
        rb.ptr = batch.ptr.detach().cpu().clone()
    return rb


def random_graph(num_nodes: int, num_edges: int, seed: int = 0, allow_isolated: bool = True) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, num_nodes, (num_edges,), generator=g)
    dst = torch.randint(0, num_nodes, (num_edges,), generator=g)
    return torch.stack([src, dst], dim=0)


FIXED_TOL = 1e-4   # north_star: fp32 relative L2 <= 1e-4 at fixed step
