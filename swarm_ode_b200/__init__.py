"""swarm_ode_b200 -- B200-native (sm_100a) implementation of the swarm-ode GNODE hot path.

Drop-in modules with the reference's signatures (scripts/train_gde.py, scripts/gnode.py) on top of a
thin C-ABI CUDA library (``libgnode_b200.so``, declared in ``include/gnode_b200.h``).  No Triton, no
multi-backend dispatch, no CPU fallback: importing the package is cheap and GPU-free, but every
compute entry point raises unless the extension is built and the tensors live on a CUDA device.
"""
from ._lib import GnodeError, set_engine, set_fold, launch_count, LIB_PATH  # noqa: F401
from .data import (Batch, Data, GraphConverter, PackedBatch, TrajectoryBatch, build_episode_batch, collate_trajectory_batches,  # noqa: F401
                   extract_positions_from_graph, spatial_edges_cuda)
from .graph import CSRGraph, csr_for  # noqa: F401
from .modules import GraphODE, GraphODEFunc, ODEFunction, SAGEConv, BoundGraphODEFunc  # noqa: F401
from .odeint import odeint, odeint_adjoint  # noqa: F401
from .hetero import HeteroData, HeteroConv, HeteroGraphODENetwork  # noqa: F401
from . import ops, synthetic  # noqa: F401
from .graphed import GraphedTrainStep  # noqa: F401

__all__ = [
    "GnodeError", "set_engine", "set_fold", "launch_count", "LIB_PATH",
    "Batch", "Data", "GraphConverter", "PackedBatch", "TrajectoryBatch", "build_episode_batch", "collate_trajectory_batches",
    "extract_positions_from_graph", "spatial_edges_cuda",
    "CSRGraph", "csr_for", "GraphedTrainStep", "odeint_adjoint",
    "GraphODE", "GraphODEFunc", "ODEFunction", "SAGEConv", "BoundGraphODEFunc",
    "HeteroData", "HeteroConv", "HeteroGraphODENetwork",
    "odeint", "ops", "synthetic",
]
