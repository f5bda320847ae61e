"""The C-ABI library loads on a GPU-less host and exports every symbol include/gnode_b200.h declares.
No compute entry point is called here (argument validation that returns before any CUDA call is)."""
import ctypes as C
import os
import re

import pytest

import swarm_ode_b200 as S
from swarm_ode_b200 import _lib

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "gnode_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(gnode_[a-z0-9_]+)\s*\(", src))
    names.discard("gnode_allreduce_fn")
    return sorted(names)


def test_library_is_built_in_tree():
    assert os.path.exists(S.LIB_PATH), "run __graft_entry__.build()"
    assert os.path.dirname(S.LIB_PATH).endswith("swarm_ode_b200")


def test_every_declared_symbol_is_exported_and_bound():
    handle = C.CDLL(S.LIB_PATH)
    declared = _declared()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in gnode_b200.h but not exported"
    assert set(declared) == set(_lib.EXPORTED_SYMBOLS), "python binding table out of sync with the header"


def test_version_and_engine_switch():
    L = _lib.lib()
    assert L.gnode_abi_version() == 1
    prev = S.set_engine("simt")
    assert S.set_engine(prev) == "simt"
    assert L.gnode_set_engine(99) == -1


def test_argument_validation_returns_error_codes_without_touching_cuda():
    L = _lib.lib()
    rc = L.gnode_csr_build(None, 0, 0, None, None, None, None, None, 0, None)
    assert rc == -1 and b"n_nodes" in L.gnode_last_error()
    rc = L.gnode_decoder_fwd(None, 10, 16, 99, None, None, None, None)
    assert rc == -1 and b"n_out" in L.gnode_last_error()
    rc = L.gnode_integrate_fixed(None, None, 2, None, None, 0, None, None, 0, None, 0, None)
    assert rc == -1 and b"graph is null" in L.gnode_last_error()
    with pytest.raises(S.GnodeError, match="status -1"):
        _lib.check(rc, "gnode_integrate_fixed")


def test_workspace_queries_scale_with_problem_size():
    L = _lib.lib()
    small = L.gnode_integrate_fixed_workspace_bytes(1000, 399, 64, _lib.GNODE_RK4_38, 0)
    big = L.gnode_integrate_fixed_workspace_bytes(100000, 399, 64, _lib.GNODE_RK4_38, 0)
    bwd = L.gnode_integrate_fixed_workspace_bytes(100000, 399, 64, _lib.GNODE_RK4_38, 1)
    assert 0 < small < big < bwd
    assert L.gnode_integrate_fixed_workspace_bytes(1000, 399, 64, _lib.GNODE_DOPRI5, 0) == 0
    assert L.gnode_integrate_dopri5_workspace_bytes(1000, 435, 64) > 10 * 1000 * 435 * 4


def test_cpu_tensors_raise_instead_of_falling_back():
    import torch
    f = S.GraphODEFunc(8, 4)
    with pytest.raises(S.GnodeError, match="CUDA"):
        f(torch.tensor(0.0), torch.zeros(3, 8), torch.zeros(2, 0, dtype=torch.long))
    with pytest.raises(S.GnodeError):
        S.odeint(lambda t, y: y, torch.zeros(3), torch.tensor([0.0, 1.0]), method="euler")
    with pytest.raises(ValueError):
        S.odeint(f.bind(torch.zeros(2, 0, dtype=torch.long)), torch.zeros(3, 8), torch.tensor([0.0, 1.0]), method="rk45")
