"""ctypes binding of ``libgnode_b200.so`` (the C ABI declared in ``include/gnode_b200.h``).

There is no CPU fallback: if the shared library is missing or an entry point is absent the import
of any compute function raises.  The library is built in-tree by ``__graft_entry__.build()``
(``make -C swarm_ode_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgnode_b200.so")

# ---- constants mirrored from include/gnode_b200.h ----
GNODE_EULER, GNODE_MIDPOINT, GNODE_RK4_38, GNODE_DOPRI5 = 0, 1, 2, 3
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TC = 0, 1, 2
PACK_U8, PACK_I16, PACK_F16, PACK_F32, PACK_BITS = 0, 1, 2, 3, 4
METHODS = {"euler": GNODE_EULER, "midpoint": GNODE_MIDPOINT, "rk4": GNODE_RK4_38, "dopri5": GNODE_DOPRI5}


class GnodeGraph(C.Structure):
    _fields_ = [("n_nodes", C.c_int64), ("n_edges", C.c_int64), ("rowptr", C.c_void_p), ("col", C.c_void_p),
                ("t_rowptr", C.c_void_p), ("t_col", C.c_void_p), ("tiles", C.c_void_p), ("tile_err", C.c_void_p),
                ("tile_rows", C.c_int32)]


class GnodeSage3Params(C.Structure):
    _fields_ = [("node_dim", C.c_int32), ("hidden_dim", C.c_int32)] + [
        (n, C.c_void_p) for n in ("w1l", "b1", "w1r", "w2l", "b2", "w2r", "w3l", "b3", "w3r")]


class GnodeSage3Grads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w1l", "b1", "w1r", "w2l", "b2", "w2r", "w3l", "b3", "w3r")]


class GnodeDopri5Stats(C.Structure):
    _fields_ = [("nfe", C.c_int64), ("n_accepted", C.c_int64), ("n_attempted", C.c_int64),
                ("first_step", C.c_double), ("last_dt", C.c_double), ("min_margin", C.c_double)]


class GnodeDopri5Trace(C.Structure):
    _fields_ = [("error_ratio", C.POINTER(C.c_double)), ("dt", C.POINTER(C.c_double)),
                ("accepted", C.POINTER(C.c_int32)), ("trace_cap", C.c_int64)]


class GnodeMlpParams(C.Structure):
    _fields_ = [("dim", C.c_int32), ("hidden_dim", C.c_int32)] + [
        (n, C.c_void_p) for n in ("w0", "b0", "w1", "b1", "w2", "b2")]


class GnodeMlpGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w0", "b0", "w1", "b1", "w2", "b2")]


class GnodeProfEntry(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("launches", C.c_int64), ("ms", C.c_double), ("flops", C.c_double),
                ("bytes", C.c_double)]


ALLREDUCE_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_void_p)
ALLREDUCE_DEV_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p)

_P = C.c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)
    "gnode_last_error": (C.c_char_p, []),
    "gnode_abi_version": (C.c_int, []),
    "gnode_set_engine": (C.c_int, [C.c_int]),
    "gnode_set_fold": (C.c_int, [C.c_int]),
    "gnode_set_dopri5_fsal": (C.c_int, [C.c_int]),
    "gnode_set_dopri5_device_allreduce": (C.c_int, [ALLREDUCE_DEV_FN, C.c_void_p]),
    "gnode_launch_count": (C.c_int64, []),
    "gnode_tc_status": (C.c_int, [_P]),
    "gnode_tc_status_async": (C.c_int, [_P, _P]),
    "gnode_prof_enable": (C.c_int, [C.c_int]),
    "gnode_prof_read": (C.c_int, [C.POINTER(GnodeProfEntry), C.c_int]),
    "gnode_csr_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "gnode_csr_build": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "gnode_csr_build_async": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "gnode_tiles_build": (C.c_int, [_P, C.c_int64, _P, _P]),
    "gnode_tiles_build_rows": (C.c_int, [_P, C.c_int64, C.c_int32, _P, _P]),
    "gnode_gemm_nt_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "gnode_gemm_nt": (C.c_int, [_P, C.c_int64, _P, C.c_int64, _P, C.c_int64, C.c_int64, C.c_int32, C.c_int32, _P,
                                C.c_int32, _P, C.c_int64, C.c_float, _P, C.c_size_t, _P]),
    "gnode_gemm_tn_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int64]),
    "gnode_gemm_tn": (C.c_int, [_P, C.c_int64, _P, C.c_int64, _P, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                C.c_float, _P, C.c_size_t, _P]),
    "gnode_sage_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32]),
    "gnode_sage_fwd": (C.c_int, [C.POINTER(GnodeGraph), _P, C.c_int32, C.c_int32, _P, _P, _P, C.c_int32, _P, _P,
                                 C.c_size_t, _P]),
    "gnode_sage_bwd": (C.c_int, [C.POINTER(GnodeGraph), _P, _P, _P, C.c_int32, C.c_int32, _P, _P, C.c_int32, _P, _P,
                                 _P, _P, _P, C.c_size_t, _P]),
    "gnode_sage_bipartite_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32]),
    "gnode_sage_bipartite_fwd": (C.c_int, [C.POINTER(GnodeGraph), C.c_int64, _P, _P, C.c_int32, C.c_int32, _P, _P, _P,
                                           C.c_float, _P, C.c_int32, _P, _P, C.c_size_t, _P]),
    "gnode_rhs_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32]),
    "gnode_rhs_fwd": (C.c_int, [C.POINTER(GnodeGraph), C.POINTER(GnodeSage3Params), _P, _P, _P, C.c_size_t, _P]),
    "gnode_rhs_bwd": (C.c_int, [C.POINTER(GnodeGraph), C.POINTER(GnodeSage3Params), _P, _P, _P,
                                C.POINTER(GnodeSage3Grads), _P, C.c_size_t, _P]),
    "gnode_integrate_fixed_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "gnode_integrate_fixed_save_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "gnode_integrate_fixed": (C.c_int, [C.POINTER(GnodeGraph), C.POINTER(GnodeSage3Params), C.c_int32, _P,
                                        C.POINTER(C.c_float), C.c_int32, _P, _P, C.c_size_t, _P, C.c_size_t, _P]),
    "gnode_integrate_fixed_flags": (C.c_int, [C.POINTER(GnodeGraph), C.POINTER(GnodeSage3Params), C.c_int32, _P,
                                              C.POINTER(C.c_float), C.c_int32, _P, _P, C.c_size_t, _P, C.c_size_t, C.c_int32, _P]),
    "gnode_integrate_fixed_decoded_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "gnode_integrate_fixed_decoded": (C.c_int, [C.POINTER(GnodeGraph), C.POINTER(GnodeSage3Params), C.c_int32, _P,
                                                C.POINTER(C.c_float), C.c_int32, _P, _P, C.c_size_t, _P, _P, C.c_int32, _P,
                                                _P, C.c_size_t, _P]),
    "gnode_integrate_fixed_bwd": (C.c_int, [C.POINTER(GnodeGraph), C.POINTER(GnodeSage3Params), C.c_int32, _P,
                                            C.POINTER(C.c_float), C.c_int32, _P, _P, C.POINTER(GnodeSage3Grads), _P,
                                            C.c_size_t, _P, C.c_size_t, _P]),
    "gnode_integrate_fixed_adjoint_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32, C.c_int32]),
    "gnode_integrate_fixed_adjoint": (C.c_int, [C.POINTER(GnodeGraph), C.POINTER(GnodeSage3Params), C.c_int32, _P,
                                                C.POINTER(C.c_float), C.c_int32, _P, _P, C.POINTER(GnodeSage3Grads), _P,
                                                C.c_size_t, _P]),
    "gnode_integrate_fixed_bwd_decoded_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "gnode_integrate_fixed_bwd_decoded": (C.c_int, [C.POINTER(GnodeGraph), C.POINTER(GnodeSage3Params), C.c_int32, _P,
                                                    C.POINTER(C.c_float), C.c_int32, _P, _P, C.c_int32,
                                                    C.POINTER(GnodeSage3Grads), _P, C.c_size_t, _P, C.c_size_t, _P]),
    "gnode_integrate_dopri5_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32]),
    "gnode_integrate_dopri5": (C.c_int, [C.POINTER(GnodeGraph), C.POINTER(GnodeSage3Params), _P,
                                         C.POINTER(C.c_double), C.c_int32, C.c_double, C.c_double, _P,
                                         C.POINTER(GnodeDopri5Stats), C.POINTER(GnodeDopri5Trace), ALLREDUCE_FN, _P,
                                         C.c_int64, _P, C.c_size_t, _P]),
    "gnode_integrate_dopri5_bwd_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32, C.c_int32]),
    "gnode_integrate_dopri5_bwd": (C.c_int, [C.POINTER(GnodeGraph), C.POINTER(GnodeSage3Params), _P,
                                             C.POINTER(C.c_double), C.c_int32, C.POINTER(C.c_double), C.c_int32, _P, _P,
                                             C.POINTER(GnodeSage3Grads), _P, C.c_size_t, _P]),
    "gnode_decoder_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32]),
    "gnode_decoder_fwd": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "gnode_decoder_fwd_copy": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P, _P, _P]),
    "gnode_decoder_bwd": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "gnode_mlp_ode_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32, C.c_int32]),
    "gnode_mlp_rhs_fwd": (C.c_int, [C.POINTER(GnodeMlpParams), _P, C.c_int64, _P, _P, C.c_size_t, _P]),
    "gnode_mlp_integrate_fixed": (C.c_int, [C.POINTER(GnodeMlpParams), C.c_int32, _P, C.c_int64,
                                            C.POINTER(C.c_float), C.c_int32, _P, _P, C.c_size_t, _P]),
    "gnode_mlp_integrate_dopri5": (C.c_int, [C.POINTER(GnodeMlpParams), _P, C.c_int64, C.POINTER(C.c_double),
                                             C.c_int32, C.c_double, C.c_double, _P, C.POINTER(GnodeDopri5Stats),
                                             C.POINTER(GnodeDopri5Trace), C.c_int64, _P, C.c_size_t, _P]),
    "gnode_gemm_k128_workspace_bytes": (C.c_size_t, [C.c_int32]),
    "gnode_gemm_k128": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int64, C.c_int32, _P, C.c_float, _P, C.c_int64, C.c_float, _P,
                                  C.c_int64, C.c_float, _P, C.c_size_t, _P]),
    "gnode_window_graphs_nodes": (C.c_int64, [C.c_int64, C.c_int32, C.c_int32]),
    "gnode_window_graphs_edge_capacity": (C.c_int64, [C.c_int64, C.c_int32, C.c_int32]),
    "gnode_window_graphs_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "gnode_window_graphs": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_int32, _P, _P, C.c_int64,
                                      _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "gnode_sage_bipartite_bwd_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32]),
    "gnode_sage_bipartite_bwd": (C.c_int, [C.POINTER(GnodeGraph), C.c_int64, C.c_int64, _P, _P, C.c_int32, C.c_int32, _P, _P,
                                           _P, _P, C.c_float, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "gnode_linear_bwd_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32]),
    "gnode_linear_bwd": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P, _P, C.c_size_t, _P]),
    "gnode_mlp_bwd_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "gnode_mlp_rhs_bwd": (C.c_int, [C.POINTER(GnodeMlpParams), _P, _P, C.c_int64, _P, C.POINTER(GnodeMlpGrads), _P,
                                    C.c_size_t, _P]),
    "gnode_mlp_integrate_fixed_bwd": (C.c_int, [C.POINTER(GnodeMlpParams), C.c_int32, _P, C.c_int64, C.POINTER(C.c_float),
                                                C.c_int32, _P, _P, C.POINTER(GnodeMlpGrads), _P, C.c_size_t, _P]),
    "gnode_mlp_integrate_dopri5_bwd": (C.c_int, [C.POINTER(GnodeMlpParams), _P, C.c_int64, C.POINTER(C.c_double), C.c_int32,
                                                 C.POINTER(C.c_double), C.c_int32, _P, _P, C.POINTER(GnodeMlpGrads), _P,
                                                 C.c_size_t, _P]),
    "gnode_spatial_edges": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_float, _P, _P, _P]),
    "gnode_unpack_features": (C.c_int, [_P, C.c_int32, C.c_int64, _P, _P]),
    "gnode_unpack_bits": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P]),
    "gnode_unpack_edges": (C.c_int, [_P, C.c_int64, _P, _P]),
    "gnode_batch_vector": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib: Optional[C.CDLL] = None


class GnodeError(RuntimeError):
    """Raised when a libgnode_b200 entry point returns a negative status."""


def lib() -> C.CDLL:
    """Load the shared library (once).  Raises loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: the sm_100a CUDA extension has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C swarm_ode_b200/csrc`). "
                "There is no CPU fallback for the GNODE path.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = handle
        env_engine = os.environ.get("GNODE_ENGINE")  # 'auto' | 'simt' | 'tc': initial GEMM engine
        if env_engine:
            handle.gnode_set_engine({"auto": ENGINE_AUTO, "simt": ENGINE_SIMT, "tc": ENGINE_TC}[env_engine])
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().gnode_last_error().decode("utf-8", "replace")
        raise GnodeError(f"{what or 'libgnode_b200'} failed (status {rc}): {msg}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise GnodeError(f"{name} must live on a CUDA device (got {t.device}); libgnode_b200 has no CPU path")
    if t.dtype != torch.float32:
        raise GnodeError(f"{name} must be float32 (got {t.dtype})")
    return t.contiguous()


class Workspace:
    """Grow-only per-device scratch buffer handed to the library (it never allocates itself)."""

    def __init__(self):
        self._bufs = {}

    def get(self, nbytes: int, device, tag: str = "main") -> torch.Tensor:
        # one buffer per (device, stream): work on different streams may overlap in time
        key = (str(device), tag, torch.cuda.current_stream(device).cuda_stream)
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            self._bufs.pop(key, None)
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._bufs[key] = buf
        return buf

    def clear(self):
        self._bufs.clear()


WORKSPACE = Workspace()


def set_engine(name: str) -> str:
    """Select the GEMM engine: 'auto' | 'simt' (fp32 FFMA parity anchor) | 'tc' (tcgen05 3xTF32)."""
    code = {"auto": ENGINE_AUTO, "simt": ENGINE_SIMT, "tc": ENGINE_TC}[name]
    prev = lib().gnode_set_engine(code)
    return {ENGINE_AUTO: "auto", ENGINE_SIMT: "simt", ENGINE_TC: "tc"}.get(prev, "auto")


def set_fold(on: bool) -> bool:
    """Folded (True, default) or direct (False) evaluation of the RK stages; returns the previous setting."""
    return bool(lib().gnode_set_fold(1 if on else 0))


def set_dopri5_fsal(on: bool) -> bool:
    """Step-level re-associations of the folded dopri5 (Z_0 handed from stage 6 of an accepted step to the next attempt,
    dense output as one projection) on (default) or off; returns the previous setting."""
    return bool(lib().gnode_set_dopri5_fsal(1 if on else 0))


def launch_count() -> int:
    return int(lib().gnode_launch_count())


def prof_enable(on: bool) -> None:
    """Start (and reset) / stop per-kernel-class CUDA-event timing inside the library."""
    lib().gnode_prof_enable(1 if on else 0)


def prof_read():
    """List of dicts {name, launches, ms, flops, bytes} accumulated since ``prof_enable(True)``."""
    cap = 128
    arr = (GnodeProfEntry * cap)()
    n = min(lib().gnode_prof_read(arr, cap), cap)
    return [dict(name=arr[i].name.decode(), launches=int(arr[i].launches), ms=float(arr[i].ms),
                 flops=float(arr[i].flops), bytes=float(arr[i].bytes)) for i in range(n)]


def tc_check(device=None) -> None:
    """Synchronise and raise if a tcgen05 kernel reported a barrier timeout (tests / debugging)."""
    dev = device if device is not None else torch.cuda.current_device()
    with torch.cuda.device(dev):
        check(lib().gnode_tc_status(stream_ptr(dev)), "gnode_tc_status")
