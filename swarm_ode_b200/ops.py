"""torch.autograd bridges onto the C ABI (one Function per fused op).

PyTorch is plumbing here: it owns device memory, streams and the parameter tensors; every FLOP of
the GNODE path runs inside ``libgnode_b200.so``.  No function in this module has a CPU or eager
fallback -- CPU tensors raise ``GnodeError``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import METHODS, GnodeError
from .graph import CSRGraph

_f32 = _lib.require_cuda_f32

# A differentiable fixed-grid solve keeps its per-stage intermediates for the backward pass when they take at
# most this fraction of the currently free device memory; otherwise the backward recomputes them.
SAVE_FRACTION = 0.5


_TOTAL_MEM: dict = {}


def _save_fits(nbytes: int, device) -> bool:
    """Keep the forward's intermediates when they take at most SAVE_FRACTION of the free device memory.  Small save
    areas (< 1/16 of the device) skip the driver query: cudaMemGetInfo is a synchronous driver call per step."""
    if SAVE_FRACTION <= 0.0:
        return False
    key = str(device)
    if key not in _TOTAL_MEM:
        _TOTAL_MEM[key] = torch.cuda.get_device_properties(device).total_memory
    if SAVE_FRACTION >= 0.25 and nbytes <= _TOTAL_MEM[key] // 16:
        return True
    free, _total = torch.cuda.mem_get_info(device)
    return nbytes <= SAVE_FRACTION * free


def _zeros_like_many(tensors: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """Zero-filled gradient buffers for ``tensors`` as views of ONE flat allocation (one fill kernel instead of one per
    parameter; every view starts on a 256-byte boundary, like a tensor of its own)."""
    if not tensors:
        return []
    offs, total = [], 0
    for t in tensors:
        offs.append(total)
        total += (t.numel() + 63) // 64 * 64
    flat = torch.zeros(total, dtype=tensors[0].dtype, device=tensors[0].device)
    return [flat[o:o + t.numel()].view(t.shape) for o, t in zip(offs, tensors)]


def _ws(nbytes: int, device) -> torch.Tensor:
    return _lib.WORKSPACE.get(nbytes, device)


def _sage3_params(D: int, H: int, w: Sequence[torch.Tensor]) -> _lib.GnodeSage3Params:
    shapes = [(H, D), (H,), (H, D), (H, H), (H,), (H, H), (D, H), (D,), (D, H)]
    for t, shp in zip(w, shapes):
        if tuple(t.shape) != shp:
            raise GnodeError(f"GraphODEFunc parameter has shape {tuple(t.shape)}, expected {shp}")
    return _lib.GnodeSage3Params(D, H, *[t.data_ptr() for t in w])


def _float_array(vals) -> "C.Array":
    return (C.c_float * len(vals))(*[float(v) for v in vals])


# ----------------------------------------------------------------------------------------------
# single SAGEConv layer
# ----------------------------------------------------------------------------------------------
class _SageConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, wl, bl, wr, graph: CSRGraph, relu: bool):
        x, wl, bl, wr = _f32(x, "x"), _f32(wl, "lin_l.weight"), _f32(bl, "lin_l.bias"), _f32(wr, "lin_r.weight")
        N, ci = x.shape
        co = wl.shape[0]
        if N != graph.num_nodes:
            raise GnodeError(f"x has {N} rows but the graph has {graph.num_nodes} nodes")
        out = torch.empty((N, co), dtype=torch.float32, device=x.device)
        L = _lib.lib()
        ws = _ws(L.gnode_sage_workspace_bytes(N, ci, co), x.device)
        with torch.cuda.device(x.device):
            _lib.check(L.gnode_sage_fwd(graph.ref(), _lib.ptr(x), ci, co, _lib.ptr(wl), _lib.ptr(bl), _lib.ptr(wr),
                                        int(relu), _lib.ptr(out), _lib.ptr(ws), ws.numel(),
                                        _lib.stream_ptr(x.device)), "gnode_sage_fwd")
        ctx.graph, ctx.relu = graph, relu
        ctx.save_for_backward(x, wl, wr, out)
        return out

    @staticmethod
    def backward(ctx, g):
        x, wl, wr, out = ctx.saved_tensors
        g = _f32(g, "grad_out")
        N, ci = x.shape
        co = wl.shape[0]
        need_x, need_wl, need_bl, need_wr = ctx.needs_input_grad[:4]
        gx = torch.empty_like(x) if need_x else None
        gwl = torch.zeros_like(wl) if need_wl else None
        gbl = torch.zeros(co, dtype=torch.float32, device=x.device) if need_bl else None
        gwr = torch.zeros_like(wr) if need_wr else None
        L = _lib.lib()
        ws = _ws(L.gnode_sage_workspace_bytes(N, ci, co), x.device)
        with torch.cuda.device(x.device):
            _lib.check(L.gnode_sage_bwd(ctx.graph.ref(), _lib.ptr(x), _lib.ptr(out), _lib.ptr(g), ci, co, _lib.ptr(wl),
                                        _lib.ptr(wr), int(ctx.relu), _lib.ptr(gx), _lib.ptr(gwl), _lib.ptr(gbl),
                                        _lib.ptr(gwr), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(x.device)),
                       "gnode_sage_bwd")
        return gx, gwl, gbl, gwr, None, None


def sage_conv(x, wl, bl, wr, graph: CSRGraph, relu: bool = False) -> torch.Tensor:
    return _SageConvFn.apply(x, wl, bl, wr, graph, relu)


# ----------------------------------------------------------------------------------------------
# GraphODEFunc (three layers) -- one RHS evaluation
# ----------------------------------------------------------------------------------------------
class _RhsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, graph: CSRGraph, *w):
        x = _f32(x, "x")
        w = [_f32(t, "param") for t in w]
        N, D = x.shape
        H = w[0].shape[0]
        if N != graph.num_nodes:
            raise GnodeError(f"x has {N} rows but the graph has {graph.num_nodes} nodes")
        p = _sage3_params(D, H, w)
        out = torch.empty_like(x)
        L = _lib.lib()
        ws = _ws(L.gnode_rhs_workspace_bytes(N, D, H), x.device)
        with torch.cuda.device(x.device):
            _lib.check(L.gnode_rhs_fwd(graph.ref(), C.byref(p), _lib.ptr(x), _lib.ptr(out), _lib.ptr(ws), ws.numel(),
                                       _lib.stream_ptr(x.device)), "gnode_rhs_fwd")
        ctx.graph = graph
        ctx.save_for_backward(x, *w)
        return out

    @staticmethod
    def backward(ctx, g):
        x, *w = ctx.saved_tensors
        g = _f32(g, "grad_out")
        N, D = x.shape
        H = w[0].shape[0]
        p = _sage3_params(D, H, w)
        gx = torch.empty_like(x)
        gw = _zeros_like_many(w)
        grads = _lib.GnodeSage3Grads(*[t.data_ptr() for t in gw])
        L = _lib.lib()
        ws = _ws(L.gnode_rhs_workspace_bytes(N, D, H), x.device)
        with torch.cuda.device(x.device):
            _lib.check(L.gnode_rhs_bwd(ctx.graph.ref(), C.byref(p), _lib.ptr(x), _lib.ptr(g), _lib.ptr(gx),
                                       C.byref(grads), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(x.device)),
                       "gnode_rhs_bwd")
        return (gx, None, *gw)


def gnode_rhs(x, graph: CSRGraph, params: Sequence[torch.Tensor]) -> torch.Tensor:
    """dx/dt of GraphODEFunc; ``params`` = (w1l, b1, w1r, w2l, b2, w2r, w3l, b3, w3r)."""
    return _RhsFn.apply(x, graph, *params)


# ----------------------------------------------------------------------------------------------
# fixed-grid integration with backprop through the solver
# ----------------------------------------------------------------------------------------------
def _fixed_forward(ctx, y0, graph: CSRGraph, method: int, t_host: Tuple[float, ...], w, decoder=None):
    """Shared forward of the fixed-grid autograd nodes: runs the solve, keeps the save area on ``ctx``.
    Returns ``(solution, contiguous parameter list)`` -- with ``decoder = (weight, bias)`` ``(solution, parameter list,
    trajectories)``: the solve and ``position_decoder`` over every time point in one library call
    (``gnode_integrate_fixed_decoded``: ``solution[0] = y0`` is written while ``y0`` streams through the decoder of the
    first time point, later time points are decoded from the 2H-wide step combination).  The caller does
    ``ctx.save_for_backward``."""
    if True:
        y0 = _f32(y0, "y0")
        w = [_f32(t, "param") for t in w]
        N, D = y0.shape
        H = w[0].shape[0]
        if N != graph.num_nodes:
            raise GnodeError(f"y0 has {N} rows but the graph has {graph.num_nodes} nodes")
        p = _sage3_params(D, H, w)
        T = len(t_host)
        sol = torch.empty((T, N, D), dtype=torch.float32, device=y0.device)
        L = _lib.lib()
        fold = bool(L.gnode_set_fold(1))     # read the current setting (set-and-restore)
        L.gnode_set_fold(1 if fold else 0)
        traj = None
        if decoder is not None:
            dec_w, dec_b = decoder
            n_out = dec_w.shape[0]
            traj = torch.empty((T, N, n_out), dtype=torch.float32, device=y0.device)
            ws = _ws(L.gnode_integrate_fixed_decoded_workspace_bytes(N, D, H, method, n_out), y0.device)
        else:
            ws = _ws(L.gnode_integrate_fixed_workspace_bytes(N, D, H, method, 0), y0.device)
        tarr = _float_array(t_host)
        # When a backward will follow, keep the per-stage intermediates (autograd's "tape") so the backward does
        # not recompute every stage -- unless they would not fit comfortably in free device memory.
        save = None
        if any(ctx.needs_input_grad) and T >= 2:   # (also true when called from _IntegrateDecodeFn: same ctx)
            nbytes = int(L.gnode_integrate_fixed_save_bytes(N, D, H, method, T))
            if nbytes > 0 and _save_fits(nbytes, y0.device):
                try:
                    save = torch.empty(nbytes, dtype=torch.uint8, device=y0.device)
                except torch.OutOfMemoryError:      # a nearly full device: the backward recomputes the stages instead
                    save = None
        with torch.cuda.device(y0.device):
            if decoder is not None:
                _lib.check(L.gnode_integrate_fixed_decoded(graph.ref(), C.byref(p), method, _lib.ptr(y0), tarr, T, _lib.ptr(sol),
                                                           _lib.ptr(save), save.numel() if save is not None else 0,
                                                           _lib.ptr(dec_w), _lib.ptr(dec_b), n_out, _lib.ptr(traj),
                                                           _lib.ptr(ws), ws.numel(), _lib.stream_ptr(y0.device)),
                           "gnode_integrate_fixed_decoded")
            else:
                _lib.check(L.gnode_integrate_fixed_flags(graph.ref(), C.byref(p), method, _lib.ptr(y0), tarr, T, _lib.ptr(sol),
                                                         _lib.ptr(save), save.numel() if save is not None else 0,
                                                         _lib.ptr(ws), ws.numel(), 0, _lib.stream_ptr(y0.device)),
                           "gnode_integrate_fixed")
        graph.schedule_tile_check()
        ctx.graph, ctx.method, ctx.t_host, ctx.save = graph, method, t_host, save
        ctx.fold = fold
        if decoder is not None:
            return sol, w, traj
        return sol, w


class _IntegrateFixedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y0, graph: CSRGraph, method: int, t_host: Tuple[float, ...], *w):
        sol, w = _fixed_forward(ctx, y0, graph, method, t_host, w)
        ctx.save_for_backward(sol, *w)
        return sol

    @staticmethod
    def backward(ctx, gsol):
        sol, *w = ctx.saved_tensors
        gsol = _f32(gsol, "grad_solution")
        T, N, D = sol.shape
        H = w[0].shape[0]
        p = _sage3_params(D, H, w)
        gy0 = torch.empty((N, D), dtype=torch.float32, device=sol.device) if ctx.needs_input_grad[0] else None
        gw = _zeros_like_many(w)
        grads = _lib.GnodeSage3Grads(*[t.data_ptr() for t in gw])
        L = _lib.lib()
        prev_fold = L.gnode_set_fold(1 if ctx.fold else 0)   # the save area's layout is the forward's
        try:
            ws = _ws(L.gnode_integrate_fixed_workspace_bytes(N, D, H, ctx.method, 1), sol.device)
            tarr = _float_array(ctx.t_host)
            with torch.cuda.device(sol.device):
                save = ctx.save
                _lib.check(L.gnode_integrate_fixed_bwd(ctx.graph.ref(), C.byref(p), ctx.method, _lib.ptr(sol), tarr, T,
                                                       _lib.ptr(gsol), _lib.ptr(gy0), C.byref(grads), _lib.ptr(save),
                                                       save.numel() if save is not None else 0, _lib.ptr(ws),
                                                       ws.numel(), _lib.stream_ptr(sol.device)),
                           "gnode_integrate_fixed_bwd")
        finally:
            L.gnode_set_fold(prev_fold)
        ctx.save = None
        return (gy0, None, None, None, *gw)


class _IntegrateFixedAdjointFn(torch.autograd.Function):
    """Fixed-grid solve whose backward is the ADJOINT method (torchdiffeq.odeint_adjoint semantics): nothing of the
    forward is kept but the solution; the backward integrates (y, a, dL/dtheta) backwards (gnode_integrate_fixed_adjoint)."""

    @staticmethod
    def forward(ctx, y0, graph: CSRGraph, method: int, t_host: Tuple[float, ...], *w):
        class _NoSave:                      # the shared forward keeps a save area only when a gradient is wanted
            needs_input_grad = (False,)
        holder = _NoSave()
        sol, w = _fixed_forward(holder, y0, graph, method, t_host, w)
        ctx.graph, ctx.method, ctx.t_host = graph, method, t_host
        ctx.save_for_backward(sol, *w)
        return sol

    @staticmethod
    def backward(ctx, gsol):
        sol, *w = ctx.saved_tensors
        gsol = _f32(gsol, "grad_solution")
        T, N, D = sol.shape
        H = w[0].shape[0]
        p = _sage3_params(D, H, w)
        gy0 = torch.empty((N, D), dtype=torch.float32, device=sol.device) if ctx.needs_input_grad[0] else None
        gw = _zeros_like_many(w)
        grads = _lib.GnodeSage3Grads(*[t.data_ptr() for t in gw])
        L = _lib.lib()
        ws = _ws(L.gnode_integrate_fixed_adjoint_workspace_bytes(N, D, H, ctx.method), sol.device)
        tarr = _float_array(ctx.t_host)
        with torch.cuda.device(sol.device):
            _lib.check(L.gnode_integrate_fixed_adjoint(ctx.graph.ref(), C.byref(p), ctx.method, _lib.ptr(sol), tarr, T,
                                                       _lib.ptr(gsol), _lib.ptr(gy0), C.byref(grads), _lib.ptr(ws),
                                                       ws.numel(), _lib.stream_ptr(sol.device)),
                       "gnode_integrate_fixed_adjoint")
        return (gy0, None, None, None, *gw)


def integrate_fixed_adjoint(y0, graph: CSRGraph, params: Sequence[torch.Tensor], t, method: str) -> torch.Tensor:
    """``integrate_fixed`` with the adjoint backward (O(1) memory in the number of steps)."""
    if method not in ("euler", "midpoint", "rk4"):
        raise ValueError(f"not a fixed-grid method: {method}")
    return _IntegrateFixedAdjointFn.apply(y0, graph, METHODS[method], _t_to_host(t), *params)


_T_CACHE: dict = {}


def time_grid_to_host(t, dtype=torch.float32) -> Tuple[float, ...]:
    """Host copy of a time grid.  A CUDA tensor costs one device->host synchronisation the first time it is seen;
    the values are cached by (storage, version), so a training loop that reuses its ``time_span`` tensor pays once."""
    if not torch.is_tensor(t):
        return tuple(float(v) for v in t)
    if not t.is_cuda:
        return tuple(float(v) for v in t.detach().to(dtype).tolist())
    key = (t.data_ptr(), t._version, tuple(t.shape), str(t.device), t.dtype, dtype)
    hit = _T_CACHE.get(key)
    if hit is None:
        vals = tuple(float(v) for v in t.detach().to("cpu", dtype).tolist())
        if len(_T_CACHE) > 64:
            _T_CACHE.clear()
        # the entry keeps the tensor alive: its address cannot be recycled for a different grid while it is cached
        hit = (t.detach(), vals)
        _T_CACHE[key] = hit
    return hit[1]


def _t_to_host(t) -> Tuple[float, ...]:
    return time_grid_to_host(t, torch.float32)


def integrate_fixed(y0, graph: CSRGraph, params: Sequence[torch.Tensor], t, method: str) -> torch.Tensor:
    """Solution [len(t), N, D] of the GraphODEFunc field on the fixed grid ``t``; differentiable."""
    if method not in ("euler", "midpoint", "rk4"):
        raise ValueError(f"not a fixed-grid method: {method}")
    return _IntegrateFixedFn.apply(y0, graph, METHODS[method], _t_to_host(t), *params)


# ----------------------------------------------------------------------------------------------
# adaptive dopri5 (forward only)
# ----------------------------------------------------------------------------------------------
@dataclass
class Dopri5Stats:
    nfe: int = 0
    n_accepted: int = 0
    n_attempted: int = 0
    first_step: float = float("nan")
    last_dt: float = float("nan")
    min_margin: float = float("nan")
    error_ratios: Optional[List[float]] = None
    dts: Optional[List[float]] = None
    accepted: Optional[List[bool]] = None


def _run_dopri5(call: Callable, trace_cap: int):
    st = _lib.GnodeDopri5Stats()
    er = (C.c_double * trace_cap)()
    dts = (C.c_double * trace_cap)()
    acc = (C.c_int32 * trace_cap)()
    tr = _lib.GnodeDopri5Trace(C.cast(er, C.POINTER(C.c_double)), C.cast(dts, C.POINTER(C.c_double)),
                               C.cast(acc, C.POINTER(C.c_int32)), trace_cap)
    call(C.byref(st), C.byref(tr))
    n = min(int(st.n_attempted), trace_cap)
    return Dopri5Stats(nfe=int(st.nfe), n_accepted=int(st.n_accepted), n_attempted=int(st.n_attempted),
                       first_step=float(st.first_step), last_dt=float(st.last_dt), min_margin=float(st.min_margin),
                       error_ratios=[er[i] for i in range(n)], dts=[dts[i] for i in range(n)],
                       accepted=[bool(acc[i]) for i in range(n)])


def _dopri5_forward(y0, graph: CSRGraph, w: Sequence[torch.Tensor], t_host: Sequence[float], rtol: float, atol: float,
                    allreduce, max_num_steps: int, trace_cap: int):
    N, D = y0.shape
    H = w[0].shape[0]
    if N != graph.num_nodes:
        raise GnodeError(f"y0 has {N} rows but the graph has {graph.num_nodes} nodes")
    p = _sage3_params(D, H, w)
    T = len(t_host)
    tarr = (C.c_double * T)(*t_host)
    sol = torch.empty((T, N, D), dtype=torch.float32, device=y0.device)
    L = _lib.lib()
    ws = _ws(L.gnode_integrate_dopri5_workspace_bytes(N, D, H), y0.device)

    if allreduce is not None:
        def _cb(buf, _user):
            s, c = allreduce(buf[0], buf[1])
            buf[0], buf[1] = float(s), float(c)
        cb = _lib.ALLREDUCE_FN(_cb)
    else:
        cb = C.cast(None, _lib.ALLREDUCE_FN)

    # device-side exchange of the norm (``allreduce.device_allreduce(tensor)``: an in-place SUM all-reduce enqueued on the
    # current stream, e.g. dist.dopri5_norm_allreduce): the library hands over the device address of the local sum, a view
    # of the workspace tensor at that address goes to the hook -- no device -> host -> device round trip per attempt
    dev_hook = getattr(allreduce, "device_allreduce", None) if allreduce is not None else None
    dev_err: list = []
    if dev_hook is not None:
        base = ws.data_ptr()

        def _cbd(ptr, _user):
            try:
                off = int(ptr) - base
                dev_hook(ws[off:off + 8].view(torch.float64))
                return 0
            except BaseException as e:      # never unwind through the C frame
                dev_err.append(e)
                return 1
        cbd = _lib.ALLREDUCE_DEV_FN(_cbd)
    else:
        cbd = C.cast(None, _lib.ALLREDUCE_DEV_FN)

    def call(st, tr):
        with torch.cuda.device(y0.device):
            L.gnode_set_dopri5_device_allreduce(cbd, None)
            try:
                rc = L.gnode_integrate_dopri5(graph.ref(), C.byref(p), _lib.ptr(y0), tarr, T, float(rtol),
                                              float(atol), _lib.ptr(sol), st, tr, cb, None, int(max_num_steps),
                                              _lib.ptr(ws), ws.numel(), _lib.stream_ptr(y0.device))
            finally:
                L.gnode_set_dopri5_device_allreduce(C.cast(None, _lib.ALLREDUCE_DEV_FN), None)
            if dev_err:
                raise dev_err[0]
            _lib.check(rc, "gnode_integrate_dopri5")

    stats = _run_dopri5(call, trace_cap)
    graph.schedule_tile_check()
    return sol, stats


class _IntegrateDopri5Fn(torch.autograd.Function):
    """Adaptive solve with backprop through the solver (the reference differentiates torchdiffeq's odeint with plain
    autograd, scripts/train_gde.py:493).  The forward's accepted steps are replayed as a fixed grid with the Dormand-Prince
    tableau; step sizes are constants of the differentiation (``gnode_integrate_dopri5_bwd``)."""

    @staticmethod
    def forward(ctx, y0, graph: CSRGraph, t_host, rtol, atol, allreduce, max_num_steps, trace_cap, holder, *w):
        y0c = _f32(y0.detach(), "y0")
        wc = [_f32(p.detach(), "param") for p in w]
        sol, stats = _dopri5_forward(y0c, graph, wc, t_host, rtol, atol, allreduce, max_num_steps, trace_cap)
        holder.append(stats)
        tau = None
        if stats.n_attempted <= len(stats.dts):
            tau = [float(t_host[0])]
            for dt, acc in zip(stats.dts, stats.accepted):
                if acc:
                    tau.append(tau[-1] + dt)            # the solver's own double-precision t1 = t0 + dt
        ctx.graph, ctx.t_host, ctx.tau = graph, tuple(t_host), tau
        ctx.save_for_backward(y0c, *wc)
        return sol

    @staticmethod
    def backward(ctx, gsol):
        y0, *w = ctx.saved_tensors
        if ctx.tau is None:
            raise GnodeError("dopri5 backward: the forward pass attempted more steps than its trace holds; raise trace_cap")
        gsol = _f32(gsol, "grad_solution")
        N, D = y0.shape
        H = w[0].shape[0]
        p = _sage3_params(D, H, w)
        T, K = len(ctx.t_host), len(ctx.tau) - 1
        gy0 = torch.empty((N, D), dtype=torch.float32, device=y0.device) if ctx.needs_input_grad[0] else None
        gw = _zeros_like_many(w)
        grads = _lib.GnodeSage3Grads(*[t.data_ptr() for t in gw])
        L = _lib.lib()
        prev_fold = L.gnode_set_fold(1)                  # the replay runs on the folded integrator
        try:
            ws = _ws(L.gnode_integrate_dopri5_bwd_workspace_bytes(N, D, H, K), y0.device)
            tau = (C.c_double * (K + 1))(*ctx.tau)
            tarr = (C.c_double * T)(*ctx.t_host)
            with torch.cuda.device(y0.device):
                _lib.check(L.gnode_integrate_dopri5_bwd(ctx.graph.ref(), C.byref(p), _lib.ptr(y0), tau, K, tarr, T,
                                                        _lib.ptr(gsol), _lib.ptr(gy0), C.byref(grads), _lib.ptr(ws),
                                                        ws.numel(), _lib.stream_ptr(y0.device)),
                           "gnode_integrate_dopri5_bwd")
        finally:
            L.gnode_set_fold(prev_fold)
        ctx.graph.schedule_tile_check()
        return (gy0, None, None, None, None, None, None, None, None, *gw)


def integrate_dopri5(y0, graph: CSRGraph, params: Sequence[torch.Tensor], t, rtol: float, atol: float,
                     allreduce: Optional[Callable[[float, float], Tuple[float, float]]] = None,
                     max_num_steps: int = 0, trace_cap: int = 4096):
    """Adaptive Dormand-Prince integration of the GraphODEFunc field.  Returns ``(solution, Dopri5Stats)``.
    Differentiable with respect to ``y0`` and the parameters (backprop through the solver).

    ``allreduce(sumsq, count) -> (sumsq, count)`` (optional) sums the error-norm pieces over
    data-parallel ranks so that all ranks take the step-size decisions of the unsharded batch.
    """
    t_host = list(time_grid_to_host(t, torch.float64))
    if torch.is_grad_enabled() and (y0.requires_grad or any(p.requires_grad for p in params)):
        holder: list = []
        sol = _IntegrateDopri5Fn.apply(y0, graph, t_host, float(rtol), float(atol), allreduce, int(max_num_steps),
                                       int(trace_cap), holder, *params)
        return sol, holder[0]
    y0 = _f32(y0.detach(), "y0")
    w = [_f32(p.detach(), "param") for p in params]
    return _dopri5_forward(y0, graph, w, t_host, rtol, atol, allreduce, max_num_steps, trace_cap)


# ----------------------------------------------------------------------------------------------
# position decoder
# ----------------------------------------------------------------------------------------------
class _DecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        x, w, b = _f32(x, "x"), _f32(w, "weight"), _f32(b, "bias")
        D = x.shape[-1]
        M = x.numel() // D
        n_out = w.shape[0]
        out = torch.empty(x.shape[:-1] + (n_out,), dtype=torch.float32, device=x.device)
        L = _lib.lib()
        with torch.cuda.device(x.device):
            _lib.check(L.gnode_decoder_fwd(_lib.ptr(x), M, D, n_out, _lib.ptr(w), _lib.ptr(b), _lib.ptr(out),
                                           _lib.stream_ptr(x.device)), "gnode_decoder_fwd")
        ctx.save_for_backward(x, w)
        return out

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        g = _f32(g, "grad_out")
        D = x.shape[-1]
        M = x.numel() // D
        n_out = w.shape[0]
        need_x, need_w, need_b = ctx.needs_input_grad
        gx = torch.empty_like(x) if need_x else None
        gw = torch.zeros_like(w) if need_w else None
        gb = torch.zeros(n_out, dtype=torch.float32, device=x.device) if need_b else None
        L = _lib.lib()
        ws = _ws(L.gnode_decoder_workspace_bytes(M, D, n_out), x.device)
        with torch.cuda.device(x.device):
            _lib.check(L.gnode_decoder_bwd(_lib.ptr(x), _lib.ptr(g), M, D, n_out, _lib.ptr(w), _lib.ptr(gx),
                                           _lib.ptr(gw), _lib.ptr(gb), _lib.ptr(ws), ws.numel(),
                                           _lib.stream_ptr(x.device)), "gnode_decoder_bwd")
        return gx, gw, gb


def decode_positions(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """``F.linear(x, weight, bias)`` for a tall-skinny decoder (n_out <= 8) over [..., D]."""
    return _DecoderFn.apply(x, weight, bias)


# ----------------------------------------------------------------------------------------------
# fixed-grid integration + position decoder as one autograd node (GraphODE.forward, scripts/train_gde.py:67-100)
# ----------------------------------------------------------------------------------------------
class _IntegrateDecodeFn(torch.autograd.Function):
    """``solution = odeint(...)`` and ``trajectories = position_decoder(solution)`` in one node, so that the backward
    pass can see HOW the solution reaches the loss.  When only ``trajectories`` carries a cotangent (the training
    loss of scripts/train_gde.py:486-490), the solve has one step and dL/dy_0 is not wanted, the cotangent of the
    solution is the rank-2 product ``g_traj[-1] @ W_dec`` and is handed to the integrator in factored form
    (``gnode_integrate_fixed_bwd_decoded``); every other case takes the general path."""

    @staticmethod
    def forward(ctx, y0, graph: CSRGraph, method: int, t_host: Tuple[float, ...], dec_w, dec_b, *w):
        ctx.set_materialize_grads(False)
        y0 = _f32(y0, "y0")
        dec_w, dec_b = _f32(dec_w, "position_decoder.weight"), _f32(dec_b, "position_decoder.bias")
        if dec_w.shape[0] > 8:
            raise GnodeError("position_decoder wider than 8 outputs is not supported by the fused decoder")
        # the solve and the decoder in one call; fills ctx.{graph,method,t_host,save,fold}
        sol, w, traj = _fixed_forward(ctx, y0, graph, method, t_host, w, decoder=(dec_w, dec_b))
        ctx.save_for_backward(sol, dec_w, *w)
        return sol, traj

    @staticmethod
    def backward(ctx, g_sol, g_traj):
        sol, dec_w, *w = ctx.saved_tensors
        T, N, D = sol.shape
        H = w[0].shape[0]
        n_out = dec_w.shape[0]
        dev = sol.device
        L = _lib.lib()
        need_y0 = ctx.needs_input_grad[0]
        need_dw, need_db = ctx.needs_input_grad[4], ctx.needs_input_grad[5]
        p = _sage3_params(D, H, w)
        want_dw, want_db = need_dw and g_traj is not None, need_db and g_traj is not None
        extra = ([dec_w] if want_dw else []) + ([dec_w.new_empty(n_out)] if want_db else [])
        bufs = _zeros_like_many(list(w) + extra)          # one fill for every parameter gradient of the step
        gw, rest = bufs[:len(w)], bufs[len(w):]
        grads = _lib.GnodeSage3Grads(*[t.data_ptr() for t in gw])
        g_dec_w = rest.pop(0) if want_dw else None
        g_dec_b = rest.pop(0) if want_db else None
        tarr = _float_array(ctx.t_host)
        if g_traj is not None:
            g_traj = _f32(g_traj, "grad_trajectories")
        fast = g_sol is None and g_traj is not None and T == 2 and ctx.fold and not need_y0 and n_out <= 8
        prev_fold = L.gnode_set_fold(1 if ctx.fold else 0)
        try:
            with torch.cuda.device(dev):
                save = ctx.save
                sbytes = save.numel() if save is not None else 0
                if fast:
                    if g_dec_w is not None or g_dec_b is not None:      # decoder parameter gradients (no dL/dsolution)
                        ws = _ws(L.gnode_decoder_workspace_bytes(T * N, D, n_out), dev)
                        _lib.check(L.gnode_decoder_bwd(_lib.ptr(sol), _lib.ptr(g_traj), T * N, D, n_out, _lib.ptr(dec_w),
                                                       None, _lib.ptr(g_dec_w), _lib.ptr(g_dec_b), _lib.ptr(ws), ws.numel(),
                                                       _lib.stream_ptr(dev)), "gnode_decoder_bwd")
                    ws = _ws(L.gnode_integrate_fixed_bwd_decoded_workspace_bytes(N, D, H, ctx.method, n_out), dev)
                    g_last = g_traj[T - 1]                               # contiguous [N, n_out] slice
                    _lib.check(L.gnode_integrate_fixed_bwd_decoded(ctx.graph.ref(), C.byref(p), ctx.method, _lib.ptr(sol),
                                                                   tarr, T, _lib.ptr(g_last), _lib.ptr(dec_w), n_out,
                                                                   C.byref(grads), _lib.ptr(save), sbytes, _lib.ptr(ws),
                                                                   ws.numel(), _lib.stream_ptr(dev)),
                               "gnode_integrate_fixed_bwd_decoded")
                    gy0 = None
                else:
                    if g_traj is not None:
                        gsol = torch.empty_like(sol)
                        ws = _ws(L.gnode_decoder_workspace_bytes(T * N, D, n_out), dev)
                        _lib.check(L.gnode_decoder_bwd(_lib.ptr(sol), _lib.ptr(g_traj), T * N, D, n_out, _lib.ptr(dec_w),
                                                       _lib.ptr(gsol), _lib.ptr(g_dec_w), _lib.ptr(g_dec_b), _lib.ptr(ws),
                                                       ws.numel(), _lib.stream_ptr(dev)), "gnode_decoder_bwd")
                        if g_sol is not None:
                            gsol.add_(_f32(g_sol, "grad_solution"))
                    elif g_sol is not None:
                        gsol = _f32(g_sol, "grad_solution")
                    else:
                        gsol = torch.zeros_like(sol)
                    gy0 = torch.empty((N, D), dtype=torch.float32, device=dev) if need_y0 else None
                    ws = _ws(L.gnode_integrate_fixed_workspace_bytes(N, D, H, ctx.method, 1), dev)
                    _lib.check(L.gnode_integrate_fixed_bwd(ctx.graph.ref(), C.byref(p), ctx.method, _lib.ptr(sol), tarr, T,
                                                           _lib.ptr(gsol), _lib.ptr(gy0), C.byref(grads), _lib.ptr(save),
                                                           sbytes, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)),
                               "gnode_integrate_fixed_bwd")
        finally:
            L.gnode_set_fold(prev_fold)
        ctx.save = None
        return (gy0, None, None, None, g_dec_w, g_dec_b, *gw)


def integrate_fixed_decode(y0, graph: CSRGraph, params: Sequence[torch.Tensor], t, method: str, dec_w, dec_b):
    """``(solution [T, N, D], trajectories [T, N, n_out])`` of GraphODE.forward on a fixed grid; differentiable."""
    if method not in ("euler", "midpoint", "rk4"):
        raise ValueError(f"not a fixed-grid method: {method}")
    return _IntegrateDecodeFn.apply(y0, graph, METHODS[method], _t_to_host(t), dec_w, dec_b, *params)


# ----------------------------------------------------------------------------------------------
# MLP vector field (ODEFunction): forward, and backprop through the field / the solver
# ----------------------------------------------------------------------------------------------
def _mlp_params(w: Sequence[torch.Tensor]):
    w0, b0, w1, b1, w2, b2 = w
    h, H = w0.shape
    if tuple(w1.shape) != (h, h) or tuple(w2.shape) != (H, h) or b0.numel() != h or b1.numel() != h or b2.numel() != H:
        raise GnodeError("ODEFunction parameters have inconsistent shapes")
    return _lib.GnodeMlpParams(H, h, *[t.data_ptr() for t in w]), H, h


def _mlp_rhs_forward(x, w):
    p, H, h = _mlp_params(w)
    M = x.shape[0]
    out = torch.empty_like(x)
    L = _lib.lib()
    ws = _ws(L.gnode_mlp_ode_workspace_bytes(M, H, h, 0), x.device)
    with torch.cuda.device(x.device):
        _lib.check(L.gnode_mlp_rhs_fwd(C.byref(p), _lib.ptr(x), M, _lib.ptr(out), _lib.ptr(ws), ws.numel(),
                                       _lib.stream_ptr(x.device)), "gnode_mlp_rhs_fwd")
    return out


class _MlpRhsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, *w):
        xc = _f32(x.detach(), "x")
        wc = [_f32(p.detach(), "param") for p in w]
        ctx.save_for_backward(xc, *wc)
        return _mlp_rhs_forward(xc, wc)

    @staticmethod
    def backward(ctx, g):
        x, *w = ctx.saved_tensors
        g = _f32(g, "grad_out")
        p, H, h = _mlp_params(w)
        M = x.shape[0]
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gw = _zeros_like_many(w)
        grads = _lib.GnodeMlpGrads(*[t.data_ptr() for t in gw])
        L = _lib.lib()
        ws = _ws(L.gnode_mlp_bwd_workspace_bytes(M, H, h, 0, 1), x.device)
        with torch.cuda.device(x.device):
            _lib.check(L.gnode_mlp_rhs_bwd(C.byref(p), _lib.ptr(x), _lib.ptr(g), M, _lib.ptr(gx), C.byref(grads),
                                           _lib.ptr(ws), ws.numel(), _lib.stream_ptr(x.device)), "gnode_mlp_rhs_bwd")
        return (gx, *gw)


def mlp_rhs(x: torch.Tensor, params: Sequence[torch.Tensor]) -> torch.Tensor:
    """One evaluation of the MLP field; differentiable."""
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params)):
        return _MlpRhsFn.apply(x, *params)
    return _mlp_rhs_forward(_f32(x.detach(), "x"), [_f32(p.detach(), "param") for p in params])


def _mlp_integrate_forward(y0, w, t, method: str, rtol, atol, max_num_steps, trace_cap):
    p, H, h = _mlp_params(w)
    M = y0.shape[0]
    L = _lib.lib()
    m = METHODS[method]
    ws = _ws(L.gnode_mlp_ode_workspace_bytes(M, H, h, m), y0.device)
    if method == "dopri5":
        t_host = [float(v) for v in (t.detach().to("cpu", torch.float64).tolist() if torch.is_tensor(t) else t)]
        T = len(t_host)
        tarr = (C.c_double * T)(*t_host)
        sol = torch.empty((T, M, H), dtype=torch.float32, device=y0.device)

        def call(st, tr):
            with torch.cuda.device(y0.device):
                _lib.check(L.gnode_mlp_integrate_dopri5(C.byref(p), _lib.ptr(y0), M, tarr, T, float(rtol), float(atol),
                                                        _lib.ptr(sol), st, tr, int(max_num_steps), _lib.ptr(ws),
                                                        ws.numel(), _lib.stream_ptr(y0.device)),
                           "gnode_mlp_integrate_dopri5")
        return sol, _run_dopri5(call, trace_cap), t_host
    t_host = _t_to_host(t)
    T = len(t_host)
    sol = torch.empty((T, M, H), dtype=torch.float32, device=y0.device)
    with torch.cuda.device(y0.device):
        _lib.check(L.gnode_mlp_integrate_fixed(C.byref(p), m, _lib.ptr(y0), M, _float_array(t_host), T, _lib.ptr(sol),
                                               _lib.ptr(ws), ws.numel(), _lib.stream_ptr(y0.device)),
                   "gnode_mlp_integrate_fixed")
    return sol, None, t_host


class _MlpIntegrateFn(torch.autograd.Function):
    """``odeint(ODEFunction, y0, t)`` with backprop through the solver (fixed grid, or dopri5 replayed over its accepted
    steps with step sizes as constants -- see _IntegrateDopri5Fn)."""

    @staticmethod
    def forward(ctx, y0, t, method, rtol, atol, max_num_steps, trace_cap, holder, *w):
        y0c = _f32(y0.detach(), "y0")
        wc = [_f32(p.detach(), "param") for p in w]
        sol, stats, t_host = _mlp_integrate_forward(y0c, wc, t, method, rtol, atol, max_num_steps, trace_cap)
        holder.append(stats)
        ctx.method, ctx.t_host, ctx.tau = method, tuple(t_host), None
        if method == "dopri5" and stats.n_attempted <= len(stats.dts):
            tau = [float(t_host[0])]
            for dt, acc in zip(stats.dts, stats.accepted):
                if acc:
                    tau.append(tau[-1] + dt)
            ctx.tau = tau
        ctx.save_for_backward(y0c, sol, *wc)
        return sol

    @staticmethod
    def backward(ctx, gsol):
        y0, sol, *w = ctx.saved_tensors
        gsol = _f32(gsol, "grad_solution")
        p, H, h = _mlp_params(w)
        M, T = y0.shape[0], len(ctx.t_host)
        gy0 = torch.empty_like(y0) if ctx.needs_input_grad[0] else None
        gw = _zeros_like_many(w)
        grads = _lib.GnodeMlpGrads(*[t.data_ptr() for t in gw])
        L = _lib.lib()
        m = METHODS[ctx.method]
        with torch.cuda.device(y0.device):
            if ctx.method == "dopri5":
                if ctx.tau is None:
                    raise GnodeError("dopri5 backward: the forward pass attempted more steps than its trace holds; raise trace_cap")
                K = len(ctx.tau) - 1
                ws = _ws(L.gnode_mlp_bwd_workspace_bytes(M, H, h, m, K), y0.device)
                tau = (C.c_double * (K + 1))(*ctx.tau)
                tarr = (C.c_double * T)(*ctx.t_host)
                _lib.check(L.gnode_mlp_integrate_dopri5_bwd(C.byref(p), _lib.ptr(y0), M, tau, K, tarr, T, _lib.ptr(gsol),
                                                            _lib.ptr(gy0), C.byref(grads), _lib.ptr(ws), ws.numel(),
                                                            _lib.stream_ptr(y0.device)), "gnode_mlp_integrate_dopri5_bwd")
            else:
                ws = _ws(L.gnode_mlp_bwd_workspace_bytes(M, H, h, m, 1), y0.device)
                _lib.check(L.gnode_mlp_integrate_fixed_bwd(C.byref(p), m, _lib.ptr(sol), M, _float_array(ctx.t_host), T,
                                                           _lib.ptr(gsol), _lib.ptr(gy0), C.byref(grads), _lib.ptr(ws),
                                                           ws.numel(), _lib.stream_ptr(y0.device)),
                           "gnode_mlp_integrate_fixed_bwd")
        return (gy0, None, None, None, None, None, None, None, *gw)


def mlp_integrate(y0: torch.Tensor, params: Sequence[torch.Tensor], t, method: str, rtol: float = 1e-7,
                  atol: float = 1e-9, max_num_steps: int = 0, trace_cap: int = 4096):
    """Integrate the MLP field; returns ``(solution [T, M, H], Dopri5Stats | None)``.  Differentiable with respect to
    ``y0`` and the parameters."""
    if torch.is_grad_enabled() and (y0.requires_grad or any(p.requires_grad for p in params)):
        holder: list = []
        sol = _MlpIntegrateFn.apply(y0, t, method, float(rtol), float(atol), int(max_num_steps), int(trace_cap), holder,
                                    *params)
        return sol, holder[0]
    sol, stats, _ = _mlp_integrate_forward(_f32(y0.detach(), "y0"), [_f32(p.detach(), "param") for p in params], t, method,
                                           rtol, atol, max_num_steps, trace_cap)
    return sol, stats


# ----------------------------------------------------------------------------------------------
# dense NT contraction (engine test surface; also the Linear layers of the secondary variants)
# ----------------------------------------------------------------------------------------------
def gemm_nt(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, act: str = "none",
            base: Optional[torch.Tensor] = None, scale: float = 1.0) -> torch.Tensor:
    """``base + scale * act(a @ w.T + bias)`` on the selected engine (no autograd)."""
    a, w = _f32(a.detach(), "a"), _f32(w.detach(), "w")
    m, k = a.shape
    n = w.shape[0]
    out = torch.empty((m, n), dtype=torch.float32, device=a.device)
    L = _lib.lib()
    ws = _ws(L.gnode_gemm_nt_workspace_bytes(n, k), a.device)
    bias_c = _f32(bias.detach(), "bias") if bias is not None else None
    base_c = _f32(base.detach(), "base") if base is not None else None
    with torch.cuda.device(a.device):
        _lib.check(L.gnode_gemm_nt(_lib.ptr(a), k, _lib.ptr(w), k, _lib.ptr(out), n, m, n, k, _lib.ptr(bias_c),
                                   {"none": 0, "relu": 1, "tanh": 2}[act], _lib.ptr(base_c), n, float(scale),
                                   _lib.ptr(ws), ws.numel(), _lib.stream_ptr(a.device)), "gnode_gemm_nt")
    return out


def gemm_tn(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, scale: float = 1.0) -> torch.Tensor:
    """``out += scale * a.T @ b`` (reduction over rows; the weight-gradient contraction) on the selected engine."""
    a, b = _f32(a.detach(), "a"), _f32(b.detach(), "b")
    rows, p = a.shape
    q = b.shape[1]
    if b.shape[0] != rows:
        raise GnodeError("gemm_tn: a and b must have the same number of rows")
    if out is None:
        out = torch.zeros((p, q), dtype=torch.float32, device=a.device)
    L = _lib.lib()
    ws = _ws(L.gnode_gemm_tn_workspace_bytes(p, q, rows), a.device)
    with torch.cuda.device(a.device):
        _lib.check(L.gnode_gemm_tn(_lib.ptr(a), p, _lib.ptr(b), q, _lib.ptr(out), out.stride(0), rows, p, q, float(scale),
                                   _lib.ptr(ws), ws.numel(), _lib.stream_ptr(a.device)), "gnode_gemm_tn")
    return out


def gemm_k128(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, bias_scale: float = 1.0,
              base: Optional[torch.Tensor] = None, base_scale: float = 1.0, base2: Optional[torch.Tensor] = None,
              scale: float = 1.0) -> torch.Tensor:
    """``base_scale * base + base2 + scale * (a @ w.T + bias_scale * bias)`` for a 128-wide reduction on the wide-output
    tcgen05 engine (csrc/gemm_k128.cu); no autograd.  Test surface of the D-wide projections of the integrator."""
    a, w = _f32(a.detach(), "a"), _f32(w.detach(), "w")
    m, k = a.shape
    n = w.shape[0]
    if k != 128 or w.shape[1] != 128:
        raise GnodeError("gemm_k128: the reduction width must be 128")
    out = torch.empty((m, n), dtype=torch.float32, device=a.device)
    L = _lib.lib()
    ws = _ws(L.gnode_gemm_k128_workspace_bytes(n), a.device)
    opt = [None if t is None else _f32(t.detach(), "operand") for t in (bias, base, base2)]
    with torch.cuda.device(a.device):
        _lib.check(L.gnode_gemm_k128(_lib.ptr(a), _lib.ptr(w), _lib.ptr(out), n, m, n, _lib.ptr(opt[0]), float(bias_scale),
                                     _lib.ptr(opt[1]), n, float(base_scale), _lib.ptr(opt[2]), n, float(scale), _lib.ptr(ws),
                                     ws.numel(), _lib.stream_ptr(a.device)), "gnode_gemm_k128")
    return out
