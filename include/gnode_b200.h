/*
 * gnode_b200.h -- C ABI of the B200-native GNODE hot path (libgnode_b200.so, sm_100a only).
 *
 * The reference (dkssud715/swarm-ode) has NO FFI / plugin layer for this path: it is plain
 * nn.Modules calling two third-party Python libraries.  Each entry point below therefore cites
 * the reference call it replaces (file:line under /root/reference) and, where the arithmetic
 * lives in an un-vendored dependency, the upstream routine it stands for:
 *
 *   gnode_csr_build          edge_index consumption inside SAGEConv.propagate
 *                            (scripts/train_gde.py:36,39,43; PyG scatter by dst)      [upstream PyG]
 *   gnode_sage_fwd/_bwd      SAGEConv(in,out)(x, edge_index)   scripts/train_gde.py:27-29,36-43
 *   gnode_rhs_fwd/_bwd       GraphODEFunc.forward(t,x,edge_index)  scripts/train_gde.py:33-45
 *   gnode_integrate_fixed    odeint(..., method='euler'|'midpoint'|'rk4') scripts/train_gde.py:78-85,
 *                            scripts/run_gnode.py:134-135                          [upstream torchdiffeq]
 *   gnode_integrate_fixed_bwd  loss.backward() through the solver  scripts/train_gde.py:493
 *   gnode_integrate_dopri5   odeint(..., method='dopri5') / default method  scripts/gnode.py:136-137,
 *                            scripts/train_gde.py:78-85 with ode_solver='dopri5'   [upstream torchdiffeq]
 *   gnode_decoder_fwd/_bwd   position_decoder over every time point  scripts/train_gde.py:88-94
 *   gnode_mlp_ode_*          ODEFunction.forward(t,x) + odeint   scripts/gnode.py:136-137,160-174,
 *                            scripts/run_gnode.py:134-135,153-167
 *   gnode_spatial_edges      GraphConverter._compute_spatial_edges  scripts/train_gde.py:228-244
 *
 * Conventions
 *   - every pointer documented "device" is a CUDA device pointer on the current device; "host" is
 *     ordinary host memory.  No allocation happens inside the library: the caller passes a
 *     workspace whose size comes from the matching *_workspace_bytes() query.
 *   - all tensors are dense row-major float32 unless stated; node state is [n_nodes, node_dim].
 *   - return value: 0 = ok, negative = error; gnode_last_error() returns a thread-local message.
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); functions documented
 *     "synchronises" wait for the stream before returning.
 *   - no pointer is retained past the call.
 */
#ifndef GNODE_B200_H
#define GNODE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* gnode_stream_t; /* cudaStream_t */

#define GNODE_OK 0
#define GNODE_ERR_ARG (-1)
#define GNODE_ERR_CUDA (-2)
#define GNODE_ERR_WORKSPACE (-3)
#define GNODE_ERR_INDEX (-4)   /* edge index out of range */
#define GNODE_ERR_SOLVER (-5)  /* dt underflow / non-finite state / max steps */

/* solver methods (torchdiffeq names: 'euler', 'midpoint', 'rk4' (3/8 rule), 'dopri5') */
#define GNODE_EULER 0
#define GNODE_MIDPOINT 1
#define GNODE_RK4_38 2
#define GNODE_DOPRI5 3

/* which kernels run the dense contractions */
#define GNODE_ENGINE_AUTO 0  /* tcgen05 (3xTF32) where the shape allows, else SIMT */
#define GNODE_ENGINE_SIMT 1  /* fp32 FFMA everywhere: the parity anchor */
#define GNODE_ENGINE_TC 2    /* force the tcgen05 path (error if the shape is unsupported) */

/* element type of packed node features (gnode_unpack_features) */
#define GNODE_PACK_U8 0
#define GNODE_PACK_I16 1
#define GNODE_PACK_F16 2
#define GNODE_PACK_F32 3
#define GNODE_PACK_BITS 4   /* column-wise bit packing: gnode_unpack_bits */

const char* gnode_last_error(void);
int gnode_abi_version(void);
/* process-wide selection of the GEMM engine (default AUTO); returns the previous value */
int gnode_set_engine(int engine);
/* process-wide selection of how the fixed-grid integrators evaluate the RK stages (default 1); returns the previous
 * value.  1 = folded: conv1 / conv3 are linear, so the stage inputs never materialise D-wide -- two D-wide
 * contractions per STEP (csrc/fold.cu).  0 = direct: every stage evaluates the full field (the straightforward
 * anchor).  Both agree to fp32 rounding.  A forward call with a save area and its backward call must use the same
 * setting. */
int gnode_set_fold(int fold);
/* process-wide switch of the two step-level re-associations of the folded dopri5 (default 1, or GNODE_DOPRI5_FSAL=0 in
 * the environment); returns the previous value.  1 = (a) Z of stage 6 of an accepted step is handed to the next attempt
 * as Z_0 = y_1 @ w1cat^T (FSAL in the folded space; a rejected step keeps its Z_0), so the D-wide input contraction runs
 * once per solve, and (b) an output time inside a step is y0 + (dt sum_s w_s(x) cat2_s) @ w3cat^T with the dense-output
 * weights w(x) of torchdiffeq's quartic (_interp_fit / _interp_evaluate of torchdiffeq/_impl/interp.py, expanded): one
 * D-wide projection.  0 = Z_0 contracted from y on every attempt and the quartic evaluated D-wide in torchdiffeq's
 * operation order.  Both agree to fp32 rounding (stands for the odeint(..., method='dopri5') call at
 * scripts/train_gde.py:78-85). */
int gnode_set_dopri5_fsal(int on);
/* number of kernels this library has launched since load (all threads) */
int64_t gnode_launch_count(void);
/* Synchronises `stream` and reports whether a tcgen05 kernel hit one of its bounded barrier waits
 * since the last call (GNODE_OK = healthy).  For tests / debugging; the hot path never calls it. */
int gnode_tc_status(gnode_stream_t stream);
/* Non-blocking form: enqueues on `stream` a copy of the status word (0 = healthy, otherwise the code of the barrier
 * that expired) into `host_word` (pinned host memory).  The Python layer ships it next to the deferred tile check of
 * every integrate call and raises from graph.poll_pending(); bench.py and scripts/train_gde.py call the blocking form
 * after their timed region / once per epoch. */
int gnode_tc_status_async(int32_t* host_word, gnode_stream_t stream);

/* Per-kernel-class timing (CUDA events on the launching stream) for roofline reporting.  Each entry
 * carries the ALGORITHMIC flops / bytes of the launches it covers (computed from the shapes). */
typedef struct {
  char name[64];
  int64_t launches;
  double ms;     /* summed device time of the class */
  double flops;  /* summed algorithmic flops         */
  double bytes;  /* summed algorithmic bytes (compulsory global reads + writes) */
} gnode_prof_entry;
int gnode_prof_enable(int on);                        /* 1: reset and start collecting, 0: stop */
int gnode_prof_read(gnode_prof_entry* out, int cap);  /* synchronises; returns the number of classes */

/* ------------------------------------------------------------------------------------------
 * Graph: destination-sorted CSR (forward gather) + source-sorted CSR (transpose, for backward).
 * Inside a row, neighbours are in ascending id order, so every reduction order is deterministic.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int64_t n_nodes;
  int64_t n_edges;
  const int32_t* rowptr;   /* device [n_nodes+1]  in-edges of node i: col[rowptr[i]..rowptr[i+1]) */
  const int32_t* col;      /* device [n_edges]    source node ids                                 */
  const int32_t* t_rowptr; /* device [n_nodes+1]  out-edges of node j                             */
  const int32_t* t_col;    /* device [n_edges]    destination node ids                            */
  /* optional (NULL = absent): tiling of the rows into runs of WHOLE graphs of at most tile_rows rows, from
   * gnode_tiles_build[_rows]; lets the integrators run their per-stage chain graph-resident (csrc/chain_fwd.cu,
   * chain_bwd.cu).  tile_err: device int32 set to 1 if an edge is found to leave its tile (the batch was not a
   * disjoint union).  tile_rows: the row limit the tiles were built with (0 = 128; 129 .. 256 for batches whose
   * graphs have up to 256 nodes). */
  const int32_t* tiles;
  int32_t* tile_err;
  int32_t tile_rows;
} gnode_graph;

/* tiles: device int32 [n_graphs + 2]; graph_ptr: device int64 [n_graphs + 1] node offsets of the graphs of a batch
 * (PyG `Batch.ptr`).  tiles[0] = number of tiles (-1 if a graph has more than 128 nodes), tiles[1 + t] = first row of
 * tile t.  Does not synchronise. */
int gnode_tiles_build(const int64_t* graph_ptr, int64_t n_graphs, int32_t* tiles, gnode_stream_t stream);   /* 128 rows */
int gnode_tiles_build_rows(const int64_t* graph_ptr, int64_t n_graphs, int32_t max_rows, int32_t* tiles,
                           gnode_stream_t stream);

size_t gnode_csr_workspace_bytes(int64_t n_nodes, int64_t n_edges);
/* edge_index: device int64 [2, n_edges] (row 0 = source j, row 1 = destination i), any order,
 * as PyG lays it out.  Fills the four CSR arrays.  Synchronises; returns GNODE_ERR_INDEX if any
 * index is outside [0, n_nodes). */
int gnode_csr_build(const int64_t* edge_index, int64_t n_edges, int64_t n_nodes,
                    int32_t* rowptr, int32_t* col, int32_t* t_rowptr, int32_t* t_col,
                    void* workspace, size_t workspace_bytes, gnode_stream_t stream);

/* Same, without synchronising: `error_flag` (device int32, caller-owned) is set to 1 when an index is out of range
 * (such edges are skipped, the arrays stay memory-safe); the caller reads it whenever convenient. */
int gnode_csr_build_async(const int64_t* edge_index, int64_t n_edges, int64_t n_nodes,
                          int32_t* rowptr, int32_t* col, int32_t* t_rowptr, int32_t* t_col, int32_t* error_flag,
                          void* workspace, size_t workspace_bytes, gnode_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Dense "NT" contraction on the selected engine (the per-node Linear layers of SAGEConv /
 * ODEFunction; reference call sites scripts/train_gde.py:27-29, scripts/gnode.py:165-171):
 *   C[m, n] = base[m, n] + scale * act( sum_k A[m, k] * B[n, k] + bias[n] )
 * A: [m, k] row stride lda; B: [n, k] (nn.Linear weight layout) row stride ldb; bias / base may be
 * NULL; act: 0 none, 1 relu, 2 tanh.  Engine AUTO/TC: tcgen05 3xTF32 (fp32-grade), SIMT: fp32 FFMA.
 * ---------------------------------------------------------------------------------------- */
size_t gnode_gemm_nt_workspace_bytes(int32_t n, int32_t k);
int gnode_gemm_nt(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                  int64_t m, int32_t n, int32_t k, const float* bias, int32_t act, const float* base,
                  int64_t ldbase, float scale, void* workspace, size_t workspace_bytes,
                  gnode_stream_t stream);

/* Dense "TN" contraction (the weight gradients of those Linear layers, autograd of
 * scripts/train_gde.py:493):   C[p, q] += scale * sum_r A[r, p] * B[r, q]
 * A: [rows, p] row stride lda, B: [rows, q] row stride ldb, C: [p, q] row stride ldc (accumulated into).
 * Deterministic (fixed-order reduction of per-CTA partials).  Engine AUTO/TC: tcgen05 3xTF32 when both
 * operands are dense (lda == p, ldb == q) and one of them is at most 128 wide; otherwise fp32 FFMA. */
size_t gnode_gemm_tn_workspace_bytes(int32_t p, int32_t q, int64_t rows);
int gnode_gemm_tn(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                  int64_t rows, int32_t p, int32_t q, float scale, void* workspace, size_t workspace_bytes,
                  gnode_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * One SAGEConv layer:  out = mean_{j in N(i)} x_j @ wl^T + bl + x_i @ wr^T   (optional ReLU)
 * wl, wr: [c_out, c_in] (PyG / nn.Linear layout), bl: [c_out].
 * ---------------------------------------------------------------------------------------- */
size_t gnode_sage_workspace_bytes(int64_t n_nodes, int32_t c_in, int32_t c_out);
int gnode_sage_fwd(const gnode_graph* g, const float* x, int32_t c_in, int32_t c_out,
                   const float* wl, const float* bl, const float* wr, int32_t relu,
                   float* out, void* workspace, size_t workspace_bytes, gnode_stream_t stream);
/* Backward of the layer above.  `out` is the layer output (used for the ReLU mask when relu=1).
 * grad_x is overwritten; grad_wl / grad_bl / grad_wr are accumulated into (+=). Any grad pointer
 * may be NULL to skip it. */
int gnode_sage_bwd(const gnode_graph* g, const float* x, const float* out, const float* grad_out,
                   int32_t c_in, int32_t c_out, const float* wl, const float* wr, int32_t relu,
                   float* grad_x, float* grad_wl, float* grad_bl, float* grad_wr,
                   void* workspace, size_t workspace_bytes, gnode_stream_t stream);

/* Bipartite SAGEConv forward, one HeteroConv relation (scripts/gnode.py:92-99,126-128; scripts/run_gnode.py:89-96):
 *   out = post( accum + scale * ( mean_{j in N(i)} x_src[j] @ wl^T + bl + x_dst[i] @ wr^T ) ),   post = relu | id
 * g: CSR whose first n_dst rows are the destination nodes and whose column ids index rows of x_src (build it with
 * n_nodes >= max(n_src, n_dst)).  accum (may be NULL) / scale fold HeteroConv(aggr='mean') over the relations of one
 * destination type -- and the ReLU after it -- into the relation calls.  out may alias accum. */
size_t gnode_sage_bipartite_workspace_bytes(int64_t n_dst, int32_t c_in, int32_t c_out);
int gnode_sage_bipartite_fwd(const gnode_graph* g, int64_t n_dst, const float* x_src, const float* x_dst,
                             int32_t c_in, int32_t c_out, const float* wl, const float* bl, const float* wr,
                             float scale, const float* accum, int32_t post_relu, float* out,
                             void* workspace, size_t workspace_bytes, gnode_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * The GNODE vector field: three SAGE layers, ReLU between them (GraphODEFunc).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t node_dim;   /* D */
  int32_t hidden_dim; /* H */
  const float *w1l, *b1, *w1r; /* conv1: [H,D], [H], [H,D] */
  const float *w2l, *b2, *w2r; /* conv2: [H,H], [H], [H,H] */
  const float *w3l, *b3, *w3r; /* conv3: [D,H], [D], [D,H] */
} gnode_sage3_params;

typedef struct { /* same shapes as gnode_sage3_params; accumulated into (+=); device */
  float *w1l, *b1, *w1r, *w2l, *b2, *w2r, *w3l, *b3, *w3r;
} gnode_sage3_grads;

size_t gnode_rhs_workspace_bytes(int64_t n_nodes, int32_t node_dim, int32_t hidden_dim);
int gnode_rhs_fwd(const gnode_graph* g, const gnode_sage3_params* p, const float* x, float* dxdt,
                  void* workspace, size_t workspace_bytes, gnode_stream_t stream);
/* vector-Jacobian product of the field at x (forward is recomputed inside). grad_x overwritten. */
int gnode_rhs_bwd(const gnode_graph* g, const gnode_sage3_params* p, const float* x,
                  const float* grad_out, float* grad_x, const gnode_sage3_grads* grads,
                  void* workspace, size_t workspace_bytes, gnode_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fixed-grid integration (grid == t, as at every reference call site).
 *   t: host float32 [n_t], strictly increasing.   sol: device [n_t, n_nodes, D]; sol[0] = y0.
 * ---------------------------------------------------------------------------------------- */
size_t gnode_integrate_fixed_workspace_bytes(int64_t n_nodes, int32_t node_dim, int32_t hidden_dim,
                                             int32_t method, int32_t backward);
/* Bytes of the optional `save` area: per solver step the stage inputs and the per-stage layer
 * intermediates (what autograd would keep on its tape).  0 for n_t < 2. */
size_t gnode_integrate_fixed_save_bytes(int64_t n_nodes, int32_t node_dim, int32_t hidden_dim,
                                        int32_t method, int32_t n_t);
/* save: NULL for a plain forward; otherwise a device buffer of >= *_save_bytes() that the matching
 * gnode_integrate_fixed_bwd call reads instead of recomputing every stage. */
int gnode_integrate_fixed(const gnode_graph* g, const gnode_sage3_params* p, int32_t method,
                          const float* y0, const float* t, int32_t n_t, float* sol,
                          void* save, size_t save_bytes,
                          void* workspace, size_t workspace_bytes, gnode_stream_t stream);
/* flags: GNODE_FIXED_SOL0_BY_CALLER -- sol[0] is NOT written (the first step reads y0 itself); the caller fills it before
 * anything reads the solution (gnode_decoder_fwd_copy does, as a side effect of decoding the first time point). */
#define GNODE_FIXED_SOL0_BY_CALLER 1
int gnode_integrate_fixed_flags(const gnode_graph* g, const gnode_sage3_params* p, int32_t method,
                                const float* y0, const float* t, int32_t n_t, float* sol,
                                void* save, size_t save_bytes,
                                void* workspace, size_t workspace_bytes, int32_t flags, gnode_stream_t stream);
/* GraphODE.forward on a fixed grid in ONE call (scripts/train_gde.py:67-100: odeint(...) at :78-85, then position_decoder
 * over every time point at :88-94): sol [n_t, n_nodes, D] and traj [n_t, n_nodes, n_out] (dec_w [n_out, D], dec_b [n_out]
 * in nn.Linear layout).  sol[0] = y0 is written while y0 streams through the decoder of the first time point; with the
 * folded integrator, 2H = 128 and n_out <= 4 every later time point is decoded from the previous one and the step's
 * 2H-wide stage combination, traj[j+1] = traj[j] + C_j @ (dec_w @ w3cat)^T + (dt_j sum c)(dec_w @ b3), so the D-wide
 * solution is written once and not read back (fp32 rounding order only); otherwise the decoder runs over sol[1:].
 * save / save_bytes as for gnode_integrate_fixed; sol must not alias y0. */
size_t gnode_integrate_fixed_decoded_workspace_bytes(int64_t n_nodes, int32_t node_dim, int32_t hidden_dim,
                                                     int32_t method, int32_t n_out);
int gnode_integrate_fixed_decoded(const gnode_graph* g, const gnode_sage3_params* p, int32_t method,
                                  const float* y0, const float* t, int32_t n_t, float* sol,
                                  void* save, size_t save_bytes,
                                  const float* dec_w, const float* dec_b, int32_t n_out, float* traj,
                                  void* workspace, size_t workspace_bytes, gnode_stream_t stream);
/* Backprop through the solver (discretise-then-optimise, like autograd through torchdiffeq's
 * fixed-grid loop).  sol is the forward output; grad_sol: device [n_t, n_nodes, D] (cotangent of
 * every saved time point); grad_y0 overwritten (NULL = not wanted: its D-wide contraction is skipped); param grads
 * accumulated (+=).  With save == NULL
 * each step's stages are recomputed from sol[j]; with the forward's save area they are not. */
/* Opt-in ADJOINT backward of the fixed-grid solvers (torchdiffeq.odeint_adjoint semantics; the reference itself trains
 * with plain odeint, /root/reference/scripts/train_gde.py:78-85, imported at :10): the augmented system (y, a, dL/dtheta)
 * is integrated backwards over every output interval with the same Runge-Kutta scheme, y being reset to the stored sol[i]
 * at every output time.  Needs only the forward's solution (no save area); same arguments as gnode_integrate_fixed_bwd. */
size_t gnode_integrate_fixed_adjoint_workspace_bytes(int64_t n_nodes, int32_t node_dim, int32_t hidden_dim, int32_t method);
int gnode_integrate_fixed_adjoint(const gnode_graph* g, const gnode_sage3_params* p, int32_t method,
                                  const float* sol, const float* t, int32_t n_t, const float* grad_sol,
                                  float* grad_y0, const gnode_sage3_grads* grads,
                                  void* workspace, size_t workspace_bytes, gnode_stream_t stream);
int gnode_integrate_fixed_bwd(const gnode_graph* g, const gnode_sage3_params* p, int32_t method,
                              const float* sol, const float* t, int32_t n_t, const float* grad_sol,
                              float* grad_y0, const gnode_sage3_grads* grads,
                              const void* save, size_t save_bytes,
                              void* workspace, size_t workspace_bytes, gnode_stream_t stream);

/* Backward of a ONE-STEP fixed-grid solve (n_t == 2, folded integrator) whose solution reaches the loss only through
 * position_decoder at the last time point -- the training step of scripts/train_gde.py:486-493.  The cotangent of y_1
 * is grad_traj_last [n_nodes, n_out] @ dec_w [n_out, D] (rank n_out <= 8): it is never formed, and neither are its two
 * D-wide contractions.  Param grads accumulated (+=); dL/dy_0 is not produced. */
size_t gnode_integrate_fixed_bwd_decoded_workspace_bytes(int64_t n_nodes, int32_t node_dim, int32_t hidden_dim,
                                                         int32_t method, int32_t n_out);
int gnode_integrate_fixed_bwd_decoded(const gnode_graph* g, const gnode_sage3_params* p, int32_t method,
                                      const float* sol, const float* t, int32_t n_t,
                                      const float* grad_traj_last, const float* dec_w, int32_t n_out,
                                      const gnode_sage3_grads* grads, const void* save, size_t save_bytes,
                                      void* workspace, size_t workspace_bytes, gnode_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Adaptive Dormand-Prince 5(4), torchdiffeq semantics (global RMS error norm over the whole
 * state tensor, Hairer initial step, FSAL, quartic dense output, fp64 time / fp32 state).
 * Synchronises once per attempted step (the accept/reject decision is made on the host).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int64_t nfe;          /* vector-field evaluations                       */
  int64_t n_accepted;   /* steps with error_ratio <= 1                    */
  int64_t n_attempted;
  double first_step;    /* dt chosen by the initial-step heuristic        */
  double last_dt;       /* dt proposed after the final step               */
  double min_margin;    /* min over attempts of |error_ratio - 1|         */
} gnode_dopri5_stats;

/* Optional per-attempt trace, host arrays of capacity trace_cap (may be NULL / 0). */
typedef struct {
  double* error_ratio;
  double* dt;
  int32_t* accepted;
  int64_t trace_cap;
} gnode_dopri5_trace;

/* Optional cross-rank hook: called on the host with the local (sum of squares, element count)
 * of a norm; must return them summed over all ranks (torch.distributed all-reduce on the Python
 * side).  NULL = single rank. */
typedef void (*gnode_allreduce_fn)(double* sumsq_and_count /* [2], in/out */, void* user);
/* Device-side form of the same exchange, registered for the CALLING THREAD (NULL clears it) and used by its following
 * gnode_integrate_dopri5 / gnode_mlp_integrate_dopri5 calls: per norm the library hands the hook the DEVICE address of the
 * local sum of squares (one double inside the call's workspace); the hook enqueues an in-place SUM all-reduce over the
 * data-parallel ranks on the solve's stream (ncclAllReduce / torch.distributed.all_reduce) and returns 0.  The host then
 * reads the global sum with the one synchronisation per attempted step that the step-size decision needs -- no
 * device -> host -> device round trip for the exchange.  The element count is summed once per solve through the host hook
 * (gnode_allreduce_fn, which must be given as well). */
typedef int (*gnode_allreduce_dev_fn)(double* device_sumsq /* [1], in/out, device memory */, void* user);
int gnode_set_dopri5_device_allreduce(gnode_allreduce_dev_fn fn, void* user);

size_t gnode_integrate_dopri5_workspace_bytes(int64_t n_nodes, int32_t node_dim, int32_t hidden_dim);
int gnode_integrate_dopri5(const gnode_graph* g, const gnode_sage3_params* p, const float* y0,
                           const double* t, int32_t n_t, double rtol, double atol, float* sol,
                           gnode_dopri5_stats* stats, const gnode_dopri5_trace* trace,
                           gnode_allreduce_fn allreduce, void* allreduce_user,
                           int64_t max_num_steps, void* workspace, size_t workspace_bytes,
                           gnode_stream_t stream);

/* Backward through gnode_integrate_dopri5 -- loss.backward() through torchdiffeq's odeint (no adjoint), as in
 * scripts/train_gde.py:493 with GraphODE(ode_solver='dopri5').  tau[0 .. n_accepted] are the accepted step times of the
 * forward pass (tau[0] = t[0]; rebuilt from its trace: tau[k+1] = tau[k] + dt of the k-th accepted attempt).  Step sizes
 * are constants of the differentiation (torchdiffeq's step-size controller runs under no_grad).  grad_sol: [n_t, N, D];
 * grad_y0 overwritten (may be NULL); parameter gradients accumulated (+=) into grads.  Folded integrator only. */
size_t gnode_integrate_dopri5_bwd_workspace_bytes(int64_t n_nodes, int32_t node_dim, int32_t hidden_dim,
                                                  int32_t n_accepted);
int gnode_integrate_dopri5_bwd(const gnode_graph* g, const gnode_sage3_params* p, const float* y0,
                               const double* tau, int32_t n_accepted, const double* t, int32_t n_t,
                               const float* grad_sol, float* grad_y0, const gnode_sage3_grads* grads,
                               void* workspace, size_t workspace_bytes, gnode_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * position_decoder = Linear(D, 2) applied to every row of the solution.
 *   x: [m, D] (m = n_t * n_nodes), w: [n_out, D], b: [n_out], out: [m, n_out]; n_out <= 8.
 * ---------------------------------------------------------------------------------------- */
size_t gnode_decoder_workspace_bytes(int64_t m, int32_t node_dim, int32_t n_out);
int gnode_decoder_fwd(const float* x, int64_t m, int32_t node_dim, int32_t n_out, const float* w,
                      const float* b, float* out, gnode_stream_t stream);
/* grad_x overwritten (may be NULL); grad_w / grad_b accumulated (+=, may be NULL). */
/* The same, and the rows of x are also written to copy_out [m, node_dim] as they stream through (NULL: no copy).  Used for
 * the first time point: `sol[0] = y0` of the solver (gnode_integrate_fixed_flags with GNODE_FIXED_SOL0_BY_CALLER) then costs no
 * pass of its own. */
int gnode_decoder_fwd_copy(const float* x, int64_t m, int32_t node_dim, int32_t n_out, const float* w,
                           const float* b, float* out, float* copy_out, gnode_stream_t stream);
int gnode_decoder_bwd(const float* x, const float* grad_out, int64_t m, int32_t node_dim,
                      int32_t n_out, const float* w, float* grad_x, float* grad_w, float* grad_b,
                      void* workspace, size_t workspace_bytes, gnode_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Node-wise MLP vector field  Linear(H,h) tanh Linear(h,h) tanh Linear(h,H)  (ODEFunction),
 * integrated with the same solvers (no graph).  Weights in nn.Linear layout.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t dim;        /* H  */
  int32_t hidden_dim; /* h  */
  const float *w0, *b0; /* [h,H], [h] */
  const float *w1, *b1; /* [h,h], [h] */
  const float *w2, *b2; /* [H,h], [H] */
} gnode_mlp_params;

typedef struct {
  float *w0, *b0, *w1, *b1, *w2, *b2; /* same shapes as gnode_mlp_params; any pointer may be NULL */
} gnode_mlp_grads;

size_t gnode_mlp_ode_workspace_bytes(int64_t m, int32_t dim, int32_t hidden_dim, int32_t method);
int gnode_mlp_rhs_fwd(const gnode_mlp_params* p, const float* x, int64_t m, float* dxdt,
                      void* workspace, size_t workspace_bytes, gnode_stream_t stream);
int gnode_mlp_integrate_fixed(const gnode_mlp_params* p, int32_t method, const float* y0, int64_t m,
                              const float* t, int32_t n_t, float* sol, void* workspace,
                              size_t workspace_bytes, gnode_stream_t stream);
int gnode_mlp_integrate_dopri5(const gnode_mlp_params* p, const float* y0, int64_t m, const double* t,
                               int32_t n_t, double rtol, double atol, float* sol,
                               gnode_dopri5_stats* stats, const gnode_dopri5_trace* trace,
                               int64_t max_num_steps, void* workspace, size_t workspace_bytes,
                               gnode_stream_t stream);

/* The D-wide projections of the folded integrator on their own (parity tests of the engine, csrc/gemm_k128.cu):
 *   C[m, n] = base_scale * base + base2 + scale * (A[m, :128] . B[n, :128] + bias_scale * bias[n])
 * A: [m, 128] dense rows (16-byte aligned); B: [n, 128]; C / base / base2: row strides ldc / ldbase / ldbase2, any
 * alignment; bias, base, base2 may be NULL (base2 needs base).  With dense rows (ldc == ldbase == n <= 400), one base
 * term and 16-byte aligned C / base the row-major kernel runs (k_gemm_k128_rows: the y_1 projection of the fixed-grid
 * solvers), otherwise the column-chunked one. */
size_t gnode_gemm_k128_workspace_bytes(int32_t n);
int gnode_gemm_k128(const float* A, const float* B, float* C, int64_t ldc, int64_t m, int32_t n, const float* bias,
                    float bias_scale, const float* base, int64_t ldbase, float base_scale, const float* base2,
                    int64_t ldbase2, float scale, void* workspace, size_t workspace_bytes, gnode_stream_t stream);

/* Whole-episode graph construction on the device: all T window graphs of one episode as their disjoint
 * union -- one GraphConverter applied step by step (scripts/train_gde.py:116-184 via :308-314) followed by
 * Batch.from_data_list (:367), bit for bit.  obs: device f32 [n_steps, n_agents, node_dim] (rows < num_agvs are AGVs:
 * position columns (3, 4), pickers (0, 1)).  Outputs (device): x [nodes, node_dim]; edge_index int64 [2, edge_capacity]
 * (row 0 = sources, row 1 = targets; the first edge_offsets[n_steps] columns are valid); batch int64 [nodes];
 * is_current_agent uint8/bool [nodes]; ptr int64 [n_steps + 1]; edge_offsets int64 [n_steps + 1] (first edge of every
 * graph).  nodes = gnode_window_graphs_nodes(...), edge_capacity >= gnode_window_graphs_edge_capacity(...). */
int64_t gnode_window_graphs_nodes(int64_t n_steps, int32_t n_agents, int32_t window);
int64_t gnode_window_graphs_edge_capacity(int64_t n_steps, int32_t n_agents, int32_t window);
size_t gnode_window_graphs_workspace_bytes(int64_t n_steps, int32_t n_agents);
int gnode_window_graphs(const float* obs, int64_t n_steps, int32_t n_agents, int32_t node_dim, int32_t num_agvs,
                        float threshold, int32_t window, float* x, int64_t* edge_index, int64_t edge_capacity,
                        int64_t* batch, uint8_t* is_current_agent, int64_t* ptr, int64_t* edge_offsets,
                        void* workspace, size_t workspace_bytes, gnode_stream_t stream);

/* Backward of the bipartite SAGEConv relation of HeteroConv (scripts/gnode.py:92-99,126-128) and of the Linear
 * embeddings / action heads (:84-90,104-115): what loss.backward() runs when the Q-network is trained.
 *   gs = scale * grad_out * [out > 0]  (mask only when `out` is given: the ReLU folded into the last relation)
 *   grad_wl += gs^T mean_j x_src, grad_bl += colsum(gs), grad_wr += gs^T x_dst          (accumulated, may be NULL)
 *   grad_x_dst = gs @ wr, grad_x_src = A^T(gs @ wl)                                      (overwritten, may be NULL) */
size_t gnode_sage_bipartite_bwd_workspace_bytes(int64_t n_dst, int32_t c_in, int32_t c_out);
int gnode_sage_bipartite_bwd(const gnode_graph* g, int64_t n_src, int64_t n_dst, const float* x_src,
                             const float* x_dst, int32_t c_in, int32_t c_out, const float* wl, const float* wr,
                             const float* grad_out, const float* out, float scale, float* grad_x_src,
                             float* grad_x_dst, float* grad_wl, float* grad_bl, float* grad_wr,
                             void* workspace, size_t workspace_bytes, gnode_stream_t stream);
/* out = act(x @ w^T + b): grad_x = gm @ w (overwritten), grad_w += gm^T x, grad_b += colsum(gm),
 * gm = grad_out * [out > 0] when `out` is given (ReLU), else grad_out.  x [m, c_in], w [c_out, c_in]. */
size_t gnode_linear_bwd_workspace_bytes(int64_t m, int32_t c_in, int32_t c_out);
int gnode_linear_bwd(const float* x, const float* w, const float* out, const float* grad_out, int64_t m,
                     int32_t c_in, int32_t c_out, float* grad_x, float* grad_w, float* grad_b, void* workspace,
                     size_t workspace_bytes, gnode_stream_t stream);

/* Backward of the MLP field and of its solves -- loss.backward() through ODEFunction / odeint in the Q-network
 * training of scripts/gnode.py:136-137,160-174 and scripts/run_gnode.py:134-135 (plain autograd, no adjoint).
 * Parameter gradients are accumulated (+=); grad_x / grad_y0 are overwritten (may be NULL).  The dopri5 form replays
 * the accepted steps tau[0 .. n_accepted] of the forward pass like gnode_integrate_dopri5_bwd. */
size_t gnode_mlp_bwd_workspace_bytes(int64_t m, int32_t dim, int32_t hidden_dim, int32_t method, int32_t n_accepted);
int gnode_mlp_rhs_bwd(const gnode_mlp_params* p, const float* x, const float* grad_out, int64_t m, float* grad_x,
                      const gnode_mlp_grads* grads, void* workspace, size_t workspace_bytes, gnode_stream_t stream);
int gnode_mlp_integrate_fixed_bwd(const gnode_mlp_params* p, int32_t method, const float* sol, int64_t m,
                                  const float* t, int32_t n_t, const float* grad_sol, float* grad_y0,
                                  const gnode_mlp_grads* grads, void* workspace, size_t workspace_bytes,
                                  gnode_stream_t stream);
int gnode_mlp_integrate_dopri5_bwd(const gnode_mlp_params* p, const float* y0, int64_t m, const double* tau,
                                   int32_t n_accepted, const double* t, int32_t n_t, const float* grad_sol,
                                   float* grad_y0, const gnode_mlp_grads* grads, void* workspace,
                                   size_t workspace_bytes, gnode_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Spatial edges of GraphConverter, batched over independent snapshots, bit-exact:
 *   pos: device float32 [n_snap, n_agents, 2] (y, x).  For every snapshot, for i < j in
 *   lexicographic order, if sqrtf((dy*dy) + (dx*dx)) < threshold (strict, float32) emit (i,j)
 *   then (j,i).  counts[s] receives the number of directed edges of snapshot s; edges (local
 *   ids, int32 pairs src,dst) are written at edges + 2 * s * n_agents*(n_agents-1).
 * ---------------------------------------------------------------------------------------- */
int gnode_spatial_edges(const float* pos, int64_t n_snap, int32_t n_agents, float threshold,
                        int32_t* counts, int32_t* edges, gnode_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Lossless narrow transport of a host batch -- `batch.to(device)` of scripts/train_gde.py:475.
 * The host packs node features in the narrowest type that reproduces every fp32 value exactly
 * (warehouse observations are small integers) and edge lists as int32; these entry points widen
 * them on the device into the fp32 `x` / int64 `edge_index` / int64 `batch` of the reference
 * contract, bit-exact.  src / dst: device memory, 16-byte aligned; n: elements.
 * ---------------------------------------------------------------------------------------- */
int gnode_unpack_features(const void* src, int32_t kind /* GNODE_PACK_* */, int64_t n, float* dst, gnode_stream_t stream);
/* Column-wise bit-packed features (swarm_ode_b200/data.py:pack_bits): src [rows, row_bytes] bytes, column c of a row at bits
 * [bit_offsets[c], bit_offsets[c + 1]) of the row's little-endian bit string (widths <= 8; device int32 [cols + 1]);
 * row_bytes a multiple of 16 with at least one spare byte after the last bit.  dst: fp32 [rows, cols] -- what
 * `batch.x.to(device)` of the reference delivers (scripts/train_gde.py:475), bit for bit. */
int gnode_unpack_bits(const void* src, int64_t rows, int32_t cols, int32_t row_bytes, const int32_t* bit_offsets,
                      float* dst, gnode_stream_t stream);
int gnode_unpack_edges(const int32_t* src, int64_t n, int64_t* dst, gnode_stream_t stream);
/* batch[i] = index of the graph that owns node i, from the graph offsets ptr [n_graphs + 1] (PyG Batch.batch / Batch.ptr) */
int gnode_batch_vector(const int64_t* ptr, int64_t n_graphs, int64_t n_nodes, int64_t* batch, gnode_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GNODE_B200_H */
