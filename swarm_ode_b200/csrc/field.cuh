// Vector fields the integrators run on, and the Butcher tableaus of the solvers the reference can
// select (torchdiffeq 'euler', 'midpoint', 'rk4' = 3/8 rule, 'dopri5').
#pragma once
#include "common.cuh"
#include "rk.cuh"

namespace gnode {

constexpr int kMaxStages = 7;

struct Tableau {
  int S;                                  // stages
  double beta[kMaxStages][kMaxStages];    // stage s input = y + dt * sum_{j<s} beta[s][j] k_j
  double c_sol[kMaxStages];
  double c_err[kMaxStages];
  double c_mid[kMaxStages];
};
const Tableau* tableau_for(int method);

// A vector field f(x) over a dense [rows, dim] fp32 state.
struct Field {
  virtual ~Field() {}
  virtual int64_t rows() const = 0;
  virtual int dim() const = 0;
  // out = scale * f(x) (+ base if base != null).  `slot` selects where intermediates are kept for a
  // later vjp (0 when nothing needs to be kept).  out may alias base.
  virtual int eval(const float* x, float* out, const float* base, float scale, int slot, cudaStream_t s) = 0;
  // gx = J_f(x)^T gk using the intermediates that eval(..., slot) kept; parameter gradients accumulate inside the field
  virtual int vjp(const float* x, int slot, const float* gk, float* gx, cudaStream_t s) {
    (void)x; (void)slot; (void)gk; (void)gx; (void)s;
    set_error("this vector field has no backward");
    return GNODE_ERR_ARG;
  }
  int64_t numel() const { return rows() * (int64_t)dim(); }
};

// ---- GraphODEFunc: three SAGE layers (scripts/train_gde.py:20-45) ----
struct Sage3Ctx : Field {
  gnode_graph g{};
  const int32_t* g_tiles = nullptr;    // whole-graph row tiles (gnode_graph.tiles), null when the batch has none
  int32_t* g_tile_err = nullptr;
  int D = 0, H = 0;
  int64_t N = 0;
  const float *b1 = nullptr, *b2 = nullptr, *b3 = nullptr;
  // packed weights (workspace): w1cat [2H, D] = [w1l; w1r], w2cat [H, 2H] = [w2l | w2r], w3cat [D, 2H] = [w3l | w3r]
  float *w1cat = nullptr, *w2cat = nullptr, *w3cat = nullptr;
  // transposes for the data gradients (NT form): w3catT [2H, D], w2catT [2H, H], w1catT [D, 2H]
  float *w1catT = nullptr, *w2catT = nullptr, *w3catT = nullptr;
  // tf32 hi/lo planes of the six packed matrices for the tcgen05 engine (null -> FFMA engine)
  float *s1 = nullptr, *s2 = nullptr, *s3 = nullptr, *s1T = nullptr, *s2T = nullptr, *s3T = nullptr;
  float *ci2 = nullptr, *ci2T = nullptr;   // chain-kernel images (chain_common.cuh) of w2cat [H x 2H] and w2catT [2H x H]
  float *ck3 = nullptr, *ck1T = nullptr;   // chunked images (gemm_k128.cu) of w3cat [D x 2H] and w1catT [D x 2H]
  bool use_tc = false;
  // images not packed yet (pack() only marks them; the first reader packs: gemm_nt through GemmNT::*_pending, the chain
  // kernels through ensure_chain_images)
  int pend_s1 = 0, pend_s2 = 0, pend_s3 = 0, pend_s1T = 0, pend_s2T = 0, pend_s3T = 0;
  int pend_ci2 = 0, pend_ci2T = 0, pend_ck3 = 0, pend_ck1T = 0;
  void use_w1(GemmNT& q) { q.Bsplit = use_tc ? s1 : nullptr; q.Bsplit_pending = &pend_s1; }
  void use_w2(GemmNT& q) { q.Bsplit = use_tc ? s2 : nullptr; q.Bsplit_pending = &pend_s2; }
  void use_w3(GemmNT& q) { q.Bsplit = use_tc ? s3 : nullptr; q.Bsplit_pending = &pend_s3; q.Bchain = use_tc ? ck3 : nullptr; q.Bchain_pending = &pend_ck3; }
  void use_w1T(GemmNT& q) { q.Bsplit = use_tc ? s1T : nullptr; q.Bsplit_pending = &pend_s1T; q.Bchain = use_tc ? ck1T : nullptr; q.Bchain_pending = &pend_ck1T; }
  void use_w2T(GemmNT& q) { q.Bsplit = use_tc ? s2T : nullptr; q.Bsplit_pending = &pend_s2T; }
  void use_w3T(GemmNT& q) { q.Bsplit = use_tc ? s3T : nullptr; q.Bsplit_pending = &pend_s3T; }
  bool skip_wgrad = false;             // vjp: data gradient only (adjoint stages whose solution weight is zero)
  float* z = nullptr;                  // [N, 2H]  x @ w1cat^T
  int n_slots = 1;
  float* cat1[kMaxStages] = {};        // [N, 2H]  [ mean(h1) | h1 ]
  float* cat2[kMaxStages] = {};        // [N, 2H]  [ mean(h2) | h2 ]
  // backward scratch
  float *gcat = nullptr, *gz = nullptr, *gv2 = nullptr, *partials = nullptr, *colpart = nullptr;
  float *dW1cat = nullptr, *dW2cat = nullptr, *dW3cat = nullptr, *db1 = nullptr, *db2 = nullptr, *db3 = nullptr;

  int64_t rows() const override { return N; }
  int dim() const override { return D; }
  int eval(const float* x, float* out, const float* base, float scale, int slot, cudaStream_t s) override;

  // carve the workspace (measuring when the arena has no base)
  void carve(Arena& a, int slots, bool backward);
  int pack(const gnode_sage3_params& p, bool backward, cudaStream_t s);
  int zero_param_grads(cudaStream_t s);
  // grad_x = J_f(x)^T gk using the intermediates kept in `slot`; parameter gradients accumulate into dW*/db*
  int vjp(const float* x, int slot, const float* gk, float* gx, cudaStream_t s) override;
  int unpack_grads(const gnode_sage3_grads& gr, cudaStream_t s);
};

// ---- folded fixed-grid integration (fold.cu): D-wide contractions once per step instead of once per stage ----
struct FoldWs {
  int S = 0;
  bool forward_only = false;           // no backward will read this solve's stage slots: the stage kernel skips cat1 / sign bits
  float *M13 = nullptr, *M13T = nullptr, *c13 = nullptr;   // w1cat @ w3cat [2H, 2H], its transpose, w1cat @ b3 [2H]
  float *sM13 = nullptr, *sM13T = nullptr;                 // tf32 hi/lo planes of the two
  float *ci13 = nullptr, *ci13T = nullptr;                 // chain-kernel images of the two
  int pend_sM13 = 0, pend_sM13T = 0, pend_ci13 = 0, pend_ci13T = 0;   // not packed yet (prepare() only marks them)
  void use_M13(const Sage3Ctx& c, GemmNT& q) { q.Bsplit = c.use_tc ? sM13 : nullptr; q.Bsplit_pending = &pend_sM13; }
  void use_M13T(const Sage3Ctx& c, GemmNT& q) { q.Bsplit = c.use_tc ? sM13T : nullptr; q.Bsplit_pending = &pend_sM13T; }
  float *z0 = nullptr, *Cbuf = nullptr;                    // [N, 2H]
  float* Cslot = nullptr;                                  // C = dt sum_s c_s cat2_s of the current step (save area or Cbuf)
  // FSAL in the folded space (dopri5): the input of the last stage is the step's solution y_1, so Z of that stage IS
  // y_1 @ w1cat^T = Z_0 of the next step.  With z_next set the graph-resident stage kernel writes it there (z_next_valid
  // says it did); with z0_ready set forward_stages takes z0 as given instead of contracting the D-wide state again.
  float* z_next = nullptr;
  bool z0_ready = false, z_next_valid = false;
  float* Vbuf = nullptr;                                   // [N, 2H] V_s of the kernel-per-op path
  float *cat1[kMaxStages] = {}, *cat2[kMaxStages] = {};   // stage slots of the current step
  uint32_t* mask[kMaxStages] = {};                         // [N, 4] ReLU sign bits per stage (chain kernels), same slots
  uint32_t* maskws = nullptr;                              // workspace copy (when nothing is saved)
  // backward
  float *G3 = nullptr, *U = nullptr, *GZ = nullptr;        // [N, 2H]
  float* gzs[kMaxStages] = {};                             // dL/dZ_s  [N, 2H]
  float* Us[kMaxStages] = {};                              // U_s of every stage (backward chain kernel), one contiguous stack
  float* gv2s[kMaxStages] = {};                            // g_v2 of every stage [N, H] (backward chain kernel), contiguous
  float *gcur = nullptr, *gnext = nullptr;                 // [N, D] cotangent ping-pong
  float *R = nullptr, *g1 = nullptr, *cs = nullptr, *partials = nullptr;

  void carve(Arena& a, const Sage3Ctx& c, int S, bool backward);
  static size_t save_floats_per_step(const Sage3Ctx& c, int S);
  void bind_slots(Sage3Ctx& c, float* save, int j);
  int prepare(Sage3Ctx& c, cudaStream_t s);
  // fills cat1 / cat2 of every stage from y; with Cout also C = dt sum_s c_sol[s] cat2_s
  // Cout2 / coef2: a second combination  dt sum_s coef2[s] cat2_s  (dopri5: the error-estimate weights c_err)
  int forward_stages(Sage3Ctx& c, const Tableau& tb, const float* y, float dt, cudaStream_t s, float* Cout = nullptr,
                     float* Cout2 = nullptr, const double* coef2 = nullptr);
  int combine_solution(Sage3Ctx& c, const Tableau& tb, float dt, float* out, cudaStream_t s);
};
// sol0_by_caller: sol[0] is not written here (the first step reads y0); the caller fills it (gnode_decoder_fwd_copy)
// Position decoder fused into the folded solve: traj[j + 1] = traj[j] + C_j @ (Wd @ w3cat)^T + (dt_j sum c) (Wd @ b3), the
// decoded y_{j+1} without reading the D-wide state again (traj[0] is the caller's; n_out <= kDecodeLRMaxOut).
constexpr int kDecodeLRMaxOut = 4;
struct DecodeLR {
  const float* Wd; int n_out;
  float* traj;                    // [n_t, N, n_out]
  float* P;                       // scratch [n_out * (2H + 1)]
};
int integrate_fixed_folded(Sage3Ctx& c, FoldWs& f, const Tableau& tb, const float* y0, const float* t, int n_t,
                           float* sol, float* save, cudaStream_t s, bool sol0_by_caller = false, const DecodeLR* dec = nullptr);
// Cotangent of the last time point given in factored form  G = g1 @ Wd  (g1 [N, n_out], Wd [n_out, D]): what the
// position decoder hands back when the loss reaches the solution only through it (scripts/train_gde.py:486-490).
struct LowRankG {
  const float* g1; const float* Wd; int n_out;
  float *WdW3, *X, *partials;     // scratch: [n_out, 2H], [n_out * 2H + n_out], decoder_wgrad_partial_floats(N, 2H, n_out)
};
int integrate_fixed_folded_bwd(Sage3Ctx& c, FoldWs& f, const Tableau& tb, const float* sol, const float* t, int n_t,
                               const float* grad_sol, float* grad_y0, const float* save, cudaStream_t s,
                               const LowRankG* lr = nullptr);
// backward through dopri5 over the accepted steps tau[0 .. n_acc] of the forward pass (fold.cu)
int integrate_dopri5_folded_bwd(Sage3Ctx& c, FoldWs& f, const float* y0, const double* tau, int n_acc, const double* t,
                                int n_t, const float* grad_sol, float* grad_y0, float* ys, cudaStream_t s);
size_t decoder_wgrad_partial_floats(int64_t M, int D, int n_out);
int decoder_wgrad(const float* x, const float* g, int64_t M, int D, int n_out, float* partials, float* total, cudaStream_t s);
int current_fold();
int current_dopri5_fsal();
// graph-resident forward chain of the folded stages (chain_fwd.cu)
size_t chain_image_floats(int n, int k);
int chain_pack_image(const float* W, int n, int k, int64_t ld, float* img, cudaStream_t s);
bool chain_shape_ok(int H);
bool chain_fwd_supported(const Sage3Ctx& c);
int chain_fwd(Sage3Ctx& c, FoldWs& f, const Tableau& tb, float dt, float* Cout, cudaStream_t s, float* Cout2 = nullptr,
              const double* coef2 = nullptr);
bool chain_bwd_supported(const Sage3Ctx& c, const FoldWs& f);
int chain_bwd(Sage3Ctx& c, FoldWs& f, const Tableau& tb, float dt, bool* has_u, cudaStream_t s);

int check_graph(const gnode_graph* g, const char* who);
int check_params(const gnode_sage3_params* p, const char* who);

// ---- generic drivers (integrate.cu) ----
// Optional per-step hook of integrate_fixed: called before step j; may re-point the field's intermediate slots
// and returns the stage-input buffers xs[1..S-1] to use for this step (stage st then evaluates into slot st).
struct StepSaver {
  virtual ~StepSaver() {}
  virtual float* const* begin_step(int j) = 0;
};
int integrate_fixed(Field& f, int method, const float* y0, const float* t, int n_t, float* sol,
                    float* const* kbuf /* S-1 buffers */, float* xs, cudaStream_t s, StepSaver* saver = nullptr,
                    bool sol0_by_caller = false);

// Backprop through explicit RK steps of a generic field in direct form (integrate.cu).  A cotangent source is a tensor
// G together with the stage weights of the output it belongs to:  y_out = y + dt sum_s w[s] k_s  (c_sol for the step's
// own result, dense-output weights for a dopri5 output inside the step).
struct RkBwdSrc { const float* G; double w[kMaxStages]; };
struct RkBwdBufs { float* kbuf[kMaxStages]; float* xs[kMaxStages]; float* gk; };
// gout = sum_q G_q + sum_s J_s^T g_k[s] (+ extra); gout may alias a source.  Stage s is evaluated into slot s.
int rk_step_bwd(Field& f, const Tableau& tb, const float* y, float dt, const RkBwdSrc* src, int n_src,
                const RkBwdBufs& b, const float* const* xs_saved, const float* extra, float* gout, cudaStream_t s);
// fixed grid: grad_sol [n_t, rows, dim]; gcur scratch [rows, dim]; grad_y0 may be null
int integrate_fixed_bwd_generic(Field& f, const Tableau& tb, const float* sol, const float* t, int n_t, const float* grad_sol,
                                float* grad_y0, const RkBwdBufs& b, float* gcur, cudaStream_t s);
// dopri5 over the accepted steps tau[0 .. n_acc]; ys scratch [n_acc, rows, dim]; gping / gpong scratch [rows, dim]
int integrate_dopri5_bwd_generic(Field& f, const float* y0, const double* tau, int n_acc, const double* t, int n_t,
                                 const float* grad_sol, float* grad_y0, const RkBwdBufs& b, float* ys, float* gping,
                                 float* gpong, cudaStream_t s);
void dopri5_dense_weights(const Tableau& tb, double x, double* w);

struct Dopri5Bufs {
  float* k[7];
  float* ya; float* yb; float* xs;
  double* partials; double* dsum;   // device
};
int integrate_dopri5(Field& f, const float* y0, const double* t, int n_t, double rtol, double atol, float* sol,
                     gnode_dopri5_stats* stats, const gnode_dopri5_trace* trace, gnode_allreduce_fn allreduce,
                     void* allreduce_user, int64_t max_num_steps, const Dopri5Bufs& b, cudaStream_t s);

}  // namespace gnode
