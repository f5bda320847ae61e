// Node-wise MLP vector field of the secondary GNODE variants:
//   ODEFunction.net = Linear(H,h) -> tanh -> Linear(h,h) -> tanh -> Linear(h,H)
//   (scripts/gnode.py:160-174, scripts/run_gnode.py:153-167), integrated by the same solver drivers
//   (odeint call sites scripts/gnode.py:136-137 -- default dopri5 -- and scripts/run_gnode.py:134-135).
#include <cstring>

#include "field.cuh"

namespace gnode {
namespace {

struct MlpCtx : Field {
  int64_t M = 0;
  int H = 0, h = 0;
  gnode_mlp_params p{};
  int n_slots = 1;
  float* a1[kMaxStages] = {};   // tanh outputs per slot [M, h]
  float* a2[kMaxStages] = {};
  // backward: transposed weights (NT form of the data gradients), scratch, parameter gradients (accumulated)
  float *w0T = nullptr, *w1T = nullptr, *w2T = nullptr;     // [H, h], [h, h], [h, H]
  float *g2 = nullptr, *g1 = nullptr, *partials = nullptr;
  float *dw0 = nullptr, *db0 = nullptr, *dw1 = nullptr, *db1 = nullptr, *dw2 = nullptr, *db2 = nullptr;
  int64_t rows() const override { return M; }
  int dim() const override { return H; }
  void carve(Arena& a, int slots = 1, bool backward = false) {
    n_slots = slots;
    for (int i = 0; i < slots; ++i) { a1[i] = a.take<float>((size_t)M * h); a2[i] = a.take<float>((size_t)M * h); }
    if (backward) {
      w0T = a.take<float>((size_t)H * h); w1T = a.take<float>((size_t)h * h); w2T = a.take<float>((size_t)h * H);
      g2 = a.take<float>((size_t)M * h); g1 = a.take<float>((size_t)M * h);
      size_t pf = gemm_tn_workspace_floats(H, h, M);
      const size_t c2 = gemm_tn_workspace_floats(h, h, M), c3 = gemm_tn_workspace_floats(h, H, M);
      if (c2 > pf) pf = c2;
      if (c3 > pf) pf = c3;
      partials = a.take<float>(pf);
      dw0 = a.take<float>((size_t)h * H); db0 = a.take<float>(h);
      dw1 = a.take<float>((size_t)h * h); db1 = a.take<float>(h);
      dw2 = a.take<float>((size_t)H * h); db2 = a.take<float>(H);
    }
  }
  int prepare_backward(cudaStream_t s) {
    PackSegHost sg[3] = {{w0T, p.w0, h, H, (int64_t)H, (int64_t)h, 1, 0},      // w0T[c, r] = w0[r, c]
                         {w1T, p.w1, h, h, (int64_t)h, (int64_t)h, 1, 0},
                         {w2T, p.w2, H, h, (int64_t)h, (int64_t)H, 1, 0}};
    GN_TRY(pack_segments(sg, 3, s));
    GN_CUDA(cudaMemsetAsync(dw0, 0, sizeof(float) * h * H, s)); GN_CUDA(cudaMemsetAsync(db0, 0, sizeof(float) * h, s));
    GN_CUDA(cudaMemsetAsync(dw1, 0, sizeof(float) * h * h, s)); GN_CUDA(cudaMemsetAsync(db1, 0, sizeof(float) * h, s));
    GN_CUDA(cudaMemsetAsync(dw2, 0, sizeof(float) * H * h, s)); GN_CUDA(cudaMemsetAsync(db2, 0, sizeof(float) * H, s));
    return GNODE_OK;
  }
  int eval(const float* x, float* out, const float* base, float scale, int slot, cudaStream_t s) override {
    float* A1 = a1[slot < n_slots ? slot : 0];
    float* A2 = a2[slot < n_slots ? slot : 0];
    GemmNT q{};
    q.A = x; q.lda = H; q.B = p.w0; q.ldb = H; q.C = A1; q.ldc = h; q.M = M; q.N = h; q.K = H; q.bias = p.b0; q.relu = 2;
    GN_TRY(gemm_nt(q, s));
    GemmNT r{};
    r.A = A1; r.lda = h; r.B = p.w1; r.ldb = h; r.C = A2; r.ldc = h; r.M = M; r.N = h; r.K = h; r.bias = p.b1; r.relu = 2;
    GN_TRY(gemm_nt(r, s));
    GemmNT u{};
    u.A = A2; u.lda = h; u.B = p.w2; u.ldb = h; u.C = out; u.ldc = H; u.M = M; u.N = H; u.K = h; u.bias = p.b2;
    u.base = base; u.ldbase = H; u.scale = scale;
    GN_TRY(gemm_nt(u, s));
    return GNODE_OK;
  }
  // gx = J^T gk;  dW*, db* += this evaluation's parameter gradients
  int vjp(const float* x, int slot, const float* gk, float* gx, cudaStream_t s) override {
    const float* A1 = a1[slot < n_slots ? slot : 0];
    const float* A2 = a2[slot < n_slots ? slot : 0];
    {  // dw2 += gk^T a2, db2 += colsum(gk)
      GemmTN q{};
      q.A = gk; q.lda = H; q.P = H; q.B = A2; q.ldb = h; q.Q = h; q.Nrows = M; q.C = dw2; q.ldc = h; q.colsumA = db2;
      GN_TRY(gemm_tn(q, partials, s));
    }
    {  // g2 = (gk @ w2) * (1 - a2^2)
      GemmNT q{};
      q.A = gk; q.lda = H; q.B = w2T; q.ldb = H; q.C = g2; q.ldc = h; q.M = M; q.N = h; q.K = H;
      GN_TRY(gemm_nt(q, s));
      GN_TRY(tanh_mask(g2, A2, g2, M * h, s));
    }
    {  // dw1 += g2^T a1, db1 += colsum(g2)
      GemmTN q{};
      q.A = g2; q.lda = h; q.P = h; q.B = A1; q.ldb = h; q.Q = h; q.Nrows = M; q.C = dw1; q.ldc = h; q.colsumA = db1;
      GN_TRY(gemm_tn(q, partials, s));
    }
    {  // g1 = (g2 @ w1) * (1 - a1^2)
      GemmNT q{};
      q.A = g2; q.lda = h; q.B = w1T; q.ldb = h; q.C = g1; q.ldc = h; q.M = M; q.N = h; q.K = h;
      GN_TRY(gemm_nt(q, s));
      GN_TRY(tanh_mask(g1, A1, g1, M * h, s));
    }
    {  // dw0 += g1^T x, db0 += colsum(g1)
      GemmTN q{};
      q.A = g1; q.lda = h; q.P = h; q.B = x; q.ldb = H; q.Q = H; q.Nrows = M; q.C = dw0; q.ldc = H; q.colsumA = db0;
      GN_TRY(gemm_tn(q, partials, s));
    }
    if (gx) {  // gx = g1 @ w0
      GemmNT q{};
      q.A = g1; q.lda = h; q.B = w0T; q.ldb = h; q.C = gx; q.ldc = H; q.M = M; q.N = H; q.K = h;
      GN_TRY(gemm_nt(q, s));
    }
    return GNODE_OK;
  }
  // grads += accumulated parameter gradients (any pointer may be null)
  int unpack(const gnode_mlp_grads& g, cudaStream_t s) {
    PackSegHost sg[6];
    int n = 0;
    auto add = [&](float* dst, const float* src, int rows, int cols) {
      if (dst) sg[n++] = PackSegHost{dst, src, rows, cols, (int64_t)cols, (int64_t)cols, 0, 1};
    };
    add(g.w0, dw0, h, H); add(g.b0, db0, 1, h); add(g.w1, dw1, h, h); add(g.b1, db1, 1, h); add(g.w2, dw2, H, h); add(g.b2, db2, 1, H);
    return n ? pack_segments(sg, n, s) : GNODE_OK;
  }
};

int check_mlp(const gnode_mlp_params* p, int64_t m, const char* who) {
  GN_ARG(p != nullptr, "%s: params is null", who);
  GN_ARG(p->dim > 0 && p->hidden_dim > 0 && m > 0, "%s: dim / hidden_dim / rows must be positive", who);
  GN_ARG(p->w0 && p->b0 && p->w1 && p->b1 && p->w2 && p->b2, "%s: null parameter pointer", who);
  return GNODE_OK;
}

}  // namespace
}  // namespace gnode

using namespace gnode;

extern "C" size_t gnode_mlp_ode_workspace_bytes(int64_t m, int32_t dim, int32_t hidden_dim, int32_t method) {
  Arena a(nullptr, 0);
  MlpCtx c;
  c.M = m; c.H = dim; c.h = hidden_dim;
  c.carve(a);
  const size_t n = (size_t)m * dim;
  const int nbuf = (method == GNODE_DOPRI5) ? 10 : kMaxStages;
  for (int i = 0; i < nbuf; ++i) a.take<float>(n);
  a.take<double>((size_t)norm_blocks((int64_t)n));
  a.take<double>(2);
  return a.off;
}

// ---- backward ----------------------------------------------------------------------------------
namespace {
void carve_mlp_bwd(Arena& a, MlpCtx& c, int S, RkBwdBufs& b, float** gcur, float** gpong, float** ys, int n_states) {
  c.carve(a, S, true);
  const size_t n = (size_t)c.M * c.H;
  b = RkBwdBufs{};
  for (int i = 0; i < S; ++i) b.kbuf[i] = a.take<float>(n);
  for (int i = 1; i < S; ++i) b.xs[i] = a.take<float>(n);
  b.gk = a.take<float>(n);
  *gcur = a.take<float>(n);
  if (gpong) *gpong = a.take<float>(n);
  if (ys) *ys = a.take<float>(n * (size_t)(n_states > 0 ? n_states : 1));
}
}  // namespace

extern "C" size_t gnode_mlp_bwd_workspace_bytes(int64_t m, int32_t dim, int32_t hidden_dim, int32_t method, int32_t n_accepted) {
  Arena a(nullptr, 0);
  MlpCtx c;
  c.M = m; c.H = dim; c.h = hidden_dim;
  const Tableau* tb = tableau_for(method);
  RkBwdBufs b;
  float *g0, *g1, *ys;
  carve_mlp_bwd(a, c, tb ? tb->S : 1, b, &g0, &g1, &ys, n_accepted);
  return a.off;
}

// grad_x = J^T grad_out of ONE evaluation of the field (ODEFunction.forward under autograd); parameter gradients
// accumulated (+=) into grads.
extern "C" int gnode_mlp_rhs_bwd(const gnode_mlp_params* p, const float* x, const float* grad_out, int64_t m, float* grad_x,
                                 const gnode_mlp_grads* grads, void* workspace, size_t workspace_bytes,
                                 gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_TRY(check_mlp(p, m, "gnode_mlp_rhs_bwd"));
  GN_ARG(x && grad_out, "gnode_mlp_rhs_bwd: null pointer");
  MlpCtx c;
  c.M = m; c.H = p->dim; c.h = p->hidden_dim; c.p = *p;
  Arena a(workspace, workspace_bytes);
  RkBwdBufs b;
  float *g0, *g1, *ys;
  carve_mlp_bwd(a, c, 1, b, &g0, &g1, &ys, 1);
  GN_ARENA_OK(a, "gnode_mlp_rhs_bwd");
  GN_TRY(c.prepare_backward(s));
  GN_TRY(c.eval(x, b.kbuf[0], nullptr, 1.f, 0, s));        // recompute the activations
  GN_TRY(c.vjp(x, 0, grad_out, grad_x, s));
  if (grads) GN_TRY(c.unpack(*grads, s));
  return GNODE_OK;
}

extern "C" int gnode_mlp_integrate_fixed_bwd(const gnode_mlp_params* p, int32_t method, const float* sol, int64_t m,
                                             const float* t, int32_t n_t, const float* grad_sol, float* grad_y0,
                                             const gnode_mlp_grads* grads, void* workspace, size_t workspace_bytes,
                                             gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_TRY(check_mlp(p, m, "gnode_mlp_integrate_fixed_bwd"));
  const Tableau* tb = tableau_for(method);
  GN_ARG(tb && method != GNODE_DOPRI5, "gnode_mlp_integrate_fixed_bwd: method %d is not a fixed-grid solver", method);
  GN_ARG(sol && t && grad_sol && n_t >= 1, "gnode_mlp_integrate_fixed_bwd: null pointer or empty time grid");
  MlpCtx c;
  c.M = m; c.H = p->dim; c.h = p->hidden_dim; c.p = *p;
  Arena a(workspace, workspace_bytes);
  RkBwdBufs b;
  float *gcur, *g1, *ys;
  carve_mlp_bwd(a, c, tb->S, b, &gcur, &g1, &ys, 1);
  GN_ARENA_OK(a, "gnode_mlp_integrate_fixed_bwd");
  GN_TRY(c.prepare_backward(s));
  GN_TRY(integrate_fixed_bwd_generic(c, *tb, sol, t, n_t, grad_sol, grad_y0, b, gcur, s));
  if (grads) GN_TRY(c.unpack(*grads, s));
  return GNODE_OK;
}

extern "C" int gnode_mlp_integrate_dopri5_bwd(const gnode_mlp_params* p, const float* y0, int64_t m, const double* tau,
                                              int32_t n_accepted, const double* t, int32_t n_t, const float* grad_sol,
                                              float* grad_y0, const gnode_mlp_grads* grads, void* workspace,
                                              size_t workspace_bytes, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_TRY(check_mlp(p, m, "gnode_mlp_integrate_dopri5_bwd"));
  GN_ARG(y0 && tau && t && grad_sol && n_t >= 1 && n_accepted >= 0, "gnode_mlp_integrate_dopri5_bwd: bad argument");
  for (int k = 0; k < n_accepted; ++k)
    GN_ARG(tau[k + 1] > tau[k], "gnode_mlp_integrate_dopri5_bwd: accepted step times must be strictly increasing");
  GN_ARG(n_t == 1 || (n_accepted >= 1 && tau[0] == t[0] && tau[n_accepted] >= t[n_t - 1]),
         "gnode_mlp_integrate_dopri5_bwd: the accepted steps do not cover the time grid");
  MlpCtx c;
  c.M = m; c.H = p->dim; c.h = p->hidden_dim; c.p = *p;
  Arena a(workspace, workspace_bytes);
  RkBwdBufs b;
  float *gping, *gpong, *ys;
  carve_mlp_bwd(a, c, 7, b, &gping, &gpong, &ys, n_accepted);
  GN_ARENA_OK(a, "gnode_mlp_integrate_dopri5_bwd");
  GN_TRY(c.prepare_backward(s));
  GN_TRY(integrate_dopri5_bwd_generic(c, y0, tau, n_accepted, t, n_t, grad_sol, grad_y0, b, ys, gping, gpong, s));
  if (grads) GN_TRY(c.unpack(*grads, s));
  return GNODE_OK;
}

extern "C" int gnode_mlp_rhs_fwd(const gnode_mlp_params* p, const float* x, int64_t m, float* dxdt, void* workspace,
                                 size_t workspace_bytes, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_TRY(check_mlp(p, m, "gnode_mlp_rhs_fwd"));
  GN_ARG(x && dxdt, "gnode_mlp_rhs_fwd: null pointer");
  MlpCtx c;
  c.M = m; c.H = p->dim; c.h = p->hidden_dim; c.p = *p;
  Arena a(workspace, workspace_bytes);
  c.carve(a);
  GN_ARENA_OK(a, "gnode_mlp_rhs_fwd");
  return c.eval(x, dxdt, nullptr, 1.f, 0, s);
}

extern "C" int gnode_mlp_integrate_fixed(const gnode_mlp_params* p, int32_t method, const float* y0, int64_t m,
                                         const float* t, int32_t n_t, float* sol, void* workspace,
                                         size_t workspace_bytes, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_TRY(check_mlp(p, m, "gnode_mlp_integrate_fixed"));
  const Tableau* tb = tableau_for(method);
  GN_ARG(tb && method != GNODE_DOPRI5, "gnode_mlp_integrate_fixed: method %d is not a fixed-grid solver", method);
  GN_ARG(y0 && t && sol && n_t >= 1, "gnode_mlp_integrate_fixed: null pointer or empty time grid");
  for (int j = 0; j + 1 < n_t; ++j) GN_ARG(t[j + 1] > t[j], "gnode_mlp_integrate_fixed: t must be strictly increasing");
  MlpCtx c;
  c.M = m; c.H = p->dim; c.h = p->hidden_dim; c.p = *p;
  Arena a(workspace, workspace_bytes);
  c.carve(a);
  const size_t n = (size_t)m * p->dim;
  float* kbuf[kMaxStages] = {};
  for (int i = 0; i < tb->S - 1; ++i) kbuf[i] = a.take<float>(n);
  float* xs = a.take<float>(n);
  GN_ARENA_OK(a, "gnode_mlp_integrate_fixed");
  return integrate_fixed(c, method, y0, t, n_t, sol, kbuf, xs, s);
}

extern "C" int gnode_mlp_integrate_dopri5(const gnode_mlp_params* p, const float* y0, int64_t m, const double* t,
                                          int32_t n_t, double rtol, double atol, float* sol, gnode_dopri5_stats* stats,
                                          const gnode_dopri5_trace* trace, int64_t max_num_steps, void* workspace,
                                          size_t workspace_bytes, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_TRY(check_mlp(p, m, "gnode_mlp_integrate_dopri5"));
  GN_ARG(y0 && t && sol && n_t >= 1, "gnode_mlp_integrate_dopri5: null pointer or empty time grid");
  for (int j = 0; j + 1 < n_t; ++j) GN_ARG(t[j + 1] > t[j], "gnode_mlp_integrate_dopri5: t must be strictly increasing");
  MlpCtx c;
  c.M = m; c.H = p->dim; c.h = p->hidden_dim; c.p = *p;
  Arena a(workspace, workspace_bytes);
  c.carve(a);
  const size_t n = (size_t)m * p->dim;
  Dopri5Bufs b{};
  for (int i = 0; i < 7; ++i) b.k[i] = a.take<float>(n);
  b.ya = a.take<float>(n);
  b.yb = a.take<float>(n);
  b.xs = a.take<float>(n);
  b.partials = a.take<double>((size_t)norm_blocks((int64_t)n));
  b.dsum = a.take<double>(2);
  GN_ARENA_OK(a, "gnode_mlp_integrate_dopri5");
  if (max_num_steps <= 0) max_num_steps = (1ll << 31) - 1;
  return integrate_dopri5(c, y0, t, n_t, rtol, atol, sol, stats, trace, nullptr, nullptr, max_num_steps, b, s);
}
