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
  float *a1 = nullptr, *a2 = nullptr;
  int64_t rows() const override { return M; }
  int dim() const override { return H; }
  void carve(Arena& a) {
    a1 = a.take<float>((size_t)M * h);
    a2 = a.take<float>((size_t)M * h);
  }
  int eval(const float* x, float* out, const float* base, float scale, int, cudaStream_t s) override {
    GemmNT q{};
    q.A = x; q.lda = H; q.B = p.w0; q.ldb = H; q.C = a1; q.ldc = h; q.M = M; q.N = h; q.K = H; q.bias = p.b0; q.relu = 2;
    GN_TRY(gemm_nt(q, s));
    GemmNT r{};
    r.A = a1; r.lda = h; r.B = p.w1; r.ldb = h; r.C = a2; r.ldc = h; r.M = M; r.N = h; r.K = h; r.bias = p.b1; r.relu = 2;
    GN_TRY(gemm_nt(r, s));
    GemmNT u{};
    u.A = a2; u.lda = h; u.B = p.w2; u.ldb = h; u.C = out; u.ldc = H; u.M = M; u.N = H; u.K = h; u.bias = p.b2;
    u.base = base; u.ldbase = H; u.scale = scale;
    GN_TRY(gemm_nt(u, s));
    return GNODE_OK;
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
