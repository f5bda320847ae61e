// Engine dispatch for the dense contractions: tcgen05 (3xTF32, gemm_tc.cu) or fp32 FFMA (gemm_simt.cu).
#include <cstdlib>

#include "common.cuh"

namespace gnode {
int gemm_nt_simt(const GemmNT& g, cudaStream_t s);
bool gemm_nt_tc_supported(const GemmNT& g);
int gemm_nt_tc(const GemmNT& g, cudaStream_t s);

bool gemm_k128_supported(const GemmNT& g);
bool gemm_k128_rows_supported(const GemmNT& g);
int gemm_k128_rows(const GemmNT& g, cudaStream_t s);

static bool rows_engine_on() {   // GNODE_ROWS_ENGINE=0 falls back to the general engine (A/B runs)
  static const bool on = [] { const char* e = std::getenv("GNODE_ROWS_ENGINE"); return !(e && e[0] == '0'); }();
  return on;
}

int gemm_nt(const GemmNT& g, cudaStream_t s) {
  const int engine = current_engine();
  // The row-major variant (dense D-wide rows moved as contiguous spans) is taken where the caller allows it: the y_1
  // projection of the fixed-grid solvers.
  const bool rows = g.rows_engine && rows_engine_on() && engine != GNODE_ENGINE_SIMT && gemm_k128_rows_supported(g);
  const bool tc = engine != GNODE_ENGINE_SIMT && gemm_nt_tc_supported(g);
  if (rows) {
    if (g.Bchain_pending && *g.Bchain_pending) {
      GN_TRY(gemm_k128_pack(g.B, g.N, g.ldb, const_cast<float*>(g.Bchain), s));
      *g.Bchain_pending = 0;
    }
  } else if (tc) {
    if (g.Bsplit_pending && *g.Bsplit_pending) {
      GN_TRY(presplit_weights(g.B, g.N, g.K, g.ldb, const_cast<float*>(g.Bsplit), s));
      *g.Bsplit_pending = 0;
    }
  }
  GN_PROF(s, 2.0 * g.M * g.N * g.K, 4.0 * ((double)g.M * g.K + (double)g.N * g.K + (double)g.M * g.N * (g.base ? 2 : 1)),
          "gemm_nt[%s] N=%d K=%d", rows ? "k128 rows" : (tc ? "tcgen05" : "ffma"), g.N, g.K);
  if (engine == GNODE_ENGINE_SIMT) return gemm_nt_simt(g, s);
  if (rows) return gemm_k128_rows(g, s);
  if (gemm_nt_tc_supported(g)) return gemm_nt_tc(g, s);
  if (engine == GNODE_ENGINE_TC) {
    set_error("gemm_nt: shape M=%lld N=%d K=%d not supported by the tcgen05 engine", (long long)g.M, g.N, g.K);
    return GNODE_ERR_ARG;
  }
  return gemm_nt_simt(g, s);
}

int gemm_tn_simt(const GemmTN& g, float* partials, cudaStream_t s);
size_t gemm_tn_simt_workspace_floats(int P, int Q, int64_t Nrows);
bool gemm_tn_tc_supported(const GemmTN& g);
int gemm_tn_tc(const GemmTN& g, float* partials, cudaStream_t s);
size_t gemm_tn_tc_workspace_floats(int P, int Q);

size_t gemm_tn_workspace_floats(int P, int Q, int64_t Nrows) {
  const size_t a = gemm_tn_simt_workspace_floats(P, Q, Nrows), b = gemm_tn_tc_workspace_floats(P, Q);
  return a > b ? a : b;
}

// weight gradients: tcgen05 when the engine allows and the operands are dense row slabs, else FFMA
int gemm_tn(const GemmTN& g, float* partials, cudaStream_t s) {
  if (g.P == 0 || g.Q == 0 || g.Nrows == 0) return GNODE_OK;
  const int engine = current_engine();
  const bool tc = engine != GNODE_ENGINE_SIMT && gemm_tn_tc_supported(g);
  GN_PROF(s, 2.0 * g.P * g.Q * g.Nrows, 4.0 * ((double)g.Nrows * (g.P + g.Q) + (double)g.P * g.Q),
          "gemm_tn[%s] P=%d Q=%d", tc ? "tcgen05" : "ffma", g.P, g.Q);
  return tc ? gemm_tn_tc(g, partials, s) : gemm_tn_simt(g, partials, s);
}
}  // namespace gnode

// ------------------------------------------------------------------------------------------------
// C ABI: the dense NT contraction on its own (used by the parity tests to isolate the engines)
// ------------------------------------------------------------------------------------------------
using namespace gnode;

extern "C" size_t gnode_gemm_nt_workspace_bytes(int32_t n, int32_t k) {
  return align_up(presplit_floats(n, k) * sizeof(float));
}

extern "C" int gnode_gemm_nt(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                             int64_t m, int32_t n, int32_t k, const float* bias, int32_t act, const float* base,
                             int64_t ldbase, float scale, void* workspace, size_t workspace_bytes,
                             gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_ARG(A && B && C && m >= 0 && n > 0 && k > 0, "gnode_gemm_nt: bad argument");
  GN_ARG(act >= 0 && act <= 2, "gnode_gemm_nt: act must be 0 (none), 1 (relu) or 2 (tanh)");
  GemmNT q{};
  q.A = A; q.lda = lda; q.B = B; q.ldb = ldb; q.C = C; q.ldc = ldc; q.M = m; q.N = n; q.K = k;
  q.bias = bias; q.relu = act; q.base = base; q.ldbase = ldbase; q.scale = scale;
  if (current_engine() != GNODE_ENGINE_SIMT) {
    Arena a(workspace, workspace_bytes);
    float* planes = a.take<float>(presplit_floats(n, k));
    GN_ARENA_OK(a, "gnode_gemm_nt");
    GN_TRY(presplit_weights(B, n, k, ldb, planes, s));
    q.Bsplit = planes;
  }
  return gemm_nt(q, s);
}

extern "C" size_t gnode_gemm_tn_workspace_bytes(int32_t p, int32_t q, int64_t rows) {
  return align_up(gemm_tn_workspace_floats(p, q, rows) * sizeof(float));
}

extern "C" int gnode_gemm_tn(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                             int64_t rows, int32_t p, int32_t q, float scale, void* workspace,
                             size_t workspace_bytes, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_ARG(A && B && C && rows >= 0 && p > 0 && q > 0, "gnode_gemm_tn: bad argument");
  Arena a(workspace, workspace_bytes);
  float* partials = a.take<float>(gemm_tn_workspace_floats(p, q, rows));
  GN_ARENA_OK(a, "gnode_gemm_tn");
  GemmTN g{};
  g.A = A; g.lda = lda; g.P = p; g.B = B; g.ldb = ldb; g.Q = q; g.Nrows = rows; g.C = C; g.ldc = ldc; g.scale = scale;
  return gemm_tn(g, partials, s);
}

// The K = 128 wide-output engine on its own (gemm_k128.cu; parity tests).  B: row-major [n, 128].
extern "C" size_t gnode_gemm_k128_workspace_bytes(int32_t n) {
  return align_up(gemm_k128_image_floats(n) * sizeof(float)) + align_up(presplit_floats(n, 128) * sizeof(float));
}

extern "C" int gnode_gemm_k128(const float* A, const float* B, float* C, int64_t ldc, int64_t m, int32_t n, const float* bias,
                               float bias_scale, const float* base, int64_t ldbase, float base_scale, const float* base2,
                               int64_t ldbase2, float scale, void* workspace, size_t workspace_bytes, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_ARG(A && B && C && m > 0 && n >= 16, "gnode_gemm_k128: bad argument");
  Arena a(workspace, workspace_bytes);
  float* img = a.take<float>(gemm_k128_image_floats(n));
  GN_ARENA_OK(a, "gnode_gemm_k128");
  GN_TRY(gemm_k128_pack(B, n, 128, img, s));
  GemmNT q{};
  q.A = A; q.lda = 128; q.B = B; q.ldb = 128; q.C = C; q.ldc = ldc; q.M = m; q.N = n; q.K = 128;
  q.bias = bias; q.bias_scale = bias_scale; q.base = base; q.ldbase = ldbase; q.base_scale = base_scale;
  q.base2 = base2; q.ldbase2 = ldbase2; q.scale = scale; q.Bchain = img;
  GN_ARG(gemm_k128_supported(q), "gnode_gemm_k128: operand A must be 16-byte aligned");
  if (rows_engine_on() && gemm_k128_rows_supported(q)) return gemm_k128_rows(q, s);   // dense rows, one base term
  // every other shape (two base terms, strided rows, N > 400): the general tcgen05 engine
  float* planes = a.take<float>(presplit_floats(n, 128));
  GN_ARENA_OK(a, "gnode_gemm_k128");
  GN_TRY(presplit_weights(B, n, 128, 128, planes, s));
  q.Bsplit = planes; q.Bchain = nullptr;
  return gemm_nt(q, s);
}
