// Shared helpers of libgnode_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <cstdio>

#include "../../include/gnode_b200.h"
#include "prof.cuh"

// Opens a profiling scope (no-op unless gnode_prof_enable(1)) covering the launches that follow in
// the enclosing block; flops / bytes are the algorithmic work of those launches.
#define GN_PROF(stream, flops, bytes, ...)                                                        \
  char prof_label__[64];                                                                          \
  prof_label__[0] = 0;                                                                            \
  if (::gnode::prof_enabled()) std::snprintf(prof_label__, sizeof(prof_label__), __VA_ARGS__);   \
  ::gnode::ProfScope prof_scope__(prof_label__, stream, (double)(flops), (double)(bytes))

namespace gnode {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
bool first_use_on_device(const void* key);   // runtime.cu: per-device one-time setup guard

#define GN_CUDA(expr)                                                                     \
  do {                                                                                    \
    cudaError_t e__ = (expr);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      ::gnode::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
      return GNODE_ERR_CUDA;                                                              \
    }                                                                                     \
  } while (0)

#define GN_LAUNCHED()                   \
  do {                                  \
    ::gnode::count_launch();            \
    GN_CUDA(cudaGetLastError());        \
  } while (0)

#define GN_TRY(expr)              \
  do {                            \
    int rc__ = (expr);            \
    if (rc__ != GNODE_OK) return rc__; \
  } while (0)

#define GN_ARG(cond, ...)                 \
  do {                                    \
    if (!(cond)) {                        \
      ::gnode::set_error(__VA_ARGS__);    \
      return GNODE_ERR_ARG;               \
    }                                     \
  } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Bump allocator over the caller-provided workspace.  With base == nullptr it only measures.
struct Arena {
  char* base;
  size_t cap;
  size_t off = 0;
  bool overflow = false;
  Arena(void* b, size_t c) : base(static_cast<char*>(b)), cap(c) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T));
    size_t at = off;
    off += bytes;
    if (base == nullptr) return nullptr;
    if (off > cap) {
      overflow = true;
      return nullptr;
    }
    return reinterpret_cast<T*>(base + at);
  }
  size_t mark() const { return off; }
  void reset(size_t m) { off = m; }
};

#define GN_ARENA_OK(arena, what)                                                              \
  do {                                                                                        \
    if ((arena).overflow) {                                                                   \
      ::gnode::set_error("%s: workspace too small (need >= %zu bytes, got %zu)", what,        \
                         (arena).off, (arena).cap);                                           \
      return GNODE_ERR_WORKSPACE;                                                             \
    }                                                                                         \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// ---- engine selection (gemm.cu) ----
int current_engine();
int current_fold();
int current_dopri5_fsal();
bool lazy_images_poisoned();   // GNODE_POISON_LAZY=1: lazily packed weight images start as NaN patterns (tests)

// ---- dense contractions (gemm_simt.cu / gemm_tc.cu) ----
// C[m, n] = epi( sum_k A[m, k] * B[n, k] )          (both operands K-contiguous, "NT")
//   epi(v) = post( base_scale * base[m, n] + base2[m, n] + scale * act(v + bias_scale * bias[n]) )
//   with act (field `relu`) = 0 identity, 1 relu, 2 tanh; post = relu when post_relu; bias / base / base2 may be null.
//   lda/ldb/ldc/ldbase/ldbase2 are row strides in elements.
struct GemmNT {
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  float* C; int64_t ldc;
  int64_t M; int N; int K;
  const float* bias = nullptr;
  int relu = 0;
  const float* base = nullptr; int64_t ldbase = 0;
  float scale = 1.0f;
  float bias_scale = 1.0f;
  float base_scale = 1.0f;
  const float* base2 = nullptr; int64_t ldbase2 = 0;
  int post_relu = 0;               // ReLU applied to the final value (after the base terms)
  const float* Bsplit = nullptr;   // optional: B pre-split into tf32 hi/lo planes (presplit_weights) -> tcgen05 engine
  const float* Bchain = nullptr;   // optional: chunked chain-format image of B (gemm_k128_pack) -> K = 128 wide-output engine
  int rows_engine = 0;             // the caller accepts the row-major K = 128 engine (three-term truncated product) when it fits
  // Lazy images: a non-null flag that reads 1 says the image behind Bsplit / Bchain has NOT been packed yet; gemm_nt packs
  // it (from B) only if the engine it picks reads it, and clears the flag.  (A training step packed 27 weight images per
  // step and read 11 of them.)
  int* Bsplit_pending = nullptr;
  int* Bchain_pending = nullptr;
};
int gemm_nt(const GemmNT& g, cudaStream_t s);
// tf32 hi/lo planes of a row-major weight matrix, zero padded to multiples of 16 (gemm_tc.cu)
size_t presplit_floats(int rows, int cols);
int presplit_weights(const float* W, int rows, int cols, int64_t ld, float* planes, cudaStream_t s);
int gemm_tc_status(cudaStream_t s, int* out);
// K = 128, wide unaligned output (gemm_k128.cu): chunked weight image of a row-major [n x 128] matrix
size_t gemm_k128_image_floats(int n);
int gemm_k128_pack(const float* W, int n, int64_t ld, float* img, cudaStream_t s);

// C[p, q] (+)= scale * sum_n A[n, p] * B[n, q]        (reduction over rows, "TN": weight gradients)
// Deterministic: split over row chunks into `partials` (workspace), reduced in fixed order.
struct GemmTN {
  const float* A; int64_t lda; int P;
  const float* B; int64_t ldb; int Q;
  int64_t Nrows;
  float* C; int64_t ldc;     // accumulated: C += scale * result
  float scale = 1.0f;
  // optional fused column sums (bias gradients):  colsumA[p] += colsumA_scale * sum_n A[n, p], same for B
  float* colsumA = nullptr; float colsumA_scale = 1.0f;
  float* colsumB = nullptr; float colsumB_scale = 1.0f;
};
size_t gemm_tn_workspace_floats(int P, int Q, int64_t Nrows);
int gemm_tn(const GemmTN& g, float* partials, cudaStream_t s);

// column sums: out[c] += scale * sum_n X[n, c]  (deterministic two-pass)
size_t colsum_workspace_floats(int C, int64_t Nrows);
int colsum_accum(const float* X, int64_t ldx, int64_t Nrows, int C, float* out, float scale,
                 float* partials, cudaStream_t s);

// out[i] += scale * sum_{z < S} partials[z * count + i], fixed order (gemm_simt.cu)
int reduce_partials_accum(const float* partials, int S, int64_t count, float* out, float scale, cudaStream_t s);

// ---- sparse mean aggregation (agg.cu) ----
// out[i, 0:C] = act( mean_{j in in(i)} in[j, 0:C] + add[i, 0:C] + bias[0:C] )
int agg_mean_fwd(const gnode_graph& g, const float* in, int64_t ld_in, float* out, int64_t ld_out,
                 int C, const float* add, int64_t ld_add, const float* bias, int relu,
                 cudaStream_t s);
// out[j, 0:C] = mask(j, c) * ( sum_{i in out(j)} gin[i, 0:C] / max(deg_in(i), 1) + add[j, 0:C] )
//   mask = (act_out[j, c] > 0) when act_out != null (ReLU backward), else 1.
int agg_mean_bwd(const gnode_graph& g, const float* gin, int64_t ld_gin, float* out, int64_t ld_out,
                 int C, const float* add, int64_t ld_add, const float* act_out, int64_t ld_act,
                 cudaStream_t s);

// ---- elementwise stage combinations (rk.cu) ----
constexpr int kMaxTerms = 8;
struct LinComb {
  float* out;
  const float* base;        // may be null (treated as 0)
  const float* in[kMaxTerms];
  float coef[kMaxTerms];
  int n_terms;
  int64_t n;                // elements
};
// out = base + sum_j coef[j] * in[j]   (terms with coef == 0 are skipped)
int lincomb(const LinComb& lc, cudaStream_t s);
// out[i] = act[i] > 0 ? g[i] : 0      (ReLU backward on a flat array)
int relu_mask(const float* g, const float* act, float* out, int64_t n, cudaStream_t s);
// out[i] = g[i] * (1 - act[i]^2)      (tanh backward on a flat array; act = tanh output)
int tanh_mask(const float* g, const float* act, float* out, int64_t n, cudaStream_t s);

}  // namespace gnode
