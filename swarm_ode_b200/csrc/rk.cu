// Elementwise pieces of the Runge-Kutta integrators (torchdiffeq semantics, restated):
// stage-input / solution combinations, the dopri5 error norm, Hairer's initial-step norms and the
// quartic dense output.  All of them are single streaming passes over [n_nodes * D] floats
// (HBM-bound), 128-bit vectorised when the pointers allow; reductions are two-pass and
// deterministic (double partials per block, fixed-order final sum).
//
// Reference call sites: odeint(...) scripts/train_gde.py:78-85, scripts/gnode.py:136-137.
#include "common.cuh"
#include "rk.cuh"

namespace gnode {
namespace {

constexpr int EW_THREADS = 256;

inline unsigned ew_blocks(int64_t n_vec) {
  int64_t b = ceil_div64(n_vec, EW_THREADS);
  const int64_t cap = (int64_t)kNumSMs * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// sum_j in_j * coef_j with each product and each partial sum rounded (mirrors torch.sum(k * c, -1))
__device__ __forceinline__ float comb_terms(const LinComb& lc, int64_t i) {
  float acc = 0.f;
  bool first = true;
#pragma unroll
  for (int j = 0; j < kMaxTerms; ++j) {
    if (j < lc.n_terms) {
      const float p = __fmul_rn(__ldg(lc.in[j] + i), lc.coef[j]);
      acc = first ? p : __fadd_rn(acc, p);
      first = false;
    }
  }
  return acc;
}

__global__ void __launch_bounds__(EW_THREADS) k_lincomb(const LinComb lc) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < lc.n; i += stride) {
    const float acc = comb_terms(lc, i);
    lc.out[i] = lc.base ? __fadd_rn(__ldg(lc.base + i), acc) : acc;
  }
}

__global__ void __launch_bounds__(EW_THREADS) k_lincomb4(const LinComb lc, int64_t n4) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    bool first = true;
#pragma unroll
    for (int j = 0; j < kMaxTerms; ++j) {
      if (j < lc.n_terms) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(lc.in[j]) + i);
        const float c = lc.coef[j];
        const float4 p = make_float4(__fmul_rn(v.x, c), __fmul_rn(v.y, c), __fmul_rn(v.z, c), __fmul_rn(v.w, c));
        acc = first ? p : make_float4(__fadd_rn(acc.x, p.x), __fadd_rn(acc.y, p.y), __fadd_rn(acc.z, p.z), __fadd_rn(acc.w, p.w));
        first = false;
      }
    }
    if (lc.base) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(lc.base) + i);
      acc = make_float4(__fadd_rn(b.x, acc.x), __fadd_rn(b.y, acc.y), __fadd_rn(b.z, acc.z), __fadd_rn(b.w, acc.w));
    }
    reinterpret_cast<float4*>(lc.out)[i] = acc;
  }
  // tail (n % 4 elements) by the first threads of block 0
  if (blockIdx.x == 0) {
    const int64_t i = n4 * 4 + threadIdx.x;
    if (i < lc.n) {
      const float acc = comb_terms(lc, i);
      lc.out[i] = lc.base ? __fadd_rn(__ldg(lc.base + i), acc) : acc;
    }
  }
}

// ---- block reduction of a double -> partials[blockIdx.x] ----
__device__ __forceinline__ void block_reduce_store(double v, double* partials) {
  __shared__ double ws[EW_THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) ws[w] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < EW_THREADS / 32; ++i) s += ws[i];
    partials[blockIdx.x] = s;
  }
}

// sum_i ( (a_i - b_i) / (atol + rtol * |y_i|) )^2         (b may be null)
// VEC = 4: 128-bit loads over the first n & ~3 elements (16-byte aligned operands), scalar tail; the per-element
// arithmetic (and its roundings) is the scalar one, the double-precision sum is order-insensitive at this level.
__device__ __forceinline__ double sq_scaled(float a, float bsub, float y, float atol, float rtol) {
  const float scale = __fadd_rn(atol, __fmul_rn(fabsf(y), rtol));
  const float q = __fdiv_rn(__fsub_rn(a, bsub), scale);
  return (double)__fmul_rn(q, q);
}
template <int VEC>
__global__ void __launch_bounds__(EW_THREADS) k_scaled_sumsq(const float* __restrict__ a, const float* __restrict__ b,
                                                              const float* __restrict__ y, float atol, float rtol,
                                                              int64_t n, double* __restrict__ partials) {
  double acc = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, tid0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t done = 0;
  if (VEC == 4) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid0; i < n4; i += stride) {
      const float4 va = __ldg(reinterpret_cast<const float4*>(a) + i), vy = __ldg(reinterpret_cast<const float4*>(y) + i);
      const float4 vb = b ? __ldg(reinterpret_cast<const float4*>(b) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      // a - 0 == a exactly, so the b == null case keeps the reference's arithmetic
      acc += sq_scaled(va.x, vb.x, vy.x, atol, rtol); acc += sq_scaled(va.y, vb.y, vy.y, atol, rtol);
      acc += sq_scaled(va.z, vb.z, vy.z, atol, rtol); acc += sq_scaled(va.w, vb.w, vy.w, atol, rtol);
    }
    done = n4 << 2;
  }
  for (int64_t i = done + tid0; i < n; i += stride) acc += sq_scaled(__ldg(a + i), b ? __ldg(b + i) : 0.f, __ldg(y + i), atol, rtol);
  block_reduce_store(acc, partials);
}

// dopri5 error ratio:  err_i = sum_j k_j,i * ce_j ;  tol_i = atol + rtol * max(|y0_i|, |y1_i|)
__device__ __forceinline__ double sq_err(float err, float y0, float y1, float atol, float rtol) {
  const float tol = __fadd_rn(atol, __fmul_rn(rtol, fmaxf(fabsf(y0), fabsf(y1))));
  const float q = __fdiv_rn(err, tol);
  return (double)__fmul_rn(q, q);
}
template <int VEC>
__global__ void __launch_bounds__(EW_THREADS) k_error_sumsq(const LinComb lc /* in/coef = k_j, dt*c_err_j; base/out unused */,
                                                             const float* __restrict__ y0, const float* __restrict__ y1,
                                                             float atol, float rtol, double* __restrict__ partials) {
  double acc = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, tid0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t done = 0;
  if (VEC == 4) {
    const int64_t n4 = lc.n >> 2;
    for (int64_t i = tid0; i < n4; i += stride) {
      float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
      bool first = true;
#pragma unroll
      for (int j = 0; j < kMaxTerms; ++j) {
        if (j < lc.n_terms) {       // same products and partial sums as comb_terms, four elements at a time
          const float4 k = __ldg(reinterpret_cast<const float4*>(lc.in[j]) + i);
          const float c = lc.coef[j];
          const float4 p = make_float4(__fmul_rn(k.x, c), __fmul_rn(k.y, c), __fmul_rn(k.z, c), __fmul_rn(k.w, c));
          e = first ? p : make_float4(__fadd_rn(e.x, p.x), __fadd_rn(e.y, p.y), __fadd_rn(e.z, p.z), __fadd_rn(e.w, p.w));
          first = false;
        }
      }
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(y0) + i), a1 = __ldg(reinterpret_cast<const float4*>(y1) + i);
      acc += sq_err(e.x, a0.x, a1.x, atol, rtol); acc += sq_err(e.y, a0.y, a1.y, atol, rtol);
      acc += sq_err(e.z, a0.z, a1.z, atol, rtol); acc += sq_err(e.w, a0.w, a1.w, atol, rtol);
    }
    done = n4 << 2;
  }
  for (int64_t i = done + tid0; i < lc.n; i += stride) acc += sq_err(comb_terms(lc, i), __ldg(y0 + i), __ldg(y1 + i), atol, rtol);
  block_reduce_store(acc, partials);
}

__global__ void k_sum_partials(const double* __restrict__ partials, int n, double* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += partials[i];
    *out = s;
  }
}

// quartic dense output of one accepted dopri5 step evaluated at x in [0,1]  (torchdiffeq _interp_fit /
// _interp_evaluate, operation order kept)
__global__ void __launch_bounds__(EW_THREADS) k_dopri_interp(const LinComb lc /* in = k_0..k_6, coef = dt*mid_j */,
                                                              const float* __restrict__ y0, const float* __restrict__ y1,
                                                              float dt, float x, float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < lc.n; i += stride) {
    const float a0 = __ldg(y0 + i), a1 = __ldg(y1 + i);
    const float f0 = __ldg(lc.in[0] + i), f1 = __ldg(lc.in[6] + i);
    const float ymid = __fadd_rn(a0, comb_terms(lc, i));
    // a = 2*dt*(f1 - f0) - 8*(y1 + y0) + 16*y_mid
    const float ca = __fadd_rn(__fsub_rn(__fmul_rn(__fmul_rn(2.f, dt), __fsub_rn(f1, f0)), __fmul_rn(8.f, __fadd_rn(a1, a0))),
                               __fmul_rn(16.f, ymid));
    // b = dt*(5*f0 - 3*f1) + 18*y0 + 14*y1 - 32*y_mid
    const float cb = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(dt, __fsub_rn(__fmul_rn(5.f, f0), __fmul_rn(3.f, f1))), __fmul_rn(18.f, a0)),
                                         __fmul_rn(14.f, a1)),
                               __fmul_rn(32.f, ymid));
    // c = dt*(f1 - 4*f0) - 11*y0 - 5*y1 + 16*y_mid
    const float cc = __fadd_rn(__fsub_rn(__fsub_rn(__fmul_rn(dt, __fsub_rn(f1, __fmul_rn(4.f, f0))), __fmul_rn(11.f, a0)),
                                         __fmul_rn(5.f, a1)),
                               __fmul_rn(16.f, ymid));
    const float cd = __fmul_rn(dt, f0);
    float total = __fadd_rn(a0, __fmul_rn(x, cd));
    float xp = x;
    xp = __fmul_rn(xp, x); total = __fadd_rn(total, __fmul_rn(xp, cc));
    xp = __fmul_rn(xp, x); total = __fadd_rn(total, __fmul_rn(xp, cb));
    xp = __fmul_rn(xp, x); total = __fadd_rn(total, __fmul_rn(xp, ca));
    out[i] = total;
  }
}

__global__ void __launch_bounds__(EW_THREADS) k_act_mask(const float* __restrict__ g, const float* __restrict__ act,
                                                          float* __restrict__ out, int64_t n, int mode) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    const float a = __ldg(act + i), gv = __ldg(g + i);
    out[i] = (mode == 1) ? (a > 0.f ? gv : 0.f) : gv * (1.f - a * a);
  }
}

struct PackSeg {
  float* dst; const float* src;
  int rows, cols;          // logical shape of the source block
  int64_t ld_src, ld_dst; int transpose;  // dst[r*ld_dst + c] = src[r*ld_src + c]  or  dst[c*ld_dst + r] = ...
  int accumulate;          // dst += src instead of =
};
constexpr int kMaxSegs = 16;
struct PackArgs { PackSeg seg[kMaxSegs]; int n; };

__global__ void k_pack(const PackArgs a) {
  for (int sgi = blockIdx.y; sgi < a.n; sgi += gridDim.y) {
    const PackSeg sg = a.seg[sgi];
    const int64_t total = (int64_t)sg.rows * sg.cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int r = (int)(i / sg.cols), c = (int)(i % sg.cols);
      float* d = sg.transpose ? sg.dst + (int64_t)c * sg.ld_dst + r : sg.dst + (int64_t)r * sg.ld_dst + c;
      const float v = sg.src[(int64_t)r * sg.ld_src + c];
      if (sg.accumulate) *d += v; else *d = v;
    }
  }
}

}  // namespace

int lincomb(const LinComb& lc_in, cudaStream_t s) {
  if (lc_in.n == 0) return GNODE_OK;
  // drop zero coefficients (keeps the summation order of the remaining terms)
  LinComb lc = lc_in;
  int m = 0;
  for (int j = 0; j < lc_in.n_terms; ++j)
    if (lc_in.coef[j] != 0.f) { lc.in[m] = lc_in.in[j]; lc.coef[m] = lc_in.coef[j]; ++m; }
  lc.n_terms = m;
  for (int j = m; j < kMaxTerms; ++j) { lc.in[j] = nullptr; lc.coef[j] = 0.f; }
  if (m == 0 && lc.base == nullptr) { GN_CUDA(cudaMemsetAsync(lc.out, 0, sizeof(float) * lc.n, s)); return GNODE_OK; }
  if (m == 0) {
    if (lc.out != lc.base) GN_CUDA(cudaMemcpyAsync(lc.out, lc.base, sizeof(float) * lc.n, cudaMemcpyDeviceToDevice, s));
    return GNODE_OK;
  }
  GN_PROF(s, 2.0 * m * lc.n, 4.0 * (double)lc.n * (m + 1 + (lc.base ? 1 : 0)), "lincomb terms=%d", m);
  bool v4 = aligned16(lc.out) && (!lc.base || aligned16(lc.base));
  for (int j = 0; j < m; ++j) v4 = v4 && aligned16(lc.in[j]);
  if (v4 && lc.n >= 4) {
    const int64_t n4 = lc.n / 4;
    k_lincomb4<<<ew_blocks(n4), EW_THREADS, 0, s>>>(lc, n4);
  } else {
    k_lincomb<<<ew_blocks(lc.n), EW_THREADS, 0, s>>>(lc);
  }
  GN_LAUNCHED();
  return GNODE_OK;
}

int relu_mask(const float* g, const float* act, float* out, int64_t n, cudaStream_t s) {
  if (n == 0) return GNODE_OK;
  GN_PROF(s, (double)n, 12.0 * (double)n, "act_mask");
  k_act_mask<<<ew_blocks(n), EW_THREADS, 0, s>>>(g, act, out, n, 1);
  GN_LAUNCHED();
  return GNODE_OK;
}

int tanh_mask(const float* g, const float* act, float* out, int64_t n, cudaStream_t s) {
  if (n == 0) return GNODE_OK;
  k_act_mask<<<ew_blocks(n), EW_THREADS, 0, s>>>(g, act, out, n, 2);
  GN_LAUNCHED();
  return GNODE_OK;
}

int norm_blocks(int64_t n) { return (int)ew_blocks(ceil_div64(n, 4)); }

int scaled_sumsq(const float* a, const float* b, const float* y, float atol, float rtol, int64_t n,
                 double* partials, double* out, cudaStream_t s) {
  GN_PROF(s, 6.0 * n, 4.0 * (double)n * (b ? 3 : 2), "scaled_sumsq");
  const int nb = norm_blocks(n);
  if (aligned16(a) && aligned16(y) && (!b || aligned16(b))) k_scaled_sumsq<4><<<nb, EW_THREADS, 0, s>>>(a, b, y, atol, rtol, n, partials);
  else k_scaled_sumsq<1><<<nb, EW_THREADS, 0, s>>>(a, b, y, atol, rtol, n, partials);
  GN_LAUNCHED();
  k_sum_partials<<<1, 32, 0, s>>>(partials, nb, out);
  GN_LAUNCHED();
  return GNODE_OK;
}

int error_sumsq(const LinComb& lc, const float* y0, const float* y1, float atol, float rtol,
                double* partials, double* out, cudaStream_t s) {
  GN_PROF(s, 20.0 * lc.n, 4.0 * (double)lc.n * (lc.n_terms + 2), "dopri5_error_norm");
  const int nb = norm_blocks(lc.n);
  bool al = aligned16(y0) && aligned16(y1);
  for (int j = 0; j < lc.n_terms; ++j) al = al && aligned16(lc.in[j]);
  if (al) k_error_sumsq<4><<<nb, EW_THREADS, 0, s>>>(lc, y0, y1, atol, rtol, partials);
  else k_error_sumsq<1><<<nb, EW_THREADS, 0, s>>>(lc, y0, y1, atol, rtol, partials);
  GN_LAUNCHED();
  k_sum_partials<<<1, 32, 0, s>>>(partials, nb, out);
  GN_LAUNCHED();
  return GNODE_OK;
}

int dopri_interp(const LinComb& lc, const float* y0, const float* y1, float dt, float x, float* out, cudaStream_t s) {
  GN_PROF(s, 40.0 * lc.n, 4.0 * (double)lc.n * (lc.n_terms + 3), "dopri5_dense_output");
  k_dopri_interp<<<ew_blocks(lc.n), EW_THREADS, 0, s>>>(lc, y0, y1, dt, x, out);
  GN_LAUNCHED();
  return GNODE_OK;
}

int pack_segments(const PackSegHost* segs, int n, cudaStream_t s) {
  if (n == 0) return GNODE_OK;
  if (n > kMaxSegs) { set_error("pack_segments: too many segments"); return GNODE_ERR_ARG; }
  PackArgs a;
  a.n = n;
  for (int i = 0; i < n; ++i) {
    a.seg[i].dst = segs[i].dst; a.seg[i].src = segs[i].src; a.seg[i].rows = segs[i].rows; a.seg[i].cols = segs[i].cols;
    a.seg[i].ld_src = segs[i].ld_src; a.seg[i].ld_dst = segs[i].ld_dst; a.seg[i].transpose = segs[i].transpose; a.seg[i].accumulate = segs[i].accumulate;
  }
  GN_PROF(s, 0.0, 0.0, "pack_segments");
  int64_t most = 1;
  for (int i = 0; i < n; ++i) {
    const int64_t tot = (int64_t)segs[i].rows * segs[i].cols;
    if (tot > most) most = tot;
  }
  int64_t gx = ceil_div64(most, 256);
  if (gx > kNumSMs * 8) gx = kNumSMs * 8;
  dim3 grid((unsigned)gx, n);
  k_pack<<<grid, 256, 0, s>>>(a);
  GN_LAUNCHED();
  return GNODE_OK;
}

}  // namespace gnode
