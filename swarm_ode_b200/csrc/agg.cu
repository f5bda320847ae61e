// Sparse mean aggregation over the CSR graph: the "edge gather + scatter-mean" of SAGEConv
// (reference call sites scripts/train_gde.py:36,39,43; PyG propagate with aggr='mean').
//
// Forward is a pull over destination-sorted CSR (no atomics, fixed summation order); backward is a
// pull over the source-sorted CSR with the 1/deg factor of the destination applied per edge.
// Rows are read with 128-bit loads when the channel count / strides allow (hidden width 64), a
// group of C/4 lanes per node row, so one warp serves 32/(C/4) rows.
#include "common.cuh"

namespace gnode {
namespace {

template <int VEC> struct Vec;
template <> struct Vec<4> {
  using T = float4;
  static __device__ __forceinline__ T zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  static __device__ __forceinline__ T ld(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  static __device__ __forceinline__ void st(float* p, T v) { *reinterpret_cast<float4*>(p) = v; }
  static __device__ __forceinline__ T add(T a, T b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
  static __device__ __forceinline__ T div(T a, float d) { return make_float4(a.x / d, a.y / d, a.z / d, a.w / d); }
  static __device__ __forceinline__ T scale(T a, float w) { return make_float4(a.x * w, a.y * w, a.z * w, a.w * w); }
  static __device__ __forceinline__ T relu(T a) { return make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f)); }
  static __device__ __forceinline__ T mask(T a, T h) {
    return make_float4(h.x > 0.f ? a.x : 0.f, h.y > 0.f ? a.y : 0.f, h.z > 0.f ? a.z : 0.f, h.w > 0.f ? a.w : 0.f);
  }
};
template <> struct Vec<1> {
  using T = float;
  static __device__ __forceinline__ T zero() { return 0.f; }
  static __device__ __forceinline__ T ld(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ void st(float* p, T v) { *p = v; }
  static __device__ __forceinline__ T add(T a, T b) { return a + b; }
  static __device__ __forceinline__ T div(T a, float d) { return a / d; }
  static __device__ __forceinline__ T scale(T a, float w) { return a * w; }
  static __device__ __forceinline__ T relu(T a) { return fmaxf(a, 0.f); }
  static __device__ __forceinline__ T mask(T a, T h) { return h > 0.f ? a : 0.f; }
};

// The kernels are issue bound, not bandwidth bound (ncu: ~140 M warp instructions for 389 k rows, DRAM at 10-20 %):
// every lane of a row group repeats the row's index arithmetic.  So a lane owns UNR column chunks of its row (fewer
// lanes per row), neighbours are taken two at a time (mean in-degree ~2) and the per-edge 1/deg is one reciprocal.
template <int VEC, int UNR>
__global__ void __launch_bounds__(256) k_agg_fwd(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                  int64_t N, const float* __restrict__ in, int64_t ld_in,
                                                  float* __restrict__ out, int64_t ld_out, int C,
                                                  const float* __restrict__ add, int64_t ld_add,
                                                  const float* __restrict__ bias, int relu, int lpr) {
  using V = Vec<VEC>;
  const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t row = gid / lpr;
  const int gl = (int)(gid % lpr);
  if (row >= N) return;
  const int b = rowptr[row], e = rowptr[row + 1];
  const float inv = 1.0f / (float)((e - b) > 1 ? (e - b) : 1);
  const int cstep = lpr * VEC;                 // columns between the chunks of one lane
  for (int c0 = gl * VEC; c0 < C; c0 += cstep * UNR) {
    typename V::T acc[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) acc[u] = V::zero();
    for (int p = b; p < e; p += 2) {
      const int j0 = col[p];
      const int j1 = (p + 1 < e) ? col[p + 1] : -1;
      const float* r0 = in + (int64_t)j0 * ld_in + c0;
      const float* r1 = in + (int64_t)(j1 >= 0 ? j1 : j0) * ld_in + c0;
      typename V::T v0[UNR], v1[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const bool ok = c0 + u * cstep < C;
        v0[u] = ok ? V::ld(r0 + u * cstep) : V::zero();
        v1[u] = (ok && j1 >= 0) ? V::ld(r1 + u * cstep) : V::zero();
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        acc[u] = V::add(acc[u], v0[u]);
        if (j1 >= 0) acc[u] = V::add(acc[u], v1[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int c = c0 + u * cstep;
      if (c >= C) break;
      typename V::T r = V::scale(acc[u], inv);
      if (add) r = V::add(r, V::ld(add + row * ld_add + c));
      if (bias) r = V::add(r, V::ld(bias + c));
      if (relu) r = V::relu(r);
      V::st(out + row * ld_out + c, r);
    }
  }
}

template <int VEC, int UNR>
__global__ void __launch_bounds__(256) k_agg_bwd(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ t_rowptr,
                                                  const int32_t* __restrict__ t_col, int64_t N,
                                                  const float* __restrict__ gin, int64_t ld_gin,
                                                  float* __restrict__ out, int64_t ld_out, int C,
                                                  const float* __restrict__ add, int64_t ld_add,
                                                  const float* __restrict__ act, int64_t ld_act, int lpr) {
  using V = Vec<VEC>;
  const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t row = gid / lpr;
  const int gl = (int)(gid % lpr);
  if (row >= N) return;
  const int b = t_rowptr[row], e = t_rowptr[row + 1];
  const int cstep = lpr * VEC;
  for (int c0 = gl * VEC; c0 < C; c0 += cstep * UNR) {
    typename V::T acc[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) acc[u] = V::zero();
    for (int p = b; p < e; p += 2) {
      const int i0 = t_col[p];
      const int i1 = (p + 1 < e) ? t_col[p + 1] : -1;
      const int ii1 = i1 >= 0 ? i1 : i0;
      // deg >= 1 since the edge (row -> i) exists
      const float w0 = 1.0f / (float)(rowptr[i0 + 1] - rowptr[i0]);
      const float w1 = 1.0f / (float)(rowptr[ii1 + 1] - rowptr[ii1]);
      const float* r0 = gin + (int64_t)i0 * ld_gin + c0;
      const float* r1 = gin + (int64_t)ii1 * ld_gin + c0;
      typename V::T v0[UNR], v1[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const bool ok = c0 + u * cstep < C;
        v0[u] = ok ? V::ld(r0 + u * cstep) : V::zero();
        v1[u] = (ok && i1 >= 0) ? V::ld(r1 + u * cstep) : V::zero();
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        acc[u] = V::add(acc[u], V::scale(v0[u], w0));
        if (i1 >= 0) acc[u] = V::add(acc[u], V::scale(v1[u], w1));
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int c = c0 + u * cstep;
      if (c >= C) break;
      typename V::T r = acc[u];
      if (add) r = V::add(r, V::ld(add + row * ld_add + c));
      if (act) r = V::mask(r, V::ld(act + row * ld_act + c));
      V::st(out + row * ld_out + c, r);
    }
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

constexpr int AGG_UNR = 2;   // column chunks per lane
inline int lanes_per_row(int C, int vec) {
  int chunks = ((C + vec - 1) / vec + AGG_UNR - 1) / AGG_UNR;
  int l = 1;
  while (l < chunks && l < 32) l <<= 1;
  return l;
}

}  // namespace

int agg_mean_fwd(const gnode_graph& g, const float* in, int64_t ld_in, float* out, int64_t ld_out, int C,
                 const float* add, int64_t ld_add, const float* bias, int relu, cudaStream_t s) {
  if (g.n_nodes == 0 || C == 0) return GNODE_OK;
  GN_PROF(s, (double)g.n_edges * C, 4.0 * ((double)g.n_edges * (C + 1) + (double)g.n_nodes * (C * (add ? 2 : 1) + 1)),
          "agg_mean_fwd C=%d", C);
  const bool vec4 = (C % 4 == 0) && (ld_in % 4 == 0) && (ld_out % 4 == 0) && aligned16(in) && aligned16(out) &&
                    (!add || ((ld_add % 4 == 0) && aligned16(add))) && (!bias || aligned16(bias));
  const int vec = vec4 ? 4 : 1;
  const int lpr = lanes_per_row(C, vec);
  const int64_t threads = g.n_nodes * lpr;
  const unsigned blocks = (unsigned)ceil_div64(threads, 256);
  if (vec4)
    k_agg_fwd<4, AGG_UNR><<<blocks, 256, 0, s>>>(g.rowptr, g.col, g.n_nodes, in, ld_in, out, ld_out, C, add, ld_add, bias, relu, lpr);
  else
    k_agg_fwd<1, AGG_UNR><<<blocks, 256, 0, s>>>(g.rowptr, g.col, g.n_nodes, in, ld_in, out, ld_out, C, add, ld_add, bias, relu, lpr);
  GN_LAUNCHED();
  return GNODE_OK;
}

int agg_mean_bwd(const gnode_graph& g, const float* gin, int64_t ld_gin, float* out, int64_t ld_out, int C,
                 const float* add, int64_t ld_add, const float* act, int64_t ld_act, cudaStream_t s) {
  if (g.n_nodes == 0 || C == 0) return GNODE_OK;
  GN_PROF(s, (double)g.n_edges * C, 4.0 * ((double)g.n_edges * (C + 1) + (double)g.n_nodes * (C * (1 + (add ? 1 : 0) + (act ? 1 : 0)) + 1)),
          "agg_mean_bwd C=%d", C);
  const bool vec4 = (C % 4 == 0) && (ld_gin % 4 == 0) && (ld_out % 4 == 0) && aligned16(gin) && aligned16(out) &&
                    (!add || ((ld_add % 4 == 0) && aligned16(add))) && (!act || ((ld_act % 4 == 0) && aligned16(act)));
  const int vec = vec4 ? 4 : 1;
  const int lpr = lanes_per_row(C, vec);
  const int64_t threads = g.n_nodes * lpr;
  const unsigned blocks = (unsigned)ceil_div64(threads, 256);
  if (vec4)
    k_agg_bwd<4, AGG_UNR><<<blocks, 256, 0, s>>>(g.rowptr, g.t_rowptr, g.t_col, g.n_nodes, gin, ld_gin, out, ld_out, C, add, ld_add, act, ld_act, lpr);
  else
    k_agg_bwd<1, AGG_UNR><<<blocks, 256, 0, s>>>(g.rowptr, g.t_rowptr, g.t_col, g.n_nodes, gin, ld_gin, out, ld_out, C, add, ld_add, act, ld_act, lpr);
  GN_LAUNCHED();
  return GNODE_OK;
}

}  // namespace gnode
