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
  static __device__ __forceinline__ T relu(T a) { return fmaxf(a, 0.f); }
  static __device__ __forceinline__ T mask(T a, T h) { return h > 0.f ? a : 0.f; }
};

template <int VEC>
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
  const float denom = (float)((e - b) > 1 ? (e - b) : 1);
  for (int c = gl * VEC; c < C; c += lpr * VEC) {
    typename V::T acc = V::zero();
    // batches of four neighbours, indices first, then all row loads in flight together (the tail batch is
    // predicated instead of serialised: most rows of the warehouse graphs have 1-3 neighbours)
    for (int p = b; p < e; p += 4) {
      int j[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) j[k] = (p + k < e) ? col[p + k] : -1;
      typename V::T v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = (j[k] >= 0) ? V::ld(in + (int64_t)j[k] * ld_in + c) : V::zero();
#pragma unroll
      for (int k = 0; k < 4; ++k) if (j[k] >= 0) acc = V::add(acc, v[k]);
    }
    acc = V::div(acc, denom);
    if (add) acc = V::add(acc, V::ld(add + row * ld_add + c));
    if (bias) acc = V::add(acc, V::ld(bias + c));
    if (relu) acc = V::relu(acc);
    V::st(out + row * ld_out + c, acc);
  }
}

template <int VEC>
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
  for (int c = gl * VEC; c < C; c += lpr * VEC) {
    typename V::T acc = V::zero();
    for (int p = b; p < e; p += 4) {
      int i[4], r0[4], r1[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) i[k] = (p + k < e) ? t_col[p + k] : -1;
#pragma unroll
      for (int k = 0; k < 4; ++k) { r0[k] = (i[k] >= 0) ? rowptr[i[k]] : 0; r1[k] = (i[k] >= 0) ? rowptr[i[k] + 1] : 1; }
      typename V::T v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = (i[k] >= 0) ? V::ld(gin + (int64_t)i[k] * ld_gin + c) : V::zero();
#pragma unroll
      for (int k = 0; k < 4; ++k)   // deg >= 1 since the edge (row -> i) exists
        if (i[k] >= 0) acc = V::add(acc, V::div(v[k], (float)(r1[k] - r0[k])));
    }
    if (add) acc = V::add(acc, V::ld(add + row * ld_add + c));
    if (act) acc = V::mask(acc, V::ld(act + row * ld_act + c));
    V::st(out + row * ld_out + c, acc);
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

inline int lanes_per_row(int C, int vec) {
  int chunks = (C + vec - 1) / vec;
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
    k_agg_fwd<4><<<blocks, 256, 0, s>>>(g.rowptr, g.col, g.n_nodes, in, ld_in, out, ld_out, C, add, ld_add, bias, relu, lpr);
  else
    k_agg_fwd<1><<<blocks, 256, 0, s>>>(g.rowptr, g.col, g.n_nodes, in, ld_in, out, ld_out, C, add, ld_add, bias, relu, lpr);
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
    k_agg_bwd<4><<<blocks, 256, 0, s>>>(g.rowptr, g.t_rowptr, g.t_col, g.n_nodes, gin, ld_gin, out, ld_out, C, add, ld_add, act, ld_act, lpr);
  else
    k_agg_bwd<1><<<blocks, 256, 0, s>>>(g.rowptr, g.t_rowptr, g.t_col, g.n_nodes, gin, ld_gin, out, ld_out, C, add, ld_add, act, ld_act, lpr);
  GN_LAUNCHED();
  return GNODE_OK;
}

}  // namespace gnode
