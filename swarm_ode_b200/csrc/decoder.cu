// position_decoder = nn.Linear(D, 2) applied to every saved time point of the solution
// (reference: scripts/train_gde.py:88-94).  One streaming pass over [n_t * n_nodes, D]: a warp per
// row, the two (<= 8) weight rows stay in L1.  Backward skips rows whose cotangent is all zero
// (training only feeds the current-agent rows of time point 1: scripts/train_gde.py:486-490).
#include "common.cuh"

namespace gnode {
namespace {

constexpr int kMaxOut = 8;
constexpr int DEC_THREADS = 256;

__global__ void __launch_bounds__(DEC_THREADS) k_decoder_fwd(const float* __restrict__ x, int64_t M, int D, int n_out,
                                                             const float* __restrict__ w, const float* __restrict__ b,
                                                             float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t m = warp; m < M; m += nwarps) {
    float acc[kMaxOut];
#pragma unroll
    for (int o = 0; o < kMaxOut; ++o) acc[o] = 0.f;
    const float* xr = x + m * D;
    for (int c = lane; c < D; c += 32) {
      const float xv = __ldg(xr + c);
#pragma unroll
      for (int o = 0; o < kMaxOut; ++o)
        if (o < n_out) acc[o] = fmaf(xv, __ldg(w + (int64_t)o * D + c), acc[o]);
    }
#pragma unroll
    for (int o = 0; o < kMaxOut; ++o) {
      if (o < n_out) {
        float v = acc[o];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        if (lane == 0) out[m * n_out + o] = v + (b ? __ldg(b + o) : 0.f);
      }
    }
  }
}

// grad_x[m, c] = sum_o g[m, o] * w[o, c]
__global__ void __launch_bounds__(DEC_THREADS) k_decoder_dgrad(const float* __restrict__ g, int64_t M, int D, int n_out,
                                                               const float* __restrict__ w, float* __restrict__ gx) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t m = warp; m < M; m += nwarps) {
    float gv[kMaxOut];
    bool any = false;
#pragma unroll
    for (int o = 0; o < kMaxOut; ++o) {
      gv[o] = (o < n_out) ? __ldg(g + m * n_out + o) : 0.f;
      any = any || (gv[o] != 0.f);
    }
    float* gr = gx + m * D;
    if (!any) {
      for (int c = lane; c < D; c += 32) gr[c] = 0.f;
      continue;
    }
    for (int c = lane; c < D; c += 32) {
      float v = 0.f;
#pragma unroll
      for (int o = 0; o < kMaxOut; ++o)
        if (o < n_out) v = fmaf(gv[o], __ldg(w + (int64_t)o * D + c), v);
      gr[c] = v;
    }
  }
}

// partials[blk][o*D + c] = sum_{m in chunk} g[m,o] * x[m,c];  partials[blk][n_out*D + o] = sum g[m,o]
__global__ void __launch_bounds__(DEC_THREADS) k_decoder_wgrad(const float* __restrict__ x, const float* __restrict__ g,
                                                               int64_t M, int D, int n_out, int64_t rows_per_block,
                                                               float* __restrict__ partials) {
  const int tid = threadIdx.x;
  const int64_t rbeg = (int64_t)blockIdx.x * rows_per_block;
  const int64_t rend = (rbeg + rows_per_block < M) ? rbeg + rows_per_block : M;
  float* my = partials + (int64_t)blockIdx.x * ((int64_t)n_out * D + n_out);
  for (int c0 = 0; c0 < D; c0 += DEC_THREADS) {
    const int c = c0 + tid;
    float acc[kMaxOut];
#pragma unroll
    for (int o = 0; o < kMaxOut; ++o) acc[o] = 0.f;
    for (int64_t m = rbeg; m < rend; ++m) {
      float gv[kMaxOut];
      bool any = false;
#pragma unroll
      for (int o = 0; o < kMaxOut; ++o) {
        gv[o] = (o < n_out) ? __ldg(g + m * n_out + o) : 0.f;  // block-uniform
        any = any || (gv[o] != 0.f);
      }
      if (!any) continue;
      const float xv = (c < D) ? __ldg(x + m * D + c) : 0.f;
#pragma unroll
      for (int o = 0; o < kMaxOut; ++o) acc[o] = fmaf(gv[o], xv, acc[o]);
    }
    if (c < D) {
#pragma unroll
      for (int o = 0; o < kMaxOut; ++o)
        if (o < n_out) my[(int64_t)o * D + c] = acc[o];
    }
  }
  if (tid < n_out) {
    float s = 0.f;
    for (int64_t m = rbeg; m < rend; ++m) s += __ldg(g + m * n_out + tid);
    my[(int64_t)n_out * D + tid] = s;
  }
}

int wgrad_blocks(int64_t M) {
  int64_t b = ceil_div64(M, 256);
  if (b > kNumSMs * 8) b = kNumSMs * 8;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace
}  // namespace gnode

using namespace gnode;

extern "C" size_t gnode_decoder_workspace_bytes(int64_t m, int32_t node_dim, int32_t n_out) {
  Arena a(nullptr, 0);
  a.take<float>((size_t)wgrad_blocks(m) * ((size_t)n_out * node_dim + n_out));
  a.take<float>((size_t)n_out * node_dim + n_out);
  return a.off;
}

extern "C" int gnode_decoder_fwd(const float* x, int64_t m, int32_t node_dim, int32_t n_out, const float* w,
                                 const float* b, float* out, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_ARG(n_out >= 1 && n_out <= kMaxOut && node_dim >= 1, "gnode_decoder_fwd: n_out must be in [1, %d]", kMaxOut);
  GN_ARG(x && w && out, "gnode_decoder_fwd: null pointer");
  if (m == 0) return GNODE_OK;
  GN_PROF(s, 2.0 * m * node_dim * n_out, 4.0 * (double)m * (node_dim + n_out), "decoder_fwd");
  int64_t blocks = ceil_div64(m * 32, DEC_THREADS);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  k_decoder_fwd<<<(unsigned)blocks, DEC_THREADS, 0, s>>>(x, m, node_dim, n_out, w, b, out);
  GN_LAUNCHED();
  return GNODE_OK;
}

extern "C" int gnode_decoder_bwd(const float* x, const float* grad_out, int64_t m, int32_t node_dim, int32_t n_out,
                                 const float* w, float* grad_x, float* grad_w, float* grad_b, void* workspace,
                                 size_t workspace_bytes, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_ARG(n_out >= 1 && n_out <= kMaxOut && node_dim >= 1, "gnode_decoder_bwd: n_out must be in [1, %d]", kMaxOut);
  GN_ARG(grad_out && w, "gnode_decoder_bwd: null pointer");
  if (m == 0) return GNODE_OK;
  GN_PROF(s, 4.0 * m * node_dim * n_out, 4.0 * (double)m * (2 * node_dim + n_out), "decoder_bwd");
  if (grad_x) {
    int64_t blocks = ceil_div64(m * 32, DEC_THREADS);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    k_decoder_dgrad<<<(unsigned)blocks, DEC_THREADS, 0, s>>>(grad_out, m, node_dim, n_out, w, grad_x);
    GN_LAUNCHED();
  }
  if (grad_w || grad_b) {
    GN_ARG(x, "gnode_decoder_bwd: x is required for the weight gradient");
    Arena a(workspace, workspace_bytes);
    const int nb = wgrad_blocks(m);
    const int64_t per = (int64_t)n_out * node_dim + n_out;
    float* partials = a.take<float>((size_t)nb * per);
    float* total = a.take<float>((size_t)per);
    GN_ARENA_OK(a, "gnode_decoder_bwd");
    const int64_t rpb = ceil_div64(m, nb);
    k_decoder_wgrad<<<nb, DEC_THREADS, 0, s>>>(x, grad_out, m, node_dim, n_out, rpb, partials);
    GN_LAUNCHED();
    // fixed-order reduction of the per-block partials, then accumulate into the caller's grads
    GN_CUDA(cudaMemsetAsync(total, 0, sizeof(float) * per, s));
    GN_TRY(reduce_partials_accum(partials, nb, per, total, 1.f, s));
    if (grad_w) {
      LinComb lc{};
      lc.out = grad_w; lc.base = grad_w; lc.in[0] = total; lc.coef[0] = 1.f;
      lc.n_terms = 1; lc.n = (int64_t)n_out * node_dim;
      GN_TRY(lincomb(lc, s));
    }
    if (grad_b) {
      LinComb lc{};
      lc.out = grad_b; lc.base = grad_b; lc.in[0] = total + (int64_t)n_out * node_dim; lc.coef[0] = 1.f;
      lc.n_terms = 1; lc.n = n_out;
      GN_TRY(lincomb(lc, s));
    }
  }
  return GNODE_OK;
}
