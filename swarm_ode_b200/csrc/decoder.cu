// position_decoder = nn.Linear(D, 2) applied to every saved time point of the solution
// (reference: scripts/train_gde.py:88-94).  One streaming pass over [n_t * n_nodes, D]: a warp per
// row, the two (<= 8) weight rows stay in L1.  Backward skips rows whose cotangent is all zero
// (training only feeds the current-agent rows of time point 1: scripts/train_gde.py:486-490).
#include "common.cuh"

namespace gnode {
namespace {

constexpr int kMaxOut = 8;
constexpr int DEC_THREADS = 256;

// Generic forward: a warp per row, weights re-read through L1 (any D, n_out <= 8).
__global__ void __launch_bounds__(DEC_THREADS) k_decoder_fwd(const float* __restrict__ x, int64_t M, int D, int n_out,
                                                             const float* __restrict__ w, const float* __restrict__ b,
                                                             float* __restrict__ out, float* __restrict__ copy_out) {
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
      if (copy_out) copy_out[m * D + c] = xv;
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

// The reference's decoder (n_out = 2, D <= 32 * CPL): each lane keeps its weight columns in registers and a warp
// streams TWO rows per iteration (2 * CPL independent loads in flight), so the pass is bound by HBM, not by L1.
template <int CPL>
__global__ void __launch_bounds__(DEC_THREADS) k_decoder_fwd2(const float* __restrict__ x, int64_t M, int D,
                                                              const float* __restrict__ w, const float* __restrict__ b,
                                                              float* __restrict__ out, float* __restrict__ copy_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float w0[CPL], w1[CPL];
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
    const int c = lane + 32 * k;
    w0[k] = (c < D) ? __ldg(w + c) : 0.f;
    w1[k] = (c < D) ? __ldg(w + D + c) : 0.f;
  }
  const float b0 = b ? __ldg(b) : 0.f, b1 = b ? __ldg(b + 1) : 0.f;
  for (int64_t m = 2 * warp; m < M; m += 2 * nwarps) {
    const bool two = (m + 1 < M);
    const float* xa = x + m * D;
    const float* xb = xa + D;
    float va[CPL], vb[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const int c = lane + 32 * k;
      va[k] = (c < D) ? __ldg(xa + c) : 0.f;
      vb[k] = (two && c < D) ? __ldg(xb + c) : 0.f;
    }
    if (copy_out) {      // the rows pass through registers anyway: also deliver them as a copy (sol[0] = y0 of the solver)
      float* ca = copy_out + m * D;
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        const int c = lane + 32 * k;
        if (c < D) { __stcs(ca + c, va[k]); if (two) __stcs(ca + D + c, vb[k]); }
      }
    }
    float a0 = 0.f, a1 = 0.f, c0 = 0.f, c1 = 0.f;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      a0 = fmaf(va[k], w0[k], a0); a1 = fmaf(va[k], w1[k], a1);
      c0 = fmaf(vb[k], w0[k], c0); c1 = fmaf(vb[k], w1[k], c1);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, s); a1 += __shfl_xor_sync(0xffffffffu, a1, s);
      c0 += __shfl_xor_sync(0xffffffffu, c0, s); c1 += __shfl_xor_sync(0xffffffffu, c1, s);
    }
    if (lane == 0) {
      out[m * 2] = a0 + b0; out[m * 2 + 1] = a1 + b1;
      if (two) { out[m * 2 + 2] = c0 + b0; out[m * 2 + 3] = c1 + b1; }
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
// Rows whose cotangent is all zero are dropped up front (training only feeds the current-agent rows of one time
// point, ~10 % of the rows: scripts/train_gde.py:486-490): the block compacts its active rows (order preserved, so
// the summation order is fixed) and streams only those rows of x, four rows in flight per thread.
constexpr int kWgRows = 256;   // rows per chunk
__global__ void __launch_bounds__(DEC_THREADS) k_decoder_wgrad(const float* __restrict__ x, const float* __restrict__ g,
                                                               int64_t M, int D, int n_out, int64_t n_chunks,
                                                               float* __restrict__ partials) {
  __shared__ float gs[kWgRows][kMaxOut];
  __shared__ int act[kWgRows];
  __shared__ int warp_cnt[DEC_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  float* my = partials + (int64_t)blockIdx.x * ((int64_t)n_out * D + n_out);
  float bsum = 0.f;                                   // bias gradient (threads < n_out)
  for (int c0 = 0; c0 < D; c0 += 2 * DEC_THREADS) {   // two columns per thread and pass (one pass for D <= 512)
    const int ca = c0 + tid, cb = c0 + DEC_THREADS + tid;
    float acca[kMaxOut], accb[kMaxOut];
#pragma unroll
    for (int o = 0; o < kMaxOut; ++o) { acca[o] = 0.f; accb[o] = 0.f; }
    for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {   // fixed chunk order per block
      const int64_t rbeg = ch * kWgRows;
      const int nrows = (int)((M - rbeg < kWgRows) ? (M - rbeg) : kWgRows);
      __syncthreads();                                 // previous chunk's gs / act are no longer read
      bool mine = false;
      if (tid < nrows) {
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) {
          const float v = (o < n_out) ? __ldg(g + (rbeg + tid) * n_out + o) : 0.f;
          gs[tid][o] = v;
          mine = mine || (v != 0.f);
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, mine);
      if (lane == 0) warp_cnt[wid] = __popc(bal);
      __syncthreads();
      int base = 0, na = 0;
#pragma unroll
      for (int i = 0; i < DEC_THREADS / 32; ++i) { if (i < wid) base += warp_cnt[i]; na += warp_cnt[i]; }
      if (mine) act[base + __popc(bal & ((1u << lane) - 1u))] = tid;   // ordered compaction
      __syncthreads();
      if (na == 0) continue;
      const float* xr = x + rbeg * D;
      int i = 0;
      for (; i + 4 <= na; i += 4) {
        const int r0 = act[i], r1 = act[i + 1], r2 = act[i + 2], r3 = act[i + 3];
        float xa[4] = {0.f, 0.f, 0.f, 0.f}, xb[4] = {0.f, 0.f, 0.f, 0.f};
        if (ca < D) { xa[0] = __ldg(xr + (int64_t)r0 * D + ca); xa[1] = __ldg(xr + (int64_t)r1 * D + ca);
                      xa[2] = __ldg(xr + (int64_t)r2 * D + ca); xa[3] = __ldg(xr + (int64_t)r3 * D + ca); }
        if (cb < D) { xb[0] = __ldg(xr + (int64_t)r0 * D + cb); xb[1] = __ldg(xr + (int64_t)r1 * D + cb);
                      xb[2] = __ldg(xr + (int64_t)r2 * D + cb); xb[3] = __ldg(xr + (int64_t)r3 * D + cb); }
        const int rr[4] = {r0, r1, r2, r3};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
          for (int o = 0; o < kMaxOut; ++o) {
            if (o < n_out) { acca[o] = fmaf(gs[rr[q]][o], xa[q], acca[o]); accb[o] = fmaf(gs[rr[q]][o], xb[q], accb[o]); }
          }
        }
      }
      for (; i < na; ++i) {
        const int r = act[i];
        const float xav = (ca < D) ? __ldg(xr + (int64_t)r * D + ca) : 0.f;
        const float xbv = (cb < D) ? __ldg(xr + (int64_t)r * D + cb) : 0.f;
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o)
          if (o < n_out) { acca[o] = fmaf(gs[r][o], xav, acca[o]); accb[o] = fmaf(gs[r][o], xbv, accb[o]); }
      }
      if (c0 == 0 && tid < n_out)
        for (int q = 0; q < na; ++q) bsum += gs[act[q]][tid];
    }
#pragma unroll
    for (int o = 0; o < kMaxOut; ++o) {
      if (o < n_out) {
        if (ca < D) my[(int64_t)o * D + ca] = acca[o];
        if (cb < D) my[(int64_t)o * D + cb] = accb[o];
      }
    }
  }
  if (tid < n_out) my[(int64_t)n_out * D + tid] = bsum;
}

int wgrad_blocks(int64_t M) {   // blocks stride over 256-row chunks
  int64_t b = ceil_div64(M, kWgRows);
  if (b > kNumSMs * 4) b = kNumSMs * 4;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace
}  // namespace gnode

namespace gnode {
// total[o * D + c] = sum_m g[m, o] * x[m, c],  total[n_out * D + o] = sum_m g[m, o]   (fixed-order reduction)
// partials: decoder_wgrad_partial_floats(M, D, n_out) floats; total: n_out * D + n_out floats (overwritten).
size_t decoder_wgrad_partial_floats(int64_t M, int D, int n_out) { return (size_t)wgrad_blocks(M) * ((size_t)n_out * D + n_out); }
int decoder_wgrad(const float* x, const float* g, int64_t M, int D, int n_out, float* partials, float* total, cudaStream_t s) {
  const int nb = wgrad_blocks(M);
  const int64_t per = (int64_t)n_out * D + n_out;
  k_decoder_wgrad<<<nb, DEC_THREADS, 0, s>>>(x, g, M, D, n_out, ceil_div64(M, kWgRows), partials);
  GN_LAUNCHED();
  GN_CUDA(cudaMemsetAsync(total, 0, sizeof(float) * per, s));
  return reduce_partials_accum(partials, nb, per, total, 1.f, s);
}
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
  return gnode_decoder_fwd_copy(x, m, node_dim, n_out, w, b, out, nullptr, stream);
}

extern "C" int gnode_decoder_fwd_copy(const float* x, int64_t m, int32_t node_dim, int32_t n_out, const float* w,
                                      const float* b, float* out, float* copy_out, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_ARG(n_out >= 1 && n_out <= kMaxOut && node_dim >= 1, "gnode_decoder_fwd: n_out must be in [1, %d]", kMaxOut);
  GN_ARG(x && w && out, "gnode_decoder_fwd: null pointer");
  GN_ARG(copy_out != x, "gnode_decoder_fwd_copy: copy_out aliases x");
  if (m == 0) return GNODE_OK;
  GN_PROF(s, 2.0 * m * node_dim * n_out, 4.0 * (double)m * (node_dim + n_out), "decoder_fwd");
  int64_t blocks = ceil_div64(m * 32, DEC_THREADS);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  if (n_out == 2 && node_dim <= 32 * 16) {
    if (node_dim <= 32 * 4) k_decoder_fwd2<4><<<(unsigned)blocks, DEC_THREADS, 0, s>>>(x, m, node_dim, w, b, out, copy_out);
    else if (node_dim <= 32 * 8) k_decoder_fwd2<8><<<(unsigned)blocks, DEC_THREADS, 0, s>>>(x, m, node_dim, w, b, out, copy_out);
    else if (node_dim <= 32 * 13) k_decoder_fwd2<13><<<(unsigned)blocks, DEC_THREADS, 0, s>>>(x, m, node_dim, w, b, out, copy_out);
    else k_decoder_fwd2<16><<<(unsigned)blocks, DEC_THREADS, 0, s>>>(x, m, node_dim, w, b, out, copy_out);
  } else {
    k_decoder_fwd<<<(unsigned)blocks, DEC_THREADS, 0, s>>>(x, m, node_dim, n_out, w, b, out, copy_out);
  }
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
    k_decoder_wgrad<<<nb, DEC_THREADS, 0, s>>>(x, grad_out, m, node_dim, n_out, ceil_div64(m, kWgRows), partials);
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
