// Folded fixed-grid Runge-Kutta integration of the GraphODEFunc field (scripts/train_gde.py:33-45 integrated by
// torchdiffeq's fixed-grid solvers, :78-85) and its backward pass (loss.backward(), :493).
//
// conv1 and conv3 are linear maps around the nonlinear 2H-wide core of the field, and every stage input of an
// explicit RK step is  x_s = y + dt * sum_{j<s} beta_sj k_j  with  k_j = cat2_j @ w3cat^T + b3.  Hence
//
//   Z_s  = x_s @ w1cat^T = Z_0 + V_s @ M13^T + (dt * sum_j beta_sj) * c13,      V_s = dt * sum_{j<s} beta_sj cat2_j
//   y_1  = y + C @ w3cat^T + (dt * sum_s c_s) * b3,                             C   = dt * sum_s c_s cat2_s
//   M13  = w1cat @ w3cat  [2H, 2H],   c13 = w1cat @ b3  [2H]        (recomputed from the weights on every call)
//
// so the D-wide state is touched by TWO dense contractions per step (Z_0 on the way in, y_1 on the way out) instead
// of two per STAGE, and no D-wide stage buffer (x_s, k_s) ever exists.  The backward pass folds the same way:
//
//   gcat2_s = dt c_s (G @ w3cat) + U_s @ M13,      U_s = dt * sum_{i>s} beta_is gz_i          (gz_i = dL/dZ_i)
//   dW3cat += G^T C + w1cat^T R,   dW1cat += GZ^T y + R @ w3cat^T + g1 (x) b3,   db3 += (dt sum c_s) colsum(G) + g1 @ w1cat
//   grad_y  = G + GZ @ w1cat,      GZ = sum_s gz_s,  R = sum_s gz_s^T V_s,  g1 = sum_s (dt sum_j beta_sj) colsum(gz_s)
//
// i.e. per step two D-wide data contractions (G @ w3cat, GZ @ w1cat) and two D-wide weight-gradient contractions.
// This is a re-association of the reference arithmetic (fp32 rounding order only; parity gate: rel-L2 <= 1e-4); the
// unfolded path (gnode_set_fold(0)) stays available as the straightforward anchor.
#include <algorithm>
#include <cstring>

#include "field.cuh"

namespace gnode {

namespace {

// M13[z, c] = sum_d w1cat[z, d] w3cat[d, c];  M13T = transpose;  c13[z] = sum_d w1cat[z, d] b3[d]   (fp64 accumulate)
// Eight lanes share one output and split the D loop (d = l, l + 8, ...); fixed-order butterfly -> deterministic.
__global__ void k_fold_weights(const float* __restrict__ w1cat, const float* __restrict__ w3cat, const float* __restrict__ b3,
                               int H2, int D, float* __restrict__ M13, float* __restrict__ M13T, float* __restrict__ c13) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = gid >> 3, l = gid & 7;                   // cat2 feature (or H2 -> the bias column), lane of the split
  const int z = blockIdx.y;
  double acc = 0.0;
  const float* wz = w1cat + (size_t)z * D;
  if (c < H2) {
    for (int d = l; d < D; d += 8) acc += (double)wz[d] * (double)w3cat[(size_t)d * H2 + c];
  } else if (c == H2) {
    for (int d = l; d < D; d += 8) acc += (double)wz[d] * (double)b3[d];
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (l == 0) {
    if (c < H2) { M13[(size_t)z * H2 + c] = (float)acc; M13T[(size_t)c * H2 + z] = (float)acc; }
    else if (c == H2) c13[z] = (float)acc;
  }
}

// small dense products applied once at the end of the backward pass (fp32, tiny):
//   dW3cat[d, c] += sum_z w1cat[z, d] R[z, c]          dW1cat[z, d] += sum_c R[z, c] w3cat[d, c] + g1[z] b3[d]
//   db3[d]       += sum_z g1[z] w1cat[z, d]
// w3catT [2H, D] (the transposed copy the backward context holds anyway): consecutive threads then read consecutive
// addresses; w3cat[d, q] across threads d is a 512-byte stride = 32 lines per warp load, 128 times (the kernel took 46 us)
__global__ void k_fold_param_grads(const float* __restrict__ w1cat, const float* __restrict__ w3cat, const float* __restrict__ w3catT,
                                   const float* __restrict__ b3,
                                   const float* __restrict__ R, const float* __restrict__ g1, int H2, int D,
                                   float* __restrict__ dW3cat, float* __restrict__ dW1cat, float* __restrict__ db3) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;   // 0..H2-1: column c of dW3cat / row z of dW1cat;  H2: db3
  if (d >= D) return;
  if (j < H2) {
    float a3 = 0.f, a1 = 0.f;
    for (int q = 0; q < H2; ++q) {
      a3 = fmaf(w1cat[(size_t)q * D + d], R[(size_t)q * H2 + j], a3);       // z = q, c = j
      const float w3 = w3catT ? w3catT[(size_t)q * D + d] : w3cat[(size_t)d * H2 + q];
      a1 = fmaf(R[(size_t)j * H2 + q], w3, a1);                              // z = j, c = q
    }
    dW3cat[(size_t)d * H2 + j] += a3;
    dW1cat[(size_t)j * D + d] += a1 + g1[j] * b3[d];
  } else {
    float a = 0.f;
    for (int q = 0; q < H2; ++q) a = fmaf(g1[q], w1cat[(size_t)q * D + d], a);
    db3[d] += a;
  }
}

size_t padf(size_t floats) { return (floats + 63) & ~(size_t)63; }

// ---- factored cotangent G = g1 @ Wd (rank n_out <= 8) ----
// WdW3[o, c] = sum_d Wd[o, d] w3cat[d, c]      (fp64 accumulate; a warp per output, lanes split the D loop)
__global__ void k_lr_prep(const float* __restrict__ Wd, const float* __restrict__ w3cat, const float* __restrict__ w3catT, int n_out,
                          int D, int H2, float* __restrict__ WdW3) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, l = threadIdx.x & 31;
  if (i >= n_out * H2) return;
  const int o = i / H2, c = i % H2;
  double acc = 0.0;
  if (w3catT) {     // the transposed copy of the backward context: the lanes read consecutive addresses
    for (int d = l; d < D; d += 32) acc += (double)Wd[(size_t)o * D + d] * (double)w3catT[(size_t)c * D + d];
  } else {
    for (int d = l; d < D; d += 32) acc += (double)Wd[(size_t)o * D + d] * (double)w3cat[(size_t)d * H2 + c];
  }
#pragma unroll
  for (int sft = 1; sft < 32; sft <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
  if (l == 0) WdW3[i] = (float)acc;
}
// G3[n, c] = sum_o g1[n, o] WdW3[o, c]          (= (g1 @ Wd) @ w3cat, 2H-wide; one float4 per thread)
__global__ void k_lr_g3(const float* __restrict__ g1, const float* __restrict__ WdW3, int64_t N, int n_out, int H2,
                        float* __restrict__ G3) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;   // float4 index
  const int q = H2 / 4;
  if (i >= N * q) return;
  // (2H = 128: shifts; the general form is a 64-bit division by a run-time value per thread of a write-bound kernel)
  const int64_t n = H2 == 128 ? (i >> 5) : i / q;
  const int c = H2 == 128 ? ((int)(i & 31) << 2) : (int)(i % q) * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int o = 0; o < n_out; ++o) {
    const float g = __ldg(g1 + n * n_out + o);
    const float4 w = __ldg(reinterpret_cast<const float4*>(WdW3 + (size_t)o * H2 + c));
    acc.x = fmaf(g, w.x, acc.x); acc.y = fmaf(g, w.y, acc.y); acc.z = fmaf(g, w.z, acc.z); acc.w = fmaf(g, w.w, acc.w);
  }
  *reinterpret_cast<float4*>(G3 + n * H2 + c) = acc;
}
// The same for 2H = 128 and n_out <= 4: a warp per row (lane = one float4 of the 512-byte row), the rank-n_out weights in
// registers, four rows in flight per warp -- the kernel is a 199 MB write and should run at the store rate (the
// thread-per-float4 form above reloads the weights per element and ran at 2.8 TB/s).
template <int NO>
__global__ void __launch_bounds__(256) k_lr_g3_128(const float* __restrict__ g1, const float* __restrict__ WdW3, int64_t N,
                                                   float* __restrict__ G3) {
  constexpr int H2 = 128, RW = 4;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float4 w[NO];
#pragma unroll
  for (int o = 0; o < NO; ++o) w[o] = __ldg(reinterpret_cast<const float4*>(WdW3 + (size_t)o * H2) + lane);
  for (int64_t n0 = warp * RW; n0 < N; n0 += n_warps * RW) {
    float g[RW][NO];
#pragma unroll
    for (int u = 0; u < RW; ++u)
#pragma unroll
      for (int o = 0; o < NO; ++o) g[u][o] = (n0 + u < N) ? __ldg(g1 + (n0 + u) * NO + o) : 0.f;
#pragma unroll
    for (int u = 0; u < RW; ++u) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int o = 0; o < NO; ++o) {
        acc.x = fmaf(g[u][o], w[o].x, acc.x); acc.y = fmaf(g[u][o], w[o].y, acc.y);
        acc.z = fmaf(g[u][o], w[o].z, acc.z); acc.w = fmaf(g[u][o], w[o].w, acc.w);
      }
      if (n0 + u < N) *(reinterpret_cast<float4*>(G3 + (n0 + u) * H2) + lane) = acc;
    }
  }
}
// dW3cat[d, c] += sum_o Wd[o, d] X[o, c];   db3[d] += cs * sum_o Wd[o, d] X[n_out * H2 + o]     (X = g1^T [C | 1])
__global__ void k_lr_finish(const float* __restrict__ Wd, const float* __restrict__ X, int n_out, int D, int H2, float cs,
                            float* __restrict__ dW3cat, float* __restrict__ db3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D * (H2 + 1)) return;
  const int d = i / (H2 + 1), c = i % (H2 + 1);
  float acc = 0.f;
  if (c < H2) {
    for (int o = 0; o < n_out; ++o) acc = fmaf(Wd[(size_t)o * D + d], X[(size_t)o * H2 + c], acc);
    dW3cat[(size_t)d * H2 + c] += acc;
  } else {
    for (int o = 0; o < n_out; ++o) acc = fmaf(Wd[(size_t)o * D + d], X[(size_t)n_out * H2 + o], acc);
    db3[d] += cs * acc;
  }
}

// ---- position decoder of the next time point from the 2H-wide step combination (integrate_fixed_folded with `dec`) ----
//   y_1 = y + C @ w3cat^T + cs b3   =>   y_1 @ Wd^T + bd = (y @ Wd^T + bd) + C @ (Wd @ w3cat)^T + cs (Wd @ b3)
// P[o * 2H + c] = sum_d Wd[o, d] w3cat[d, c],   P[n_out * 2H + o] = sum_d Wd[o, d] b3[d]      (fp64 accumulate, a warp per entry)
__global__ void k_dec_prep(const float* __restrict__ Wd, const float* __restrict__ w3cat, const float* __restrict__ b3, int n_out,
                           int D, int H2, float* __restrict__ P) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, l = threadIdx.x & 31;
  if (i >= n_out * (H2 + 1)) return;
  const int o = i / (H2 + 1), c = i % (H2 + 1);
  double acc = 0.0;
  if (c < H2) { for (int d = l; d < D; d += 32) acc += (double)Wd[(size_t)o * D + d] * (double)w3cat[(size_t)d * H2 + c]; }
  else { for (int d = l; d < D; d += 32) acc += (double)Wd[(size_t)o * D + d] * (double)b3[d]; }
#pragma unroll
  for (int sft = 1; sft < 32; sft <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
  if (l == 0) P[c < H2 ? o * H2 + c : n_out * H2 + o] = (float)acc;
}
// next[n, o] = prev[n, o] + sum_c C[n, c] P[o, c] + cs P[o, 2H]      (2H = 128: a warp per row, one float4 of C per lane,
// four rows in flight per warp; fixed-order butterfly -> deterministic)
template <int NO>
__global__ void __launch_bounds__(256) k_decode_step(const float* __restrict__ C, const float* __restrict__ P, const float* __restrict__ prev,
                                                     float* __restrict__ next, int64_t N, float cs) {
  constexpr int H2 = 128, RW = 4;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float4 pw[NO];
#pragma unroll
  for (int o = 0; o < NO; ++o) pw[o] = __ldg(reinterpret_cast<const float4*>(P + (size_t)o * H2) + lane);
  for (int64_t r0 = warp * RW; r0 < N; r0 += n_warps * RW) {
    float4 v[RW];
#pragma unroll
    for (int u = 0; u < RW; ++u)
      v[u] = (r0 + u < N) ? __ldcs(reinterpret_cast<const float4*>(C + (r0 + u) * H2) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    float acc[RW][NO];
#pragma unroll
    for (int u = 0; u < RW; ++u)
#pragma unroll
      for (int o = 0; o < NO; ++o)
        acc[u][o] = fmaf(v[u].x, pw[o].x, fmaf(v[u].y, pw[o].y, fmaf(v[u].z, pw[o].z, v[u].w * pw[o].w)));
#pragma unroll
    for (int sft = 16; sft >= 1; sft >>= 1)
#pragma unroll
      for (int u = 0; u < RW; ++u)
#pragma unroll
        for (int o = 0; o < NO; ++o) acc[u][o] += __shfl_xor_sync(0xffffffffu, acc[u][o], sft);
    if (lane < RW * NO) {
      const int u = lane / NO, o = lane % NO;
      float val = 0.f;
#pragma unroll
      for (int uu = 0; uu < RW; ++uu)
#pragma unroll
        for (int oo = 0; oo < NO; ++oo)
          if (uu == u && oo == o) val = acc[uu][oo];
      if (r0 + u < N) next[(r0 + u) * NO + o] = prev[(r0 + u) * NO + o] + val + cs * __ldg(P + (size_t)NO * H2 + o);
    }
  }
}

}  // namespace

void FoldWs::carve(Arena& a, const Sage3Ctx& c, int S_, bool backward) {
  S = S_;
  const size_t nh = (size_t)c.N * 2 * c.H, nd = (size_t)c.N * c.D;
  const int H2 = 2 * c.H;
  M13 = a.take<float>((size_t)H2 * H2);
  M13T = a.take<float>((size_t)H2 * H2);
  c13 = a.take<float>(H2);
  sM13 = a.take<float>(presplit_floats(H2, H2));
  sM13T = a.take<float>(presplit_floats(H2, H2));
  if (chain_shape_ok(c.H)) {
    ci13 = a.take<float>(chain_image_floats(H2, H2));
    ci13T = a.take<float>(chain_image_floats(H2, H2));
  }
  z0 = a.take<float>(nh);
  Cbuf = a.take<float>(nh);
  Vbuf = a.take<float>(nh);
  if (chain_shape_ok(c.H)) maskws = a.take<uint32_t>((size_t)S * padf((size_t)c.N * 4));
  if (backward) {
    G3 = a.take<float>(nh);
    U = a.take<float>(nh);
    GZ = a.take<float>(nh);
    for (int i = 0; i < S; ++i) gzs[i] = a.take<float>(nh);
    if (chain_shape_ok(c.H)) {
      float* us = a.take<float>(nh * S);
      float* gv = a.take<float>(nh / 2 * S);
      for (int i = 0; i < S; ++i) { Us[i] = us ? us + (size_t)i * nh : nullptr; gv2s[i] = gv ? gv + (size_t)i * (nh / 2) : nullptr; }
    }
    gcur = a.take<float>(nd);
    gnext = a.take<float>(nd);
    R = a.take<float>((size_t)H2 * H2);
    g1 = a.take<float>(H2);
    cs = a.take<float>(H2);
    size_t pf = gemm_tn_workspace_floats(c.D, H2, c.N);
    const size_t cand[3] = {gemm_tn_workspace_floats(H2, c.D, c.N), gemm_tn_workspace_floats(H2, H2, c.N * S),
                            gemm_tn_workspace_floats(c.H, H2, c.N * S)};
    for (size_t v : cand) if (v > pf) pf = v;
    partials = a.take<float>(pf);
  }
}

size_t FoldWs::save_floats_per_step(const Sage3Ctx& c, int S_) {
  return (size_t)(2 * S_ + 1) * padf((size_t)c.N * 2 * c.H) +   // cat1, cat2 per stage; C
         (chain_shape_ok(c.H) ? (size_t)S_ * padf((size_t)c.N * 4) : 0);   // ReLU sign bits per stage
}

// point the stage slots of step j at the save area (or at the workspace when save == null)
void FoldWs::bind_slots(Sage3Ctx& c, float* save, int j) {
  const size_t nh = padf((size_t)c.N * 2 * c.H);
  if (save) {
    float* p = save + (size_t)j * save_floats_per_step(c, S);
    for (int st = 0; st < S; ++st) { cat1[st] = p + (size_t)st * nh; cat2[st] = p + (size_t)(S + st) * nh; }   // two contiguous stacks
    Cslot = p + (size_t)2 * S * nh;
    uint32_t* mp = reinterpret_cast<uint32_t*>(p + (size_t)(2 * S + 1) * nh);
    for (int st = 0; st < S; ++st) mask[st] = chain_shape_ok(c.H) ? mp + (size_t)st * padf((size_t)c.N * 4) : nullptr;
  } else {
    for (int st = 0; st < S; ++st) mask[st] = maskws ? maskws + (size_t)st * padf((size_t)c.N * 4) : nullptr;
    for (int st = 0; st < S; ++st) { cat1[st] = c.cat1[st]; cat2[st] = c.cat2[st]; }
    Cslot = Cbuf;
  }
}

int FoldWs::prepare(Sage3Ctx& c, cudaStream_t s) {
  const int H2 = 2 * c.H;
  {
    GN_PROF(s, 2.0 * H2 * (H2 + 1) * c.D, 0.0, "fold_weights");
    dim3 grid((unsigned)ceil_div64((int64_t)(H2 + 1) * 8, 256), (unsigned)H2);
    k_fold_weights<<<grid, 256, 0, s>>>(c.w1cat, c.w3cat, c.b3, H2, c.D, M13, M13T, c13);
    GN_LAUNCHED();
  }
  pend_sM13 = pend_sM13T = c.use_tc ? 1 : 0;                                    // packed by their first reader
  pend_ci13 = pend_ci13T = (c.use_tc && chain_shape_ok(c.H)) ? 1 : 0;
  if (c.use_tc && lazy_images_poisoned()) {
    GN_CUDA(cudaMemsetAsync(sM13, 0xFF, sizeof(float) * presplit_floats(H2, H2), s));
    GN_CUDA(cudaMemsetAsync(sM13T, 0xFF, sizeof(float) * presplit_floats(H2, H2), s));
    if (chain_shape_ok(c.H)) {
      GN_CUDA(cudaMemsetAsync(ci13, 0xFF, sizeof(float) * chain_image_floats(H2, H2), s));
      GN_CUDA(cudaMemsetAsync(ci13T, 0xFF, sizeof(float) * chain_image_floats(H2, H2), s));
    }
  }
  return GNODE_OK;
}

// Stages of one step: fills cat1[st], cat2[st] from y.  Z_0 is left in z0.
int FoldWs::forward_stages(Sage3Ctx& c, const Tableau& tb, const float* y, float dt, cudaStream_t s, float* Cout, float* Cout2,
                           const double* coef2) {
  const int H = c.H, H2 = 2 * c.H;
  const int64_t N = c.N;
  const int64_t nh = N * H2;
  if (!z0_ready) {  // Z_0 = y @ w1cat^T
    GemmNT q{};
    q.A = y; q.lda = c.D; q.B = c.w1cat; q.ldb = c.D; q.C = z0; q.ldc = H2; q.M = N; q.N = H2; q.K = c.D;
    c.use_w1(q);
    GN_TRY(gemm_nt(q, s));
  }
  z_next_valid = false;
  if (chain_fwd_supported(c)) {   // all stages of the step, graph-resident
    GN_TRY(chain_fwd(c, *this, tb, dt, Cout, s, Cout2, coef2));
    z_next_valid = z_next != nullptr && tb.S > 1;
    return GNODE_OK;
  }
  for (int st = 0; st < S; ++st) {
    const float* z = z0;
    if (st > 0) {
      // V_st = dt * sum_{j<st} beta[st][j] cat2_j ;  Z_st = Z_0 + V_st @ M13^T + (dt sum_j beta) c13
      LinComb lc{};
      lc.out = Vbuf; lc.base = nullptr; lc.n = nh; lc.n_terms = 0;
      double bsum = 0.0;
      for (int j = 0; j < st; ++j) {
        lc.in[lc.n_terms] = cat2[j]; lc.coef[lc.n_terms] = (float)tb.beta[st][j] * dt; ++lc.n_terms;
        bsum += tb.beta[st][j];
      }
      GN_TRY(lincomb(lc, s));
      GemmNT q{};
      q.A = Vbuf; q.lda = H2; q.B = M13; q.ldb = H2; q.C = c.z; q.ldc = H2; q.M = N; q.N = H2; q.K = H2;
      q.bias = c13; q.bias_scale = (float)bsum * dt; q.base = z0; q.ldbase = H2;
      use_M13(c, q);
      GN_TRY(gemm_nt(q, s));
      z = c.z;
    }
    float* c1 = cat1[st];
    float* c2 = cat2[st];
    GN_TRY(agg_mean_fwd(c.g, z, H2, c1 + H, H2, H, z + H, H2, c.b1, 1, s));          // h1
    GN_TRY(agg_mean_fwd(c.g, c1 + H, H2, c1, H2, H, nullptr, 0, nullptr, 0, s));     // A(h1)
    {
      GemmNT q{};
      q.A = c1; q.lda = H2; q.B = c.w2cat; q.ldb = H2; q.C = c2 + H; q.ldc = H2; q.M = N; q.N = H; q.K = H2;
      q.bias = c.b2; q.relu = 1;
      c.use_w2(q);
      GN_TRY(gemm_nt(q, s));                                                          // h2
    }
    GN_TRY(agg_mean_fwd(c.g, c2 + H, H2, c2, H2, H, nullptr, 0, nullptr, 0, s));     // A(h2)
  }
  if (Cout) GN_TRY(combine_solution(c, tb, dt, Cout, s));
  if (Cout && Cout2 && coef2) {
    LinComb lc{};
    lc.out = Cout2; lc.base = nullptr; lc.n = nh; lc.n_terms = 0;
    for (int j = 0; j < S; ++j)
      if (coef2[j] != 0.0) { lc.in[lc.n_terms] = cat2[j]; lc.coef[lc.n_terms] = (float)coef2[j] * dt; ++lc.n_terms; }
    GN_TRY(lincomb(lc, s));
  }
  return GNODE_OK;
}

// C = dt * sum_s c_sol[s] cat2_s   -> Cbuf
int FoldWs::combine_solution(Sage3Ctx& c, const Tableau& tb, float dt, float* out, cudaStream_t s) {
  LinComb lc{};
  lc.out = out; lc.base = nullptr; lc.n = c.N * 2 * c.H; lc.n_terms = 0;
  for (int st = 0; st < S; ++st) { lc.in[lc.n_terms] = cat2[st]; lc.coef[lc.n_terms] = (float)tb.c_sol[st] * dt; ++lc.n_terms; }
  return lincomb(lc, s);
}

int integrate_fixed_folded(Sage3Ctx& c, FoldWs& f, const Tableau& tb, const float* y0, const float* t, int n_t,
                           float* sol, float* save, cudaStream_t s, bool sol0_by_caller, const DecodeLR* dec) {
  const int H2 = 2 * c.H;
  const int64_t n = c.numel();
  if (sol != y0 && !sol0_by_caller) GN_CUDA(cudaMemcpyAsync(sol, y0, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
  GN_TRY(f.prepare(c, s));
  if (dec) {
    if (H2 != 128 || dec->n_out < 1 || dec->n_out > kDecodeLRMaxOut) { set_error("integrate_fixed_folded: decoder shape not supported"); return GNODE_ERR_ARG; }
    GN_PROF(s, 2.0 * dec->n_out * (H2 + 1) * c.D, 0.0, "decoder_prep");
    k_dec_prep<<<(unsigned)ceil_div64((int64_t)dec->n_out * (H2 + 1) * 32, 256), 256, 0, s>>>(dec->Wd, c.w3cat, c.b3, dec->n_out, c.D, H2, dec->P);
    GN_LAUNCHED();
  }
  f.forward_only = save == nullptr;     // without a save area a later backward recomputes the stages itself
  double csum = 0.0;
  for (int st = 0; st < tb.S; ++st) csum += tb.c_sol[st];
  for (int j = 0; j + 1 < n_t; ++j) {
    const float dt = t[j + 1] - t[j];
    const float* y = (j == 0 && sol0_by_caller) ? y0 : sol + (int64_t)j * n;
    float* y1 = sol + (int64_t)(j + 1) * n;
    f.bind_slots(c, save, j);
    GN_TRY(f.forward_stages(c, tb, y, dt, s, f.Cslot));
    GemmNT q{};   // y_1 = y + C @ w3cat^T + (dt sum c) b3
    q.A = f.Cslot; q.lda = H2; q.B = c.w3cat; q.ldb = H2; q.C = y1; q.ldc = c.D; q.M = c.N; q.N = c.D; q.K = H2;
    q.bias = c.b3; q.bias_scale = (float)csum * dt; q.base = y; q.ldbase = c.D;
    c.use_w3(q); q.rows_engine = 1;
    GN_TRY(gemm_nt(q, s));
    if (dec) {
      const float* prev = dec->traj + (int64_t)j * c.N * dec->n_out;
      float* next = dec->traj + (int64_t)(j + 1) * c.N * dec->n_out;
      const float cs = (float)csum * dt;
      GN_PROF(s, 2.0 * c.N * H2 * dec->n_out, 4.0 * (double)c.N * (H2 + 2 * dec->n_out), "decoder_step");
      const unsigned grid = (unsigned)std::min<int64_t>(ceil_div64(c.N, 4 * 8), (int64_t)kNumSMs * 8);
      switch (dec->n_out) {
        case 1: k_decode_step<1><<<grid, 256, 0, s>>>(f.Cslot, dec->P, prev, next, c.N, cs); break;
        case 2: k_decode_step<2><<<grid, 256, 0, s>>>(f.Cslot, dec->P, prev, next, c.N, cs); break;
        case 3: k_decode_step<3><<<grid, 256, 0, s>>>(f.Cslot, dec->P, prev, next, c.N, cs); break;
        default: k_decode_step<4><<<grid, 256, 0, s>>>(f.Cslot, dec->P, prev, next, c.N, cs); break;
      }
      GN_LAUNCHED();
    }
  }
  return GNODE_OK;
}

// Backward of ONE folded step whose stage slots (f.cat1 / f.cat2 / sign bits) and C (f.Cslot, combined with tb.c_sol) are
// in place: accumulates the parameter gradients (dW*, db*, R, g1) for the cotangent G of
//     y_out = y + dt * sum_s tb.c_sol[s] k_s
// and leaves GZ = sum_s dL/dZ_s in f.GZ.  tb.c_sol may be any weight vector (dense-output weights for dopri5).
// G may be given in factored form (lr).
static int fold_step_bwd(Sage3Ctx& c, FoldWs& f, const Tableau& tb, float dt, const float* y, const float* G,
                         const LowRankG* lr, cudaStream_t s) {
  const int S = tb.S, H = c.H, H2 = 2 * c.H;
  const int64_t N = c.N, nh = N * H2;
  double csum = 0.0;
  for (int st = 0; st < S; ++st) csum += tb.c_sol[st];
  if (lr) {  // G = g1 @ Wd is never formed:  G3 = g1 @ (Wd @ w3cat)
    GN_PROF(s, 2.0 * N * lr->n_out * H2, 4.0 * (double)N * (H2 + lr->n_out), "lowrank_G3");
    k_lr_prep<<<(unsigned)ceil_div64((int64_t)lr->n_out * H2 * 32, 256), 256, 0, s>>>(lr->Wd, c.w3cat, c.w3catT, lr->n_out, c.D, H2, lr->WdW3);
    GN_LAUNCHED();
    if (H2 == 128 && lr->n_out <= 4) {
      const unsigned grid = (unsigned)std::min<int64_t>(ceil_div64(N, 4 * 8), (int64_t)kNumSMs * 8);
      switch (lr->n_out) {
        case 1: k_lr_g3_128<1><<<grid, 256, 0, s>>>(lr->g1, lr->WdW3, N, f.G3); break;
        case 2: k_lr_g3_128<2><<<grid, 256, 0, s>>>(lr->g1, lr->WdW3, N, f.G3); break;
        case 3: k_lr_g3_128<3><<<grid, 256, 0, s>>>(lr->g1, lr->WdW3, N, f.G3); break;
        default: k_lr_g3_128<4><<<grid, 256, 0, s>>>(lr->g1, lr->WdW3, N, f.G3); break;
      }
    } else {
      k_lr_g3<<<(unsigned)ceil_div64(N * (H2 / 4), 256), 256, 0, s>>>(lr->g1, lr->WdW3, N, lr->n_out, H2, f.G3);
    }
    GN_LAUNCHED();
  } else {  // G3 = G @ w3cat     [N, 2H]
    GemmNT q{};
    q.A = G; q.lda = c.D; q.B = c.w3catT; q.ldb = c.D; q.C = f.G3; q.ldc = H2; q.M = N; q.N = H2; q.K = c.D;
    c.use_w3T(q);
    GN_TRY(gemm_nt(q, s));
  }
  if (chain_bwd_supported(c, f)) {
    // all stages of this step in one graph-resident kernel (chain_bwd.cu), then the weight-gradient contractions
    // over its outputs: dW2cat += g_v2_s^T cat1_s (db2 += colsum g_v2_s), R += U_s^T cat2_s (g1 += colsum U_s)
    bool has_u[kMaxStages];
    GN_TRY(chain_bwd(c, f, tb, dt, has_u, s));
    // stages stacked along the row dimension: one contraction per operand pair when the slots are contiguous
    bool stacked = true;
    int n_u = 0;
    for (int st = 0; st < S; ++st) {
      if (st > 0 && (f.cat1[st] != f.cat1[st - 1] + nh || f.cat2[st] != f.cat2[st - 1] + nh)) stacked = false;
      if (has_u[st]) { if (st != n_u) stacked = false; ++n_u; }
    }
    for (int st = 0; st < (stacked ? 1 : S); ++st) {
      GemmTN q{};
      q.A = f.gv2s[st]; q.lda = H; q.P = H; q.B = f.cat1[st]; q.ldb = H2; q.Q = H2; q.Nrows = stacked ? N * S : N;
      q.C = c.dW2cat; q.ldc = H2; q.colsumA = c.db2;
      GN_TRY(gemm_tn(q, f.partials, s));
    }
    for (int st = 0; st < (stacked ? (n_u > 0 ? 1 : 0) : S); ++st) {
      if (!stacked && !has_u[st]) continue;
      GemmTN q{};
      q.A = f.Us[st]; q.lda = H2; q.P = H2; q.B = f.cat2[st]; q.ldb = H2; q.Q = H2; q.Nrows = stacked ? N * n_u : N;
      q.C = f.R; q.ldc = H2; q.colsumA = f.g1;
      GN_TRY(gemm_tn(q, f.partials, s));
    }
  } else {
  for (int st = S - 1; st >= 0; --st) {
      float* gz = f.gzs[st];
      // U_st = dt * sum_{i>st} beta[i][st] gz_i
      LinComb lu{};
      lu.out = f.U; lu.base = nullptr; lu.n = nh; lu.n_terms = 0;
      for (int i = st + 1; i < S; ++i)
        if (tb.beta[i][st] != 0.0) { lu.in[lu.n_terms] = f.gzs[i]; lu.coef[lu.n_terms] = (float)tb.beta[i][st] * dt; ++lu.n_terms; }
      const bool has_u = lu.n_terms > 0;
      const float cs_dt = (float)tb.c_sol[st] * dt;
      if (!has_u && cs_dt == 0.f) {   // the stage does not influence the output
        GN_CUDA(cudaMemsetAsync(gz, 0, sizeof(float) * nh, s));
        continue;
      }
      if (has_u) {
        GN_TRY(lincomb(lu, s));
        {
          // R += U_st^T @ cat2_st  [2H, 2H]  (= sum_s gz_s^T V_s regrouped by cat2_j, so that V_s need not be kept);
          // g1 += colsum(U_st)  (= sum_s (dt sum_j beta_sj) colsum(gz_s)), fused into the same pass
          GemmTN q{};
          q.A = f.U; q.lda = H2; q.P = H2; q.B = f.cat2[st]; q.ldb = H2; q.Q = H2; q.Nrows = N; q.C = f.R; q.ldc = H2;
          q.colsumA = f.g1;
          GN_TRY(gemm_tn(q, f.partials, s));
        }
        GemmNT q{};   // gcat = dt c_st G3 + U @ M13
        q.A = f.U; q.lda = H2; q.B = f.M13T; q.ldb = H2; q.C = c.gcat; q.ldc = H2; q.M = N; q.N = H2; q.K = H2;
        q.base = f.G3; q.ldbase = H2; q.base_scale = cs_dt;
        f.use_M13T(c, q);
        GN_TRY(gemm_nt(q, s));
      } else {
        LinComb lg{};
        lg.out = c.gcat; lg.base = nullptr; lg.n = nh; lg.n_terms = 1; lg.in[0] = f.G3; lg.coef[0] = cs_dt;
        GN_TRY(lincomb(lg, s));
      }
      const float* c1 = f.cat1[st];
      const float* c2 = f.cat2[st];
      // ---- conv3 -> conv2 ----   g_v2 = (A^T(gcat_l) + gcat_r) * [h2 > 0]
      GN_TRY(agg_mean_bwd(c.g, c.gcat, H2, c.gv2, H, H, c.gcat + H, H2, c2 + H, H2, s));
      {  // gcat = g_v2 @ w2cat   [N, 2H]
        GemmNT q{};
        q.A = c.gv2; q.lda = H; q.B = c.w2catT; q.ldb = H; q.C = c.gcat; q.ldc = H2; q.M = N; q.N = H2; q.K = H;
        c.use_w2T(q);
        GN_TRY(gemm_nt(q, s));
      }
      {  // dW2cat += g_v2^T @ cat1
        GemmTN q{};
        q.A = c.gv2; q.lda = H; q.P = H; q.B = c1; q.ldb = H2; q.Q = H2; q.Nrows = N; q.C = c.dW2cat; q.ldc = H2;
        q.colsumA = c.db2;                                          // db2 += colsum(g_v2), fused
        GN_TRY(gemm_tn(q, f.partials, s));
      }
      // ---- conv2 -> conv1 ----   g_u1 = (A^T(gcat_l) + gcat_r) * [h1 > 0] -> gz[:, H:] ; A^T(g_u1) -> gz[:, :H]
      GN_TRY(agg_mean_bwd(c.g, c.gcat, H2, gz + H, H2, H, c.gcat + H, H2, c1 + H, H2, s));
      GN_TRY(agg_mean_bwd(c.g, gz + H, H2, gz, H2, H, nullptr, 0, nullptr, 0, s));
    }
    // ---- D-wide parameter gradients of this step ----
    {
      LinComb lz{};
      lz.out = f.GZ; lz.base = nullptr; lz.n = nh; lz.n_terms = 0;
      for (int st = 0; st < S; ++st) { lz.in[lz.n_terms] = f.gzs[st]; lz.coef[lz.n_terms] = 1.f; ++lz.n_terms; }
      GN_TRY(lincomb(lz, s));
    }
  }
  if (lr) {  // dW3cat += Wd^T (g1^T C),  db3 += (dt sum c_s) Wd^T colsum(g1): only rows with a cotangent are read
    GN_PROF(s, 2.0 * N * lr->n_out * H2, 4.0 * (double)N * lr->n_out, "lowrank_dW3");
    GN_TRY(decoder_wgrad(f.Cslot, lr->g1, N, H2, lr->n_out, lr->partials, lr->X, s));
    k_lr_finish<<<(unsigned)ceil_div64((int64_t)c.D * (H2 + 1), 256), 256, 0, s>>>(lr->Wd, lr->X, lr->n_out, c.D, H2,
                                                                                  (float)csum * dt, c.dW3cat, c.db3);
    GN_LAUNCHED();
  } else {  // dW3cat += G^T @ C     [D, 2H]   (C kept by the forward pass)
    GemmTN q{};
    q.A = G; q.lda = c.D; q.P = c.D; q.B = f.Cslot; q.ldb = H2; q.Q = H2; q.Nrows = N; q.C = c.dW3cat; q.ldc = H2;
    q.colsumA = c.db3; q.colsumA_scale = (float)csum * dt;       // db3 += (dt sum c_s) colsum(G), fused
    GN_TRY(gemm_tn(q, f.partials, s));
  }
  {  // dW1cat += GZ^T @ y    [2H, D];   colsum(GZ) = sum_s colsum(gz_s): its right half is db1
    GN_CUDA(cudaMemsetAsync(f.cs, 0, sizeof(float) * H2, s));
    GemmTN q{};
    q.A = f.GZ; q.lda = H2; q.P = H2; q.B = y; q.ldb = c.D; q.Q = c.D; q.Nrows = N; q.C = c.dW1cat; q.ldc = c.D;
    q.colsumA = f.cs;
    GN_TRY(gemm_tn(q, f.partials, s));
    LinComb l1{};
    l1.out = c.db1; l1.base = c.db1; l1.n = H; l1.n_terms = 1; l1.in[0] = f.cs + H; l1.coef[0] = 1.f;
    GN_TRY(lincomb(l1, s));
  }
  return GNODE_OK;
}

int integrate_fixed_folded_bwd(Sage3Ctx& c, FoldWs& f, const Tableau& tb, const float* sol, const float* t, int n_t,
                               const float* grad_sol, float* grad_y0, const float* save, cudaStream_t s,
                               const LowRankG* lr) {
  const int S = tb.S, H = c.H, H2 = 2 * c.H;
  const int64_t N = c.N, n = c.numel(), nh = N * H2;
  GN_TRY(f.prepare(c, s));
  // R and g1 are consecutive arena blocks (FoldWs::carve): one fill
  GN_CUDA(cudaMemsetAsync(f.R, 0, (size_t)(reinterpret_cast<char*>(f.g1 + H2) - reinterpret_cast<char*>(f.R)), s));

  // G = cotangent of y_{j+1}: explicit part from grad_sol plus what flowed back from later steps
  const float* G = lr ? nullptr : grad_sol + (int64_t)(n_t - 1) * n;
  float* gout = f.gcur;
  for (int j = n_t - 2; j >= 0; --j) {
    const float dt = t[j + 1] - t[j];
    const float* y = sol + (int64_t)j * n;
    f.bind_slots(c, const_cast<float*>(save), j);
    if (!save) GN_TRY(f.forward_stages(c, tb, y, dt, s, f.Cslot));  // recompute this step's stages (and C)
    GN_TRY(fold_step_bwd(c, f, tb, dt, y, G, lr, s));
    if (j == 0 && grad_y0 == nullptr) break;   // nobody asked for dL/dy_0: skip its D-wide contraction
    {  // cotangent of y_j:  G + GZ @ w1cat + grad_sol[j]
      GemmNT q{};
      q.A = f.GZ; q.lda = H2; q.B = c.w1catT; q.ldb = H2; q.C = gout; q.ldc = c.D; q.M = N; q.N = c.D; q.K = H2;
      q.base = G; q.ldbase = c.D; q.base2 = grad_sol + (int64_t)j * n; q.ldbase2 = c.D;
      c.use_w1T(q);
      GN_TRY(gemm_nt(q, s));
    }
    G = gout;
    gout = (gout == f.gcur) ? f.gnext : f.gcur;
  }
  if (grad_y0) GN_CUDA(cudaMemcpyAsync(grad_y0, G, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
  {
    GN_PROF(s, 4.0 * H2 * H2 * c.D, 0.0, "fold_param_grads");
    dim3 grid((unsigned)ceil_div64(c.D, 128), (unsigned)(H2 + 1));
    k_fold_param_grads<<<grid, 128, 0, s>>>(c.w1cat, c.w3cat, c.w3catT, c.b3, f.R, f.g1, H2, c.D, c.dW3cat, c.dW1cat, c.db3);
    GN_LAUNCHED();
  }
  return GNODE_OK;
}

// ------------------------------------------------------------------------------------------------
// Backward through dopri5 (backprop through the solver, like the reference's loss.backward() through
// torchdiffeq's odeint; step sizes are constants of the differentiation, as under torchdiffeq's
// @torch.no_grad() step-size controller).
//
// The forward pass chose the accepted steps tau_0 < tau_1 < ... (trace of gnode_integrate_dopri5).  With those fixed,
// the solve is a fixed-grid integration with the 7-stage Dormand-Prince tableau, and every requested output time in
// (tau_k, tau_{k+1}] is the dense-output quartic of step k, which is itself a stage combination
//     y(x) = y_k + dt sum_s w_s(x) k_s,      x = (t - tau_k) / dt
//     w_s(x) = x d_s0 + x^2 (d_s6 - 4 d_s0 - 5 c_s + 16 m_s) + x^3 (5 d_s0 - 3 d_s6 + 14 c_s - 32 m_s)
//            + x^4 (2 d_s6 - 2 d_s0 - 8 c_s + 16 m_s)                      (c = c_sol, m = c_mid, d = Kronecker delta)
// (oracle/torchdiffeq_ref.py:_interp_fit_dopri5 expanded; w(1) = c_sol).  So the backward of step k is fold_step_bwd
// once per cotangent source: the cotangent of y_{k+1} with weights c_sol, and every output inside the step with
// weights w(x).  States y_k are replayed first (deterministic kernels: bitwise the forward's), the stages of a step are
// recomputed when its backward runs.
void dopri5_dense_weights(const Tableau& tb, double x, double* w) {
  const double x2 = x * x, x3 = x2 * x, x4 = x3 * x;
  for (int s = 0; s < 7; ++s) {
    const double d0 = s == 0 ? 1.0 : 0.0, d6 = s == 6 ? 1.0 : 0.0, cs = tb.c_sol[s], ms = tb.c_mid[s];
    w[s] = x * d0 + x2 * (d6 - 4.0 * d0 - 5.0 * cs + 16.0 * ms) + x3 * (5.0 * d0 - 3.0 * d6 + 14.0 * cs - 32.0 * ms) +
           x4 * (2.0 * d6 - 2.0 * d0 - 8.0 * cs + 16.0 * ms);
  }
}

int integrate_dopri5_folded_bwd(Sage3Ctx& c, FoldWs& f, const float* y0, const double* tau, int n_acc, const double* t,
                                int n_t, const float* grad_sol, float* grad_y0, float* ys /* [n_acc + 1, N, D] */,
                                cudaStream_t s) {
  const Tableau& tb = *tableau_for(GNODE_DOPRI5);
  const int H2 = 2 * c.H;
  const int64_t N = c.N, n = c.numel();
  GN_TRY(f.prepare(c, s));
  // R and g1 are consecutive arena blocks (FoldWs::carve): one fill
  GN_CUDA(cudaMemsetAsync(f.R, 0, (size_t)(reinterpret_cast<char*>(f.g1 + H2) - reinterpret_cast<char*>(f.R)), s));
  f.bind_slots(c, nullptr, 0);
  double csum = 0.0;
  for (int st = 0; st < tb.S; ++st) csum += tb.c_sol[st];

  // ---- replay the accepted steps: ys[k] = y(tau_k) ----
  GN_CUDA(cudaMemcpyAsync(ys, y0, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
  for (int k = 0; k + 1 < n_acc; ++k) {            // the last state is never a step input
    const float dt = (float)(tau[k + 1] - tau[k]);
    const float* y = ys + (int64_t)k * n;
    GN_TRY(f.forward_stages(c, tb, y, dt, s, f.Cslot));
    GemmNT q{};   // y_{k+1} = y_k + C @ w3cat^T + (dt sum c) b3
    q.A = f.Cslot; q.lda = H2; q.B = c.w3cat; q.ldb = H2; q.C = ys + (int64_t)(k + 1) * n; q.ldc = c.D; q.M = N; q.N = c.D; q.K = H2;
    q.bias = c.b3; q.bias_scale = (float)csum * dt; q.base = y; q.ldbase = c.D;
    c.use_w3(q);
    q.rows_engine = 1;   // the replay takes no step-size decisions: the row-major engine may compute it
    GN_TRY(gemm_nt(q, s));
  }

  // ---- backward over the steps, last to first ----
  const float* Gnext = nullptr;      // cotangent of y_{k+1} flowing back from later steps
  float* gout = f.gcur;
  int out_hi = n_t - 1;              // outputs still to be assigned to a step (descending)
  for (int k = n_acc - 1; k >= 0; --k) {
    const double t0 = tau[k], t1 = tau[k + 1];
    const float dt = (float)(t1 - t0);
    const float* y = ys + (int64_t)k * n;
    // outputs of this step: t0 < t_i <= t1 (the forward interpolates an output in the first step that reaches it)
    int out_lo = out_hi;
    while (out_lo >= 1 && t[out_lo] > t0) --out_lo;
    const int n_out_here = out_hi - out_lo;        // outputs out_lo + 1 .. out_hi
    const int n_src = n_out_here + (Gnext ? 1 : 0);
    if (n_src == 0) continue;                       // nothing depends on this step (cannot happen before the last output)
    bool stages_ready = false, first = true;
    for (int q = 0; q < n_src; ++q) {
      Tableau tw = tb;
      const float* G;
      if (q < n_out_here) {
        const int i = out_lo + 1 + q;
        dopri5_dense_weights(tb, (t[i] - t0) / (t1 - t0), tw.c_sol);
        G = grad_sol + (int64_t)i * n;
      } else {
        G = Gnext;
      }
      if (!stages_ready) {
        GN_TRY(f.forward_stages(c, tw, y, dt, s, f.Cslot));     // stages of this step and C for the first source
        stages_ready = true;
      } else {
        GN_TRY(f.combine_solution(c, tw, dt, f.Cslot, s));      // same stages, this source's weights
      }
      GN_TRY(fold_step_bwd(c, f, tw, dt, y, G, nullptr, s));
      {  // cotangent of y_k  (+)=  G + GZ @ w1cat
        GemmNT g{};
        g.A = f.GZ; g.lda = H2; g.B = c.w1catT; g.ldb = H2; g.C = gout; g.ldc = c.D; g.M = N; g.N = c.D; g.K = H2;
        g.base = G; g.ldbase = c.D;
        if (!first) { g.base2 = gout; g.ldbase2 = c.D; }
        c.use_w1T(g);
        GN_TRY(gemm_nt(g, s));
      }
      first = false;
    }
    out_hi = out_lo;
    Gnext = gout;
    gout = (gout == f.gcur) ? f.gnext : f.gcur;
  }
  if (grad_y0) {
    // dL/dy_0 = what flowed back through the steps + the cotangent of the output at t[0] (sol[0] = y0)
    LinComb lc{};
    lc.out = grad_y0; lc.base = grad_sol; lc.n = n; lc.n_terms = 0;
    if (Gnext) { lc.in[0] = Gnext; lc.coef[0] = 1.f; lc.n_terms = 1; }
    GN_TRY(lincomb(lc, s));
  }
  {
    GN_PROF(s, 4.0 * H2 * H2 * c.D, 0.0, "fold_param_grads");
    dim3 grid((unsigned)ceil_div64(c.D, 128), (unsigned)(H2 + 1));
    k_fold_param_grads<<<grid, 128, 0, s>>>(c.w1cat, c.w3cat, c.w3catT, c.b3, f.R, f.g1, H2, c.D, c.dW3cat, c.dW1cat, c.db3);
    GN_LAUNCHED();
  }
  return GNODE_OK;
}

}  // namespace gnode
