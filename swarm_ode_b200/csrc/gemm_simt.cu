// fp32 FFMA dense contractions -- the parity anchor of the GNODE path (GNODE_ENGINE_SIMT) and
// the fallback for shapes the tcgen05 path does not take.
//
// Stands for the per-node Linear layers inside SAGEConv (lin_l / lin_r; reference call sites
// scripts/train_gde.py:27-29) and their weight gradients (autograd of scripts/train_gde.py:493).
//
// One kernel, two operand layouts:
//   NT  C[m,n] = sum_k A[m,k] B[n,k]      activations x weights (both K-contiguous)
//   TN  C[p,q] = sum_r A[r,p] B[r,q]      weight gradients: reduction over node rows, split over
//                                          row chunks (grid.z) into partials, reduced in fixed order.
// Tile 128x64x16, 128 threads, 8x8 register micro-tile, double-buffered shared memory.
#include "common.cuh"

namespace gnode {
namespace {

constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN_ = 8, THREADS = 128;
constexpr int PAD = 4;

struct SgemmArgs {
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  float* C; int64_t ldc;
  int64_t M; int64_t N; int64_t K;
  int64_t kchunk;          // K range per grid.z slice (== K when not split)
  int64_t c_split_stride;  // elements between partial outputs of consecutive grid.z slices
  const float* bias; int relu;
  const float* base; int64_t ldbase;
  float scale;
  float bias_scale, base_scale;
  const float* base2; int64_t ldbase2;
  int post_relu;
};

template <bool TRANS>
__global__ void __launch_bounds__(THREADS) k_sgemm(const SgemmArgs a) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];

  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int64_t n0 = (int64_t)blockIdx.y * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * a.kchunk;
  const int64_t kend = (kbeg + a.kchunk < a.K) ? kbeg + a.kchunk : a.K;
  const int nk = (int)((kend - kbeg + BK - 1) / BK);

  float ra[BM * BK / THREADS];  // 16
  float rb[BN * BK / THREADS];  // 8

  auto load_global = [&](int kt) {
    const int64_t k0 = kbeg + (int64_t)kt * BK;
    if (!TRANS) {
      const int k = tid & (BK - 1);
      const int mb = tid >> 4;  // 0..7
      const bool kok = (k0 + k) < kend;
#pragma unroll
      for (int i = 0; i < BM * BK / THREADS; ++i) {
        const int64_t m = m0 + mb + i * (THREADS / BK);
        ra[i] = (kok && m < a.M) ? __ldg(a.A + m * a.lda + k0 + k) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < BN * BK / THREADS; ++i) {
        const int64_t n = n0 + mb + i * (THREADS / BK);
        rb[i] = (kok && n < a.N) ? __ldg(a.B + n * a.ldb + k0 + k) : 0.f;
      }
    } else {
      const int64_t m = m0 + tid;  // BM == THREADS
#pragma unroll
      for (int i = 0; i < BK; ++i) {
        const int64_t k = k0 + i;
        ra[i] = (k < kend && m < a.M) ? __ldg(a.A + k * a.lda + m) : 0.f;
      }
      const int64_t n = n0 + (tid & (BN - 1));
      const int kb = tid >> 6;  // 0..1
#pragma unroll
      for (int i = 0; i < BN * BK / THREADS; ++i) {
        const int64_t k = k0 + kb + i * (THREADS / BN);
        rb[i] = (k < kend && n < a.N) ? __ldg(a.B + k * a.ldb + n) : 0.f;
      }
    }
  };
  auto store_smem = [&](int buf) {
    if (!TRANS) {
      const int k = tid & (BK - 1);
      const int mb = tid >> 4;
#pragma unroll
      for (int i = 0; i < BM * BK / THREADS; ++i) As[buf][k][mb + i * (THREADS / BK)] = ra[i];
#pragma unroll
      for (int i = 0; i < BN * BK / THREADS; ++i) Bs[buf][k][mb + i * (THREADS / BK)] = rb[i];
    } else {
#pragma unroll
      for (int i = 0; i < BK; ++i) As[buf][i][tid] = ra[i];
      const int n = tid & (BN - 1);
      const int kb = tid >> 6;
#pragma unroll
      for (int i = 0; i < BN * BK / THREADS; ++i) Bs[buf][kb + i * (THREADS / BN)][n] = rb[i];
    }
  };

  const int tx = tid & 7;   // n group
  const int ty = tid >> 3;  // m group (0..15)
  float acc[TM][TN_];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN_; ++j) acc[i][j] = 0.f;

  if (nk > 0) {
    load_global(0);
    store_smem(0);
  }
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_global(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float av[TM], bv[TN_];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * TN_]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * TN_ + 4]);
      av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w;
      av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
      bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
      bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN_; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_smem(buf ^ 1);
    __syncthreads();
  }

  float* C = a.C + (int64_t)blockIdx.z * a.c_split_stride;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t m = m0 + ty * TM + i;
    if (m >= a.M) continue;
#pragma unroll
    for (int j = 0; j < TN_; ++j) {
      const int64_t n = n0 + tx * TN_ + j;
      if (n >= a.N) continue;
      float v = acc[i][j];
      if (a.bias) v += a.bias_scale * __ldg(a.bias + n);
      if (a.relu == 1) v = fmaxf(v, 0.f);
      else if (a.relu == 2) v = tanhf(v);
      v *= a.scale;
      if (a.base) v += a.base_scale * __ldg(a.base + m * a.ldbase + n);
      if (a.base2) v += __ldg(a.base2 + m * a.ldbase2 + n);
      if (a.post_relu) v = fmaxf(v, 0.f);
      C[m * a.ldc + n] = v;
    }
  }
}

// C[p*ldc + q] += scale * sum_s partials[s][p*Q + q]     (fixed order over s)
__global__ void k_reduce_partials(const float* __restrict__ partials, int S, int64_t PQ, int Q,
                                  float* __restrict__ C, int64_t ldc, float scale) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= PQ) return;
  float s = 0.f;
  for (int z = 0; z < S; ++z) s += partials[(int64_t)z * PQ + i];
  int64_t p = i / Q, q = i % Q;
  C[p * ldc + q] += scale * s;
}

// Same sum for many partials (S >= 64): a block owns 32 outputs, its 8 warps each take a contiguous slice of s, and the
// 8 slice sums are added in slice order -- still a fixed order, but 8 x more loads in flight and 8 x more blocks.
constexpr int RP_SLICES = 8;
__global__ void __launch_bounds__(32 * RP_SLICES) k_reduce_partials_sliced(const float* __restrict__ partials, int S,
                                                                            int64_t PQ, int Q, float* __restrict__ C,
                                                                            int64_t ldc, float scale) {
  __shared__ float part[RP_SLICES][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t i = blockIdx.x * 32LL + lane;
  const int per = (S + RP_SLICES - 1) / RP_SLICES;
  const int z0 = w * per, z1 = (z0 + per < S) ? z0 + per : S;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (i < PQ) {
    int z = z0;
    for (; z + 4 <= z1; z += 4) {
      s0 += partials[(int64_t)z * PQ + i];
      s1 += partials[(int64_t)(z + 1) * PQ + i];
      s2 += partials[(int64_t)(z + 2) * PQ + i];
      s3 += partials[(int64_t)(z + 3) * PQ + i];
    }
    for (; z < z1; ++z) s0 += partials[(int64_t)z * PQ + i];
  }
  part[w][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (w == 0 && i < PQ) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < RP_SLICES; ++k) s += part[k][lane];
    int64_t p = i / Q, q = i % Q;
    C[p * ldc + q] += scale * s;
  }
}

static void launch_reduce_partials(const float* partials, int S, int64_t PQ, int Q, float* C, int64_t ldc, float scale,
                                   cudaStream_t s) {
  if (S >= 64)
    k_reduce_partials_sliced<<<(unsigned)ceil_div64(PQ, 32), 32 * RP_SLICES, 0, s>>>(partials, S, PQ, Q, C, ldc, scale);
  else
    k_reduce_partials<<<(unsigned)ceil_div64(PQ, 256), 256, 0, s>>>(partials, S, PQ, Q, C, ldc, scale);
}

constexpr int CS_THREADS = 256;
constexpr int CS_COLS = 64;
// partials[blk][c] = sum over this block's row chunk of X[r, c]
__global__ void __launch_bounds__(CS_THREADS) k_colsum_partial(const float* __restrict__ X, int64_t ldx,
                                                                int64_t Nrows, int C, int64_t rows_per_block,
                                                                float* __restrict__ partials) {
  __shared__ float red[CS_THREADS];
  const int tid = threadIdx.x;
  const int cl = tid & (CS_COLS - 1);
  const int rl = tid >> 6;  // 0..3
  const int64_t rbeg = (int64_t)blockIdx.x * rows_per_block;
  const int64_t rend = (rbeg + rows_per_block < Nrows) ? rbeg + rows_per_block : Nrows;
  for (int c0 = 0; c0 < C; c0 += CS_COLS) {
    const int c = c0 + cl;
    float acc = 0.f;
    if (c < C)
      for (int64_t r = rbeg + rl; r < rend; r += CS_THREADS / CS_COLS) acc += __ldg(X + r * ldx + c);
    red[tid] = acc;
    __syncthreads();
    if (tid < CS_COLS && c < C)
      partials[(int64_t)blockIdx.x * C + c] = (red[tid] + red[tid + 64]) + (red[tid + 128] + red[tid + 192]);
    __syncthreads();
  }
}

int tn_splits(int P, int Q, int64_t Nrows) {
  int64_t tiles = ceil_div64(P, BM) * ceil_div64(Q, BN);
  int64_t want = ceil_div64((int64_t)kNumSMs * 4, tiles);
  int64_t max_s = ceil_div64(Nrows, (int64_t)BK * 8);
  if (max_s < 1) max_s = 1;
  if (want > max_s) want = max_s;
  if (want < 1) want = 1;
  return (int)want;
}

int colsum_blocks(int64_t Nrows) {
  int64_t b = ceil_div64(Nrows, 512);
  if (b > kNumSMs * 4) b = kNumSMs * 4;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

int gemm_nt_simt(const GemmNT& g, cudaStream_t s) {
  if (g.M == 0 || g.N == 0) return GNODE_OK;
  SgemmArgs a;
  a.A = g.A; a.lda = g.lda; a.B = g.B; a.ldb = g.ldb; a.C = g.C; a.ldc = g.ldc;
  a.M = g.M; a.N = g.N; a.K = g.K; a.kchunk = g.K > 0 ? g.K : 1; a.c_split_stride = 0;
  a.bias = g.bias; a.relu = g.relu; a.base = g.base; a.ldbase = g.ldbase; a.scale = g.scale;
  a.bias_scale = g.bias_scale; a.base_scale = g.base_scale; a.base2 = g.base2; a.ldbase2 = g.ldbase2;
  a.post_relu = g.post_relu;
  dim3 grid((unsigned)ceil_div64(g.M, BM), (unsigned)ceil_div64(g.N, BN), 1);
  k_sgemm<false><<<grid, THREADS, 0, s>>>(a);
  GN_LAUNCHED();
  return GNODE_OK;
}

size_t gemm_tn_simt_workspace_floats(int P, int Q, int64_t Nrows) {
  const size_t a = (size_t)tn_splits(P, Q, Nrows) * (size_t)P * (size_t)Q;
  const size_t b = (size_t)colsum_blocks(Nrows) * (size_t)(P > Q ? P : Q);
  return a > b ? a : b;
}

int gemm_tn_simt(const GemmTN& g, float* partials, cudaStream_t s) {
  if (g.P == 0 || g.Q == 0 || g.Nrows == 0) return GNODE_OK;
  const int S = tn_splits(g.P, g.Q, g.Nrows);
  int64_t kchunk = ceil_div64(ceil_div64(g.Nrows, S), BK) * BK;
  if (kchunk < BK) kchunk = BK;
  SgemmArgs a;
  a.A = g.A; a.lda = g.lda; a.B = g.B; a.ldb = g.ldb; a.C = partials; a.ldc = g.Q;
  a.M = g.P; a.N = g.Q; a.K = g.Nrows; a.kchunk = kchunk; a.c_split_stride = (int64_t)g.P * g.Q;
  a.bias = nullptr; a.relu = 0; a.base = nullptr; a.ldbase = 0; a.scale = 1.f;
  a.bias_scale = 1.f; a.base_scale = 1.f; a.base2 = nullptr; a.ldbase2 = 0; a.post_relu = 0;
  dim3 grid((unsigned)ceil_div64(g.P, BM), (unsigned)ceil_div64(g.Q, BN), (unsigned)S);
  k_sgemm<true><<<grid, THREADS, 0, s>>>(a);
  GN_LAUNCHED();
  const int64_t PQ = (int64_t)g.P * g.Q;
  launch_reduce_partials(partials, S, PQ, g.Q, g.C, g.ldc, g.scale, s);
  GN_LAUNCHED();
  if (g.colsumA) GN_TRY(colsum_accum(g.A, g.lda, g.Nrows, g.P, g.colsumA, g.colsumA_scale, partials, s));
  if (g.colsumB) GN_TRY(colsum_accum(g.B, g.ldb, g.Nrows, g.Q, g.colsumB, g.colsumB_scale, partials, s));
  return GNODE_OK;
}

// out[i] += scale * sum_{z < S} partials[z * count + i]   (fixed order)
int reduce_partials_accum(const float* partials, int S, int64_t count, float* out, float scale, cudaStream_t s) {
  if (count == 0) return GNODE_OK;
  launch_reduce_partials(partials, S, count, (int)count, out, count, scale, s);
  GN_LAUNCHED();
  return GNODE_OK;
}

size_t colsum_workspace_floats(int C, int64_t Nrows) { return (size_t)colsum_blocks(Nrows) * (size_t)C; }

int colsum_accum(const float* X, int64_t ldx, int64_t Nrows, int C, float* out, float scale,
                 float* partials, cudaStream_t s) {
  if (C == 0) return GNODE_OK;
  GN_PROF(s, (double)Nrows * C, 4.0 * (double)Nrows * C, "colsum C=%d", C);
  const int nb = colsum_blocks(Nrows);
  const int64_t rpb = ceil_div64(Nrows > 0 ? Nrows : 1, nb);
  k_colsum_partial<<<nb, CS_THREADS, 0, s>>>(X, ldx, Nrows, C, rpb, partials);
  GN_LAUNCHED();
  launch_reduce_partials(partials, nb, C, C, out, C, scale, s);
  GN_LAUNCHED();
  return GNODE_OK;
}

}  // namespace gnode
