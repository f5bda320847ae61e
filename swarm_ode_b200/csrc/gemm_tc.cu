// tcgen05 (5th-gen tensor core) engine for the dense contractions of the GNODE path, sm_100a only.
//
//   C[m, n] = epi( sum_k A[m, k] * B[n, k] )        A: fp32 activations [M, K] (row stride lda, any alignment)
//                                                    B: fp32 weights, PRE-SPLIT into tf32 hi / lo planes
//
// fp32-grade accuracy on the tensor pipe by the 3xTF32 split: a = a_hi + a_lo with a_hi = rna_tf32(a),
// a_lo = rna_tf32(a - a_hi); the product uses a_lo*b_hi + a_hi*b_lo + a_hi*b_hi (the dropped a_lo*b_lo
// term is < 2^-22 relative), accumulated in fp32 in tensor memory (TMEM).
//
// One CTA = one 128 x BN output tile (BN <= 256, multiple of 16), 5 warps:
//   warps 0-3  producers: coalesced global loads (lanes along K), hi/lo split in registers, scalar
//              conflict-free st.shared into the UMMA K-major no-swizzle layout, fence.proxy.async,
//              mbarrier arrive;  afterwards the same warps run the epilogue (tcgen05.ld of their TMEM
//              lane quadrant -> smem transpose -> coalesced global stores with bias/act/scale/base).
//   warp 4     TMEM alloc/dealloc; one elected lane issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8)
//              and tcgen05.commit to release smem stages / publish the accumulator.
// Two CTAs fit per SM (<= 100 KB smem, <= 256 TMEM columns each) so one tile's epilogue overlaps the
// other's main loop.
//
// Every mbarrier wait is bounded: on timeout the kernel records a status word and finishes instead of
// hanging (the host turns that into GNODE_ERR_CUDA).
#include "common.cuh"

namespace gnode {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 16;               // K elements per stage: 4 chunks of 16 bytes, 2 MMA K-steps
constexpr int CHUNKS = BK / 4;
constexpr int PROD_THREADS = 128;
constexpr int THREADS = 160;
constexpr int LBO_A = BM * 16 + 16;  // bytes between consecutive 16-byte K-chunks of the A tile (+16: bank spread)
constexpr uint32_t SPIN_LIMIT = 1u << 24;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(addr), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: returns false (and flags the status word) on timeout
__device__ __forceinline__ bool mbar_wait(uint32_t addr, uint32_t parity, int* status, int code) {
  for (uint32_t i = 0; i < SPIN_LIMIT; ++i)
    if (mbar_try_wait(addr, parity)) return true;
  if (status) atomicExch(status, code);
  return false;
}

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// UMMA shared-memory descriptor, K-major, no swizzle: 8-row x 16-byte core matrices, rows contiguous
// (SBO = 128 B between 8-row groups), LBO bytes between the two K-chunks of one K=8 step.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((128u >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE
}

// instruction descriptor: D = f32, A = B = tf32, both K-major, M = 128, N = bn
__device__ __forceinline__ uint32_t make_idesc(int bn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
}

struct Args {
  const float* A; int64_t lda;
  const float* Bhi; const float* Blo; int64_t ldb;  // pre-split weight planes, zero padded to [Npad16, Kpad16]
  float* C; int64_t ldc;
  int64_t M; int N; int K;
  int bn;        // N-tile width (multiple of 16, <= 256)
  int nstage;
  const float* bias; int act;
  const float* base; int64_t ldbase;
  float scale;
  int* status;
};

__global__ void __launch_bounds__(THREADS) k_gemm_tc(const Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full[4];
  __shared__ __align__(8) uint64_t bar_empty[4];
  __shared__ __align__(8) uint64_t bar_accum;
  __shared__ uint32_t tmem_holder;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * a.bn;
  const int npad = (a.N + 15) & ~15;
  const int bn = (npad - n0 < a.bn) ? (npad - n0) : a.bn;   // this tile's MMA N (multiple of 16)
  const int nkb = (a.K + BK - 1) / BK;
  const int NST = a.nstage;
  const uint32_t lbo_b = (uint32_t)a.bn * 16u + 16u;
  const uint32_t a_plane = CHUNKS * LBO_A;                   // bytes of one A plane (hi or lo) per stage
  const uint32_t b_plane = CHUNKS * lbo_b;
  const uint32_t stage_bytes = 2 * a_plane + 2 * b_plane;
  const uint32_t smem_base = smem_u32(smem);
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < a.bn) tmem_cols <<= 1;

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(smem_u32(&bar_full[s]), PROD_THREADS);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_accum), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_holder;

  if (warp < 4) {
    // =========================== producers ===========================
    const int half = lane >> 4;   // which of the two rows of this warp instruction
    const int kk = lane & 15;     // k inside the stage
    const uint32_t koff = (uint32_t)(kk >> 2) * 1u;  // chunk index
    bool ok = true;
    for (int kb = 0; kb < nkb && ok; ++kb) {
      const int s = kb % NST, it = kb / NST;
      if (it > 0) ok = mbar_wait(smem_u32(&bar_empty[s]), (uint32_t)((it - 1) & 1), a.status, 1);
      uint8_t* st = smem + (size_t)s * stage_bytes;
      float* a_hi = reinterpret_cast<float*>(st);
      float* a_lo = reinterpret_cast<float*>(st + a_plane);
      float* b_hi = reinterpret_cast<float*>(st + 2 * a_plane);
      float* b_lo = reinterpret_cast<float*>(st + 2 * a_plane + b_plane);
      const int64_t gk = (int64_t)kb * BK + kk;
      const bool kin = gk < a.K;
      // ---- A: 128 rows x 16 k ; this warp covers rows 8q + warp + 4*half, q = 0..15 ----
      float v[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int row = 8 * q + warp + 4 * half;
        const int64_t gm = m0 + row;
        v[q] = (kin && gm < a.M) ? __ldg(a.A + gm * a.lda + gk) : 0.f;
      }
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int row = 8 * q + warp + 4 * half;
        const uint32_t hi = to_tf32(v[q]);
        const uint32_t lo = to_tf32(v[q] - __uint_as_float(hi));
        const uint32_t w = (koff * LBO_A + (uint32_t)row * 16u + (uint32_t)(kk & 3) * 4u) >> 2;
        a_hi[w] = __uint_as_float(hi);
        a_lo[w] = __uint_as_float(lo);
      }
      // ---- B: bn rows x 16 k from the pre-split planes (already tf32-exact, zero padded) ----
      const int nq = bn >> 3;
      for (int q0 = 0; q0 < nq; q0 += 8) {
        float h[8], l[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int q = q0 + u;
          const int row = 8 * q + warp + 4 * half;
          const bool rin = q < nq;
          const int64_t off = (int64_t)(n0 + row) * a.ldb + gk;
          h[u] = rin ? __ldg(a.Bhi + off) : 0.f;
          l[u] = rin ? __ldg(a.Blo + off) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int q = q0 + u;
          if (q < nq) {
            const int row = 8 * q + warp + 4 * half;
            const uint32_t w = (koff * lbo_b + (uint32_t)row * 16u + (uint32_t)(kk & 3) * 4u) >> 2;
            b_hi[w] = h[u];
            b_lo[w] = l[u];
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA (async proxy)
      mbar_arrive(smem_u32(&bar_full[s]));
    }

    // =========================== epilogue ===========================
    ok = ok && mbar_wait(smem_u32(&bar_accum), 0u, a.status, 3);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* stg = reinterpret_cast<float*>(smem) + warp * (32 * 33);     // all MMAs retired: stage memory is free
    const uint32_t taddr_w = tmem_base + ((uint32_t)(32 * warp) << 16);
    for (int c0 = 0; c0 < bn; c0 += 32) {
      uint32_t r[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
            "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
            "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr_w + (uint32_t)c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(r[j]);   // row = lane, col = j
      __syncwarp();
      const int col = n0 + c0 + lane;
      const bool cin = (c0 + lane < bn) && (col < a.N);
      const float bv = (cin && a.bias) ? __ldg(a.bias + col) : 0.f;
      for (int rr = 0; rr < 32; ++rr) {
        const int64_t gm = m0 + 32 * warp + rr;
        if (gm < a.M && cin) {
          float x = stg[rr * 33 + lane] + bv;
          if (a.act == 1) x = fmaxf(x, 0.f);
          else if (a.act == 2) x = tanhf(x);
          x *= a.scale;
          if (a.base) x += __ldg(a.base + gm * a.ldbase + col);
          a.C[gm * a.ldc + col] = x;
        }
      }
      __syncwarp();
    }
  } else {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = make_idesc(bn);
    bool ok = true;
    for (int kb = 0; kb < nkb && ok; ++kb) {
      const int s = kb % NST, it = kb / NST;
      ok = mbar_wait(smem_u32(&bar_full[s]), (uint32_t)(it & 1), a.status, 2);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t sa = smem_base + (uint32_t)s * stage_bytes;
        const uint32_t a_hi = sa, a_lo = sa + a_plane, b_hi = sa + 2 * a_plane, b_lo = sa + 2 * a_plane + b_plane;
#pragma unroll
        for (int j = 0; j < BK / 8; ++j) {
          const uint64_t dah = make_desc(a_hi + 2 * j * LBO_A, LBO_A);
          const uint64_t dal = make_desc(a_lo + 2 * j * LBO_A, LBO_A);
          const uint64_t dbh = make_desc(b_hi + 2 * j * lbo_b, lbo_b);
          const uint64_t dbl = make_desc(b_lo + 2 * j * lbo_b, lbo_b);
          umma_tf32(tmem_base, dal, dbh, idesc, (kb > 0 || j > 0) ? 1u : 0u);   // small terms first
          umma_tf32(tmem_base, dah, dbl, idesc, 1u);
          umma_tf32(tmem_base, dah, dbh, idesc, 1u);
        }
        umma_commit(smem_u32(&bar_empty[s]));                 // frees the stage once these MMAs retire
        if (kb == nkb - 1) umma_commit(smem_u32(&bar_accum)); // accumulator complete
      }
      __syncwarp();
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
  }
}

// split a row-major weight matrix into tf32 hi / lo planes, zero padded to [rows_pad, cols_pad]
__global__ void k_presplit(const float* __restrict__ W, int rows, int cols, int64_t ld, float* __restrict__ hi,
                           float* __restrict__ lo, int rows_pad, int cols_pad) {
  const int64_t total = (int64_t)rows_pad * cols_pad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols_pad), c = (int)(i % cols_pad);
    const float v = (r < rows && c < cols) ? W[(int64_t)r * ld + c] : 0.f;
    const uint32_t h = to_tf32(v);
    hi[i] = __uint_as_float(h);
    lo[i] = __uint_as_float(to_tf32(v - __uint_as_float(h)));
  }
}

}  // namespace tc

size_t presplit_floats(int rows, int cols) {
  const size_t rp = (size_t)((rows + 15) & ~15), cp = (size_t)((cols + 15) & ~15);
  return 2 * rp * cp;
}

int presplit_weights(const float* W, int rows, int cols, int64_t ld, float* planes, cudaStream_t s) {
  const int rp = (rows + 15) & ~15, cp = (cols + 15) & ~15;
  const int64_t total = (int64_t)rp * cp;
  int64_t blocks = ceil_div64(total, 256);
  if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
  tc::k_presplit<<<(unsigned)blocks, 256, 0, s>>>(W, rows, cols, ld, planes, planes + total, rp, cp);
  GN_LAUNCHED();
  return GNODE_OK;
}

namespace {
__device__ int g_tc_status_word = 0;   // barrier-timeout status of the tcgen05 kernels (0 = ok)
int* status_ptr() {
  static int* p = nullptr;
  if (!p) cudaGetSymbolAddress(reinterpret_cast<void**>(&p), g_tc_status_word);
  return p;
}
}

bool gemm_nt_tc_supported(const GemmNT& g) {
  return g.Bsplit != nullptr && g.M >= 1 && g.N >= 16 && g.K >= 1;
}

int gemm_nt_tc(const GemmNT& g, cudaStream_t s) {
  if (g.M == 0 || g.N == 0) return GNODE_OK;
  int* status_dev = status_ptr();
  if (!status_dev) { set_error("gemm_nt_tc: cannot resolve the status symbol"); return GNODE_ERR_CUDA; }
  const int npad = (g.N + 15) & ~15, kpad = (g.K + 15) & ~15;
  const int ntiles = (npad + 255) / 256;
  int bn = ((npad + ntiles - 1) / ntiles + 15) & ~15;
  tc::Args a;
  a.A = g.A; a.lda = g.lda;
  a.Bhi = g.Bsplit; a.Blo = g.Bsplit + (size_t)npad * kpad; a.ldb = kpad;
  a.C = g.C; a.ldc = g.ldc; a.M = g.M; a.N = g.N; a.K = g.K; a.bn = bn;
  a.bias = g.bias; a.act = g.relu; a.base = g.base; a.ldbase = g.ldbase; a.scale = g.scale;
  a.status = status_dev;
  const size_t stage = 2 * (size_t)tc::CHUNKS * tc::LBO_A + 2 * (size_t)tc::CHUNKS * ((size_t)bn * 16 + 16);
  int nstage = (int)((100 * 1024) / stage);
  if (nstage > 4) nstage = 4;
  if (nstage < 2) nstage = 2;
  a.nstage = nstage;
  const size_t smem = stage * nstage;
  static size_t attr_set = 0;
  if (smem > attr_set) {
    GN_CUDA(cudaFuncSetAttribute(tc::k_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(110 * 1024)));
    attr_set = 110 * 1024;
  }
  dim3 grid((unsigned)ceil_div64(g.M, tc::BM), (unsigned)ntiles, 1);
  tc::k_gemm_tc<<<grid, tc::THREADS, smem, s>>>(a);
  GN_LAUNCHED();
  return GNODE_OK;
}

// reads and clears the tcgen05 status word (0 = ok); synchronises the stream
int gemm_tc_status(cudaStream_t s, int* out) {
  *out = 0;
  int* status_dev = status_ptr();
  if (!status_dev) return GNODE_OK;
  GN_CUDA(cudaMemcpyAsync(out, status_dev, sizeof(int), cudaMemcpyDeviceToHost, s));
  GN_CUDA(cudaStreamSynchronize(s));
  if (*out != 0) GN_CUDA(cudaMemsetAsync(status_dev, 0, sizeof(int), s));
  return GNODE_OK;
}

}  // namespace gnode

// Synchronises the stream and reports whether any tcgen05 kernel hit a barrier timeout since the last
// call (0 = healthy).  Meant for tests / debugging; the hot path never calls it.
extern "C" int gnode_tc_status(gnode_stream_t stream) {
  int st = 0;
  int rc = gnode::gemm_tc_status(static_cast<cudaStream_t>(stream), &st);
  if (rc != GNODE_OK) return rc;
  if (st != 0) {
    gnode::set_error("tcgen05 kernel barrier timeout (code %d: 1 = producer waiting for a free stage, 2 = MMA waiting "
                     "for operands, 3 = epilogue waiting for the accumulator)", st);
    return GNODE_ERR_CUDA;
  }
  return GNODE_OK;
}
