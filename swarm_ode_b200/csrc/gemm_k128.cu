// NT contraction with a 2H = 128 wide reduction and a WIDE, arbitrarily aligned output (sm_100a, tcgen05):
//
//   C[m, n] = base_scale * base[m, n] + scale * ( sum_k A[m, k] B[n, k] + bias_scale * bias[n] )
//
// This is the shape of the D-wide projections of the folded integrator (fold.cu): y_1 = y + C @ w3cat^T + b3.  N = D
// (399: rows 4-byte aligned only), K = 128, so the kernel is all epilogue: 1.44 GB of base + output traffic against
// 0.2 GB of operand.  The A tile [128 x 128] is loaded once, coalesced, into the UMMA operand layout and read in place by
// tcgen05.mma.kind::tf32; the three-term product of the chain kernels (chain_common.cuh: tile residual as a bf16
// operand in tensor memory, weight images streamed per K block of 16) gives fp32-grade accuracy without a converted
// copy.  (A column-chunked predecessor of the row-major kernel below -- 80-column accumulator chunks, transposed
// epilogue -- measured the same as the general engine of gemm_tc.cu and was removed in round 2; shapes the row-major
// kernel does not take go to gemm_tc.cu.)
#include <cstdlib>

#include "chain_common.cuh"

namespace gnode {
namespace k128 {
using namespace chain;

// Output columns per accumulator chunk: five chunks cover the row (all of it stays in tensor memory, 448 columns next to
// the 64-column residual operand): 80 for N <= 400 (D = 399: the medium warehouse), 88 for N <= 440 (D = 435: 19 AGVs +
// 9 pickers).  The weight image is chunked accordingly.
constexpr int NCH_A = 80, NCH_B = 88, MAXN_A = 400, MAXN_B = 440;
__host__ __device__ constexpr int nch_of(int n) { return n <= MAXN_A ? NCH_A : NCH_B; }
constexpr int NKB = W2H / KB16;                       // 8

struct Args {
  const float* A; const float* img;                   // A [M, 128] dense rows; chunked weight image
  float* C; int64_t ldc;
  int64_t M; int N; int n_chunks;
  const float* bias; float bias_scale;
  const float* base; int64_t ldbase; float base_scale;
  const float* base2; int64_t ldbase2;
  float scale;
  int* status;
  int dev_flags;   // K128_TRACE builds: bit 0 = weight producer fetches half of every stage (timing experiment)
};

// chunked weight image: chunk c holds rows [NCH c, NCH c + NCH) of W [n x 128] (zeros beyond n), each chunk in the chain
// format (chain_common.cuh) with one stage per K block of 16
__global__ void k_pack_k128(const float* __restrict__ W, int n, int64_t ld, uint4* __restrict__ img, int n_chunks, int NCH) {
  const int units_per_chunk_plane = NCH + 1, units_per_stage = 10 * units_per_chunk_plane;
  const int64_t total = (int64_t)n_chunks * NKB * units_per_stage;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int u = (int)(i % units_per_stage);
    const int kb = (int)((i / units_per_stage) % NKB);
    const int c = (int)(i / ((int64_t)units_per_stage * NKB));
    const int chunk = u / units_per_chunk_plane, rl = u % units_per_chunk_plane;
    const int row = c * NCH + rl;
    uint4 out = make_uint4(0u, 0u, 0u, 0u);
    if (rl < NCH && row < n) {
      const float* src = W + (size_t)row * ld + kb * KB16;
      uint32_t v[4];
      if (chunk < 8) {
        const int cc = chunk & 3;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float w = src[4 * cc + e];
          const uint32_t hb = (__float_as_uint(w) + 0x1000u) & 0xFFFFE000u;
          v[e] = chunk < 4 ? hb : __float_as_uint(w - __uint_as_float(hb));
        }
      } else {
        const int cc = chunk - 8;
#pragma unroll
        for (int e = 0; e < 4; ++e) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(v[e]) : "f"(src[8 * cc + 2 * e + 1]), "f"(src[8 * cc + 2 * e]));
      }
      out = make_uint4(v[0], v[1], v[2], v[3]);
    }
    img[i] = out;
  }
}

}  // namespace k128


// ---------------------------------------------------------------------------------------------------------------------
// ROW-MAJOR variant for C = base_scale * base + scale * (A B^T + bias_scale * bias) with DENSE rows (ldc = ldbase = N):
// the y_1 = y + C w3cat^T + b3 projection of every folded step.
//
// Copy probes (scripts/dev/probe_l2prefetch.cu, probe_pitch.cu, probe_spanpipe.cu) show what bounds the chunked kernels
// above: writing a [M x 399] matrix in row PIECES (any width below the row) runs at 2.2 TB/s read+write, because every
// piece ends inside a 32-byte sector of a 1596-byte-pitch row, while moving the same matrix as contiguous spans through
// shared memory with cp.async.bulk runs at 4.8 - 6.5 TB/s.  So here all N <= 400 accumulator columns of a 128-row block
// stay in tensor memory (5 chunks of 80 columns, single buffered, one CTA per SM) and the epilogue is row major:
//
//   warp 10  pulls the base rows of the block into a ring of four 16-row SPAN slots (one 1-D bulk copy of 16 N floats
//            each: contiguous and 16-byte aligned for any N),
//   warps 2-9 (two per TMEM lane quadrant, lane = row) read their 32 rows x 200 columns of the accumulator and update
//            the span in place (pitch N words: conflict free for odd N),
//   warp 11  pushes the finished span back with one bulk store.
//
// The A tile of the next block is fetched with cp.async under the epilogue.  The main loop (5 chunks x 4 double stages of
// the weight image, ~11.6 us per block) commits one barrier per accumulator chunk, so the two quadrants whose spans were
// prefetched update them chunk by chunk under the MMAs; the other two follow when the first round has been stored (block
// period 20.7 us, traced).  TMEM is full, so the main loop of block i+1 cannot overlap the epilogue of block i.  A ragged
// last group (< 16 rows) is updated directly in global memory.
namespace k128r {
using namespace chain;
#ifdef K128_TRACE
__device__ long long g_k128r_trace[128];
#define RT(i) do { if (blockIdx.x == 0 && ehf == 0 && lane == 0 && it == 2) g_k128r_trace[8 * eq + (i)] = clock64(); } while (0)
#else
#define RT(i) do { } while (0)
#endif
using k128::NKB;

constexpr int GR = 16, NSLOT = 4;
static_assert(NSLOT == 4 && TM / GR == 2 * NSLOT, "the span pipeline is written for two rounds of four 16-row slots per block");
constexpr int ALO = 448, TMEM_ALL = 512;
constexpr int WARP_LOAD = 10, WARP_STORE = 11, NTHREADS = 12 * 32;
// Shape variants.  The span slots take 4 x 16 x MAXN floats; what is left next to the A tile is the weight ring:
//   N <= 400: 80-column chunks, ring of 2 slots x 2 K blocks;  N <= 440: 88-column chunks, ring of 3 slots x 1 K block.
template <int NCH_, int KPS_, int RING_, int MAXN_>
struct Cfg {
  static constexpr int NCH = NCH_, KPS = KPS_, RING = RING_, MAXN = MAXN_;
  static constexpr int KSTAGE = stage_bytes(NCH_);                                      // bytes per K block of 16
  static constexpr int T_OFF = 0, B_OFF = t_bytes_of(TM), S_OFF = B_OFF + RING_ * KPS_ * KSTAGE;
  static constexpr int BIAS_OFF = S_OFF + NSLOT * GR * MAXN_ * 4;
  static constexpr int SMEM = BIAS_OFF + MAXN_ * 4;                                     // one CTA per SM
  static_assert(5 * NCH_ <= ALO && NCH_ % 8 == 0 && 5 * NCH_ >= MAXN_, "five accumulator chunks next to the residual operand");
  static_assert(S_OFF % 16 == 0 && SMEM <= 232448, "span slots must be 16-byte aligned and fit");
};
using CfgA = Cfg<k128::NCH_A, 2, 2, k128::MAXN_A>;     // 221888 bytes
using CfgB = Cfg<k128::NCH_B, 1, 3, k128::MAXN_B>;     // 223168 bytes

// 8 TMEM columns starting at `col` of lane quadrant `eq` -> r[]
__device__ __forceinline__ void tmem_ld8(uint32_t tmem_base, int eq, uint32_t col, uint32_t (&r)[8]) {
  const uint32_t taddr = tmem_base + ((uint32_t)(32 * eq) << 16) + col;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}

__device__ __forceinline__ void tmem_ld4(uint32_t tmem_base, int eq, uint32_t col, uint32_t (&r)[4]) {
  const uint32_t taddr = tmem_base + ((uint32_t)(32 * eq) << 16) + col;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr));
}

// W accumulator columns [c0, c0 + W) of this thread's row: span slot in place (whole 16-row group) or global memory.
// has_base = false: C = scale * acc + bias (the slot holds no base rows and is written, not updated).
template <int W>
__device__ __forceinline__ void span_update(float* my_row, const float* biasS, int c0, const uint32_t (&r)[W], int N,
                                            float base_scale, float scale, bool whole, bool valid, const float* brow,
                                            float* crow, bool has_base) {
  if (whole) {
    if (c0 + W <= N) {
      // all W base values first (W independent shared-memory loads in flight), then the stores: updating four columns at
      // a time made every group wait for the stores of the one before it (the compiler cannot prove biasS and my_row
      // distinct) behind a branch on has_base per group
      float o[W];
      if (has_base) {
#pragma unroll
        for (int j = 0; j < W; ++j) o[j] = base_scale * my_row[c0 + j];
      } else {
#pragma unroll
        for (int j = 0; j < W; ++j) o[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < W; j += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(biasS + c0 + j);   // same address in every lane: broadcast
        o[j] = o[j] + scale * __uint_as_float(r[j]) + b4.x;
        o[j + 1] = o[j + 1] + scale * __uint_as_float(r[j + 1]) + b4.y;
        o[j + 2] = o[j + 2] + scale * __uint_as_float(r[j + 2]) + b4.z;
        o[j + 3] = o[j + 3] + scale * __uint_as_float(r[j + 3]) + b4.w;
      }
#pragma unroll
      for (int j = 0; j < W; ++j) my_row[c0 + j] = o[j];
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j)
        if (c0 + j < N) my_row[c0 + j] = (has_base ? base_scale * my_row[c0 + j] : 0.f) + scale * __uint_as_float(r[j]) + biasS[c0 + j];
    }
  } else if (valid) {
#pragma unroll
    for (int j = 0; j < W; ++j)
      if (c0 + j < N) crow[c0 + j] = (has_base ? base_scale * __ldg(brow + c0 + j) : 0.f) + scale * __uint_as_float(r[j]) + biasS[c0 + j];
  }
}

__device__ __forceinline__ void bulk_store_1d(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

template <class CFG>
__global__ void __launch_bounds__(NTHREADS, 1) k_gemm_k128_rows(const k128::Args a) {
  constexpr int NCH = CFG::NCH, KPS = CFG::KPS, RING = CFG::RING, MAXN = CFG::MAXN, KSTAGE = CFG::KSTAGE;
  constexpr int T_OFF = CFG::T_OFF, B_OFF = CFG::B_OFF, S_OFF = CFG::S_OFF, BIAS_OFF = CFG::BIAS_OFF;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_b_full[RING];
  __shared__ __align__(8) uint64_t bar_b_empty[RING];
  __shared__ __align__(8) uint64_t bar_a_ready, bar_chunk_full[5];   // accumulator chunk c complete
  __shared__ __align__(8) uint64_t bar_full[NSLOT], bar_done[NSLOT], bar_free[NSLOT];
  __shared__ uint32_t tmem_holder;
  __shared__ int dead_flag;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  int* const status = a.status;
  volatile int* dead = &dead_flag;
  uint8_t* const T = smem + T_OFF;
  const uint32_t smem_base = smem_u32(smem);
  const int64_t M = a.M;
  const int N = a.N;
  const int64_t m_tiles = (M + TM - 1) / TM;
  const int n_chunks = a.n_chunks;
  constexpr int lbo_t = lbo_t_of(TM);
  const uint32_t slot_bytes = (uint32_t)(GR * N * 4);

  if (tid == 0) {
    dead_flag = 0;
    for (int s = 0; s < RING; ++s) { mbar_init(smem_u32(&bar_b_full[s]), 1); mbar_init(smem_u32(&bar_b_empty[s]), 1); }
    mbar_init(smem_u32(&bar_a_ready), WORKERS / 32);
    for (int c = 0; c < 5; ++c) mbar_init(smem_u32(&bar_chunk_full[c]), 1);
    for (int s = 0; s < NSLOT; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_done[s]), 2); mbar_init(smem_u32(&bar_free[s]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"((uint32_t)TMEM_ALL));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  {
    float* biasS = reinterpret_cast<float*>(smem + BIAS_OFF);
    for (int i = tid; i < MAXN; i += NTHREADS) biasS[i] = (a.bias && i < N) ? a.scale * a.bias_scale * __ldg(a.bias + i) : 0.f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_holder;

  if (warp == 0) {
    // =========================== weight-image producer ===========================
    {
      uint32_t s = 0, ph = 0;
      bool first_lap = true;
      const uint8_t* img = reinterpret_cast<const uint8_t*>(a.img);
      for (int64_t t = blockIdx.x; t < m_tiles; t += gridDim.x) {
        for (int c = 0; c < n_chunks; ++c) {
          for (int kb = 0; kb < NKB; kb += KPS) {     // one bulk copy per KPS K blocks (they are adjacent in the image)
            if (!first_lap) wait_bar(smem_u32(&bar_b_empty[s]), ph ^ 1u, dead, status, 51);
            if (elect_one()) {
              const uint32_t bar = smem_u32(&bar_b_full[s]);
              mbar_expect_tx(bar, KPS * KSTAGE);
              bulk_load_1d(smem_base + B_OFF + s * (KPS * KSTAGE), img + ((size_t)c * NKB + kb) * KSTAGE, KPS * KSTAGE, bar);
            }
            __syncwarp();
            if (++s == (uint32_t)RING) { s = 0; ph ^= 1u; first_lap = false; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    uint32_t sb = 0, pb = 0, it = 0;
    for (int64_t t = blockIdx.x; t < m_tiles; t += gridDim.x, ++it) {
      // A tile + residual operand in place; the workers arrive after draining the previous block's accumulators
      wait_bar(smem_u32(&bar_a_ready), it & 1u, dead, status, 53);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int c = 0; c < n_chunks; ++c) {
        for (int kb = 0; kb < NKB; kb += KPS) {
          wait_bar(smem_u32(&bar_b_full[sb]), pb, dead, status, 52);
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < KPS; ++j)
              issue_kblock(tmem_base + (uint32_t)(NCH * c), tmem_base + ALO, smem_base + T_OFF, (uint32_t)lbo_t,
                           smem_base + B_OFF + (sb * KPS + j) * KSTAGE, NCH, kb + j, kb + j == 0);
            umma_commit(smem_u32(&bar_b_empty[sb]));
            if (kb + KPS == NKB) umma_commit(smem_u32(&bar_chunk_full[c]));
          }
          __syncwarp();
          if (++sb == (uint32_t)RING) { sb = 0; pb ^= 1u; }
        }
      }
    }
  } else if (warp == WARP_LOAD) {
    // =========================== base spans -> slots ===========================
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t t = blockIdx.x; t < m_tiles; t += gridDim.x, ++it) {
        for (int g = 0; g < TM / GR; ++g) {
          const uint32_t s = (uint32_t)g % NSLOT, lap = 2u * it + (uint32_t)g / NSLOT;
          if (lap > 0) wait_bar(smem_u32(&bar_free[s]), (lap - 1) & 1u, dead, status, 54);
          const int64_t r0 = t * TM + (int64_t)g * GR;
          const bool whole = r0 + GR <= M;
          const uint32_t bar = smem_u32(&bar_full[s]);
          const bool fetch = whole && a.base != nullptr;       // without a base term the slot is only handed over
          mbar_expect_tx(bar, fetch ? slot_bytes : 0u);
          if (fetch) bulk_load_1d(smem_base + S_OFF + s * slot_bytes, a.base + r0 * N, slot_bytes, bar);
        }
      }
    }
  } else if (warp == WARP_STORE) {
    // =========================== finished spans -> C ===========================
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t t = blockIdx.x; t < m_tiles; t += gridDim.x, ++it) {
        for (int round = 0; round < TM / GR / NSLOT; ++round) {
          // the four slots of a round: push each as soon as its rows are final, then free them in order as soon as the
          // stores have READ them (about a microsecond) -- the loader is waiting for the slots.  (Freeing a slot one
          // store late serialised the four quadrants of a block: traced at 40 us per block instead of 28.)
          const uint32_t lap = 2u * it + (uint32_t)round;
#pragma unroll
          for (int s = 0; s < NSLOT; ++s) {
            wait_bar(smem_u32(&bar_done[s]), lap & 1u, dead, status, 55);
            const int64_t r0 = t * TM + (int64_t)(round * NSLOT + s) * GR;
            if (r0 + GR <= M) bulk_store_1d(a.C + r0 * N, smem_base + S_OFF + s * slot_bytes, slot_bytes);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory"); mbar_arrive(smem_u32(&bar_free[0]));
          asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); mbar_arrive(smem_u32(&bar_free[1]));
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); mbar_arrive(smem_u32(&bar_free[2]));
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); mbar_arrive(smem_u32(&bar_free[3]));
        }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else {
    // =========================== workers ===========================
    const int wt = tid - 64;
    const int cw = warp - 2;
    const int eq = warp & 3, ehf = cw >> 2;            // TMEM lane quadrant; which half of the columns
    const float scale = a.scale, base_scale = a.base_scale;
    const float* const biasS = reinterpret_cast<const float*>(smem + BIAS_OFF);
    const uint32_t s0 = (uint32_t)(2 * eq) % NSLOT;
    float* const my_row = reinterpret_cast<float*>(smem + S_OFF) + (size_t)(s0 * GR + lane) * N;   // slots s0, s0+1 are adjacent

    auto fetch_tile = [&](int64_t t_) {
      const int64_t m0_ = t_ * TM;
      const int nr_ = (int)((M - m0_ < TM) ? (M - m0_) : TM);
#pragma unroll 4
      for (int idx = wt; idx < TM * NCHUNK; idx += WORKERS) {
        const int r = idx >> 5, c4 = idx & 31;
        const uint32_t dst = smem_base + T_OFF + (uint32_t)c4 * lbo_t + (uint32_t)r * 16u;
        if (r < nr_) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(a.A + (size_t)(m0_ + r) * W2H + 4 * c4) : "memory");
        } else {
          *reinterpret_cast<float4*>(T + (size_t)c4 * lbo_t + r * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto publish_tile = [&]() {
      asm volatile("cp.async.wait_all;" ::: "memory");
      worker_sync_w();                                   // every granule landed; every worker is past its TMEM reads
      residual_to_tmem(T, lbo_t, TM, tmem_base, eq, lane, 16 * ehf, 0, ALO);
      residual_to_tmem(T, lbo_t, TM, tmem_base, eq, lane, 16 * ehf + 8, 0, ALO);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_a_ready));
    };

    if ((int64_t)blockIdx.x < m_tiles) { fetch_tile(blockIdx.x); publish_tile(); }
    uint32_t it = 0;
    for (int64_t t = blockIdx.x; t < m_tiles; t += gridDim.x, ++it) {
      const int64_t m0 = t * TM;
      RT(0);
      const bool more = t + gridDim.x < m_tiles;
      const uint32_t par = it & 1u;
      // Quadrants 0 / 1 have their spans waiting (prefetched under the main loop) and update them chunk by chunk as the
      // accumulator chunks complete -- four fifths of their update runs under the MMAs.  Quadrants 2 / 3 get their slots
      // only after the first round has been stored: they start the next block's A tile first (the tile is free once the
      // last chunk is complete).
      if (eq >= 2) {
        wait_bar(smem_u32(&bar_chunk_full[n_chunks - 1]), par, dead, status, 56);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        RT(1);
        if (more) fetch_tile(t + gridDim.x);
      }
      RT(2);
      const uint32_t lap = 2u * it + (uint32_t)(eq >> 1);
      wait_bar(smem_u32(&bar_full[s0]), lap & 1u, dead, status, 57);
      wait_bar(smem_u32(&bar_full[s0 + 1]), lap & 1u, dead, status, 57);
      RT(3);
      const int64_t grow = m0 + 32 * eq + lane;
      const bool valid = grow < M;
      const bool whole = m0 + 32 * eq + (lane & 16) + GR <= M;    // my 16-row group went through the slot
      const bool has_base = a.base != nullptr;
      const float* brow = has_base ? a.base + grow * a.ldbase : nullptr;
      float* crow = a.C + grow * a.ldc;
      for (int c = 0; c < n_chunks; ++c) {
        if (eq < 2) {
          wait_bar(smem_u32(&bar_chunk_full[c]), par, dead, status, 56);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (c == n_chunks - 1) RT(1);
        }
        const int c0 = NCH * c + (NCH / 2) * ehf;       // my 40 (44) columns of the chunk
        {
          // every accumulator load of the chunk is issued before the one wait
          uint32_t r[32], r8[8], r4[4];
          tmem_ld32(tmem_base, eq, (uint32_t)c0, r);
          tmem_ld8(tmem_base, eq, (uint32_t)(c0 + 32), r8);
          if (NCH / 2 > 40) tmem_ld4(tmem_base, eq, (uint32_t)(c0 + 40), r4);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          span_update<32>(my_row, biasS, c0, r, N, base_scale, scale, whole, valid, brow, crow, has_base);
          span_update<8>(my_row, biasS, c0 + 32, r8, N, base_scale, scale, whole, valid, brow, crow, has_base);
          if (NCH / 2 > 40) span_update<4>(my_row, biasS, c0 + 40, r4, N, base_scale, scale, whole, valid, brow, crow, has_base);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) { mbar_arrive(smem_u32(&bar_done[s0])); mbar_arrive(smem_u32(&bar_done[s0 + 1])); }
      RT(4);
      if (more && eq < 2) fetch_tile(t + gridDim.x);
      if (more) publish_tile();
      RT(5);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_ALL));
  }
}

}  // namespace k128r

namespace tc { int* status_ptr(); }
#ifdef K128_TRACE
extern "C" int gnode_k128r_trace(long long* out128) {
  return (int)cudaMemcpyFromSymbol(out128, k128r::g_k128r_trace, sizeof(long long) * 128);
}
#endif

size_t gemm_k128_image_floats(int n) {
  const int nch = k128::nch_of(n);
  return (size_t)((n + nch - 1) / nch) * k128::NKB * chain::stage_bytes(nch) / 4;
}

int gemm_k128_pack(const float* W, int n, int64_t ld, float* img, cudaStream_t s) {
  const int nch = k128::nch_of(n);
  const int n_chunks = (n + nch - 1) / nch;
  const int64_t total = (int64_t)n_chunks * k128::NKB * 10 * (nch + 1);
  k128::k_pack_k128<<<(unsigned)ceil_div64(total, 256), 256, 0, s>>>(W, n, ld, reinterpret_cast<uint4*>(img), n_chunks, nch);
  GN_LAUNCHED();
  return GNODE_OK;
}

bool gemm_k128_supported(const GemmNT& g) {
  return g.Bchain != nullptr && g.K == chain::W2H && g.lda == chain::W2H && g.relu == 0 && g.post_relu == 0 && g.M > 0 &&
         g.N >= 16 && (reinterpret_cast<uintptr_t>(g.A) & 15) == 0;
}

// dense rows on both sides, at most one base term, N <= 440: the row-major kernel
bool gemm_k128_rows_supported(const GemmNT& g) {
  return gemm_k128_supported(g) && g.base2 == nullptr && g.N <= k128::MAXN_B && g.ldc == g.N &&
         (g.base == nullptr || (g.ldbase == g.N && (reinterpret_cast<uintptr_t>(g.base) & 15) == 0)) &&
         (reinterpret_cast<uintptr_t>(g.C) & 15) == 0;
}

static void k128_args(const GemmNT& g, int* status_dev, k128::Args& a) {
  a.A = g.A; a.img = g.Bchain; a.C = g.C; a.ldc = g.ldc; a.M = g.M; a.N = g.N;
  a.n_chunks = (g.N + k128::nch_of(g.N) - 1) / k128::nch_of(g.N);
  a.bias = g.bias; a.bias_scale = g.bias_scale;
  a.base = g.base; a.ldbase = g.ldbase; a.base_scale = g.base_scale;
  a.base2 = g.base ? g.base2 : nullptr; a.ldbase2 = g.ldbase2;
  a.scale = g.scale;
  a.status = status_dev;
  a.dev_flags = 0;
#ifdef K128_TRACE
  { const char* e = std::getenv("K128_DEV_FLAGS"); a.dev_flags = e ? atoi(e) : 0; }
#endif
}

int gemm_k128_rows(const GemmNT& g, cudaStream_t s) {
  int* status_dev = tc::status_ptr();
  if (!status_dev) { set_error("gemm_k128_rows: status symbol unavailable"); return GNODE_ERR_CUDA; }
  k128::Args a{};
  k128_args(g, status_dev, a);
  const int64_t m_tiles = (g.M + chain::TM - 1) / chain::TM;
  unsigned grid = (unsigned)(m_tiles < kNumSMs ? m_tiles : kNumSMs);
#ifdef K128_TRACE
  { const char* e = std::getenv("K128_GRID"); if (e && atoi(e) > 0 && (unsigned)atoi(e) < grid) grid = (unsigned)atoi(e); }
#endif
  if (g.N <= k128::MAXN_A) {
    if (first_use_on_device(reinterpret_cast<const void*>(&k128r::k_gemm_k128_rows<k128r::CfgA>))) {
      GN_CUDA(cudaFuncSetAttribute(k128r::k_gemm_k128_rows<k128r::CfgA>, cudaFuncAttributeMaxDynamicSharedMemorySize, k128r::CfgA::SMEM));
    }
    k128r::k_gemm_k128_rows<k128r::CfgA><<<grid, k128r::NTHREADS, k128r::CfgA::SMEM, s>>>(a);
  } else {
    if (first_use_on_device(reinterpret_cast<const void*>(&k128r::k_gemm_k128_rows<k128r::CfgB>))) {
      GN_CUDA(cudaFuncSetAttribute(k128r::k_gemm_k128_rows<k128r::CfgB>, cudaFuncAttributeMaxDynamicSharedMemorySize, k128r::CfgB::SMEM));
    }
    k128r::k_gemm_k128_rows<k128r::CfgB><<<grid, k128r::NTHREADS, k128r::CfgB::SMEM, s>>>(a);
  }
  GN_LAUNCHED();
  return GNODE_OK;
}

}  // namespace gnode
