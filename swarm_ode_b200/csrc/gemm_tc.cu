// tcgen05 (5th-gen tensor core) engine for the dense contractions of the GNODE path, sm_100a only.
//
//   C[m, n] = base[m, n] + scale * act( sum_k A[m, k] * B[n, k] + bias[n] )
//     A: fp32 activations [M, K], row stride lda (rows need NOT be 16-byte aligned: D = 399)
//     B: fp32 weights, pre-packed once per call into per-stage shared-memory images of tf32 hi/lo planes
//
// fp32-grade accuracy on the tensor pipe by the 3xTF32 split: a = a_hi + a_lo, a_hi = rna_tf32(a),
// a_lo = rna_tf32(a - a_hi); products a_lo*b_hi + a_hi*b_lo + a_hi*b_hi accumulate in fp32 in TMEM.
//
// Persistent kernel, one CTA per SM, eleven warps, three decoupled shared-memory rings:
//   warp 0      A producer: RAW fp32 tiles through cp.async.bulk.tensor into a DEEP ring (up to 10 stages,
//               ~100 KB in flight per SM -- what Little's law asks for at HBM latency).  Rows that are not
//               16-byte aligned are viewed as "super-rows" of 4 rows (pitch = multiple of 16 B); four
//               16-byte-aligned boxes per stage pick the four row phases, the converter applies the shift.
//   warp 10     B producer: one cp.async.bulk per K-block of the pre-packed weight image (already in UMMA
//               layout) into its own ring.
//   warps 2-5   converters: raw ring -> registers -> hi/lo split -> conflict-free st.shared into the UMMA
//               K-major no-swizzle A-operand ring -> fence.proxy.async -> mbarrier arrive.
//   warp 1      TMEM alloc; one lane issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) and the tcgen05.commit
//               that release the A-operand / B stages and publish the accumulator.
//   warps 6-9   epilogue: tcgen05.ld of their TMEM lane quadrant -> smem transpose -> coalesced global
//               stores with bias / activation / scale / base.  The accumulator is double buffered in
//               TMEM, so the epilogue of tile i overlaps the main loop of tile i+1.
//
// Every mbarrier wait is bounded: on timeout the kernel records a status word and finishes instead of
// hanging (gnode_tc_status() turns that into an error).
#include <cuda.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace gnode {
int gemm_nt_simt(const GemmNT& g, cudaStream_t s);

namespace tc {

constexpr int BM = 128;
constexpr int BK = 16;                 // K elements per stage: 4 chunks of 16 bytes, 2 MMA K-steps
constexpr int CHUNKS = BK / 4;
constexpr int LBO_A = BM * 16 + 16;    // bytes between consecutive 16-byte K-chunks of the A planes (+16: bank spread)
constexpr int A_PLANE = CHUNKS * LBO_A;                 // 8256
constexpr int AOP_BYTES = 2 * A_PLANE;                  // hi + lo planes of one A-operand stage
// Raw A stage.  Aligned rows (J = 1): one box of 128 rows x 16 floats.  Unaligned rows (J = 4): TMA box
// starts must be 16-byte aligned in global memory, so each of the four row-phase boxes starts at the aligned
// address below its first element and is 20 floats wide; the converter applies the 0..3 element shift.
__host__ __device__ constexpr int raw_stride(int J) { return J == 4 ? 20 : 16; }       // floats per raw row
__host__ __device__ constexpr int raw_bytes(int J) { return BM * raw_stride(J) * 4; }  // 8192 / 10240
constexpr int STAGING_BYTES = 4 * 32 * 33 * 4;          // 16896
constexpr int CONV_WARPS = 8;                           // converter warps (2 per SM sub-partition)
// The converter chain of a stage (wait raw -> loads, split -> wait operand slot -> stores -> proxy fence -> arrive) is a serial
// latency per stage however little work it holds: the eight warps form CONV_GROUPS groups of four that take alternate stages
// (each group converts whole stages), so two chains are in flight per CTA (same finding as in gemm_tn_tc.cu).
// Measured (profiles/r2_ab_tc.txt): K = 399, N = 128 (Z_0: main-loop bound) 0.42 -> 0.36 ms with two groups; N = 399, K = 128
// (epilogue bound) 4 - 7 % slower: the launcher picks two groups when K >= N, one otherwise.
constexpr int CONV_THREADS = CONV_WARPS * 32;
constexpr int WARP_EPI0 = 2 + CONV_WARPS;               // first epilogue warp (10: 10 % 4 == 2, quadrants 2,3,0,1)
constexpr int WARP_BPROD = WARP_EPI0 + 4;               // B producer warp
constexpr int THREADS = (WARP_BPROD + 1) * 32;          // 480
constexpr int MAX_RAW = 10, MAX_AOP = 4, MAX_B = 4;
#ifdef TC_TRACE
__device__ long long g_tc_trace[16];
#define TW(idx, stmt) do { const long long t0_ = clock64(); stmt; if (blockIdx.x == 0 && lane == 0) g_tc_trace[idx] += clock64() - t0_; } while (0)
#else
#define TW(idx, stmt) do { stmt; } while (0)
#endif
struct Args {
  const float* Bimg;   // pre-packed weight images: [n_tiles][nkb][hi plane | lo plane], each plane CHUNKS * lbo_b bytes
  float* C; int64_t ldc;
  int64_t M; int N; int K;
  int lda;             // row stride of A in elements
  int J;               // rows per TMA super-row: 1 (16-byte aligned rows) or 4
  int bn;              // N-tile width (multiple of 16, <= 256)
  int n_tiles;         // ceil(Npad16 / bn)
  int n_raw, n_aop, n_b;  // ring depths
  int64_t m_tiles;
  const float* bias; int act;
  const float* base; int64_t ldbase;
  float scale;
  float bias_scale, base_scale;
  const float* base2; int64_t ldbase2;
  int post_relu;
  int* status;
};

// HAS_BASE: 0 none, 1 base, 2 base + base2.  All 32 (64) base loads of the chunk are issued before the first use,
// so the epilogue pays the L2 / HBM latency once per 32 x 32 chunk instead of once per 8 rows.
template <int ACT, int HAS_BASE>
__device__ __forceinline__ void store_rows(const float* __restrict__ stg, int lane, float bv, float scale, float* crow,
                                           int64_t ldc, const float* brow, int64_t ldbase, float base_scale,
                                           const float* b2row, int64_t ldbase2, int rmax, bool cin, bool post_relu) {
  float bs[32];
  if (HAS_BASE >= 1) {
#pragma unroll
    for (int u = 0; u < 32; ++u) bs[u] = (cin && u < rmax) ? __ldg(brow + (int64_t)u * ldbase) : 0.f;
  }
  float b2[32];
  if (HAS_BASE == 2) {
#pragma unroll
    for (int u = 0; u < 32; ++u) b2[u] = (cin && u < rmax) ? __ldg(b2row + (int64_t)u * ldbase2) : 0.f;
  }
#pragma unroll
  for (int u = 0; u < 32; ++u) {
    float x = stg[u * 33 + lane] + bv;
    if (ACT == 1) x = fmaxf(x, 0.f);
    if (ACT == 2) x = tanhf(x);
    x *= scale;
    if (HAS_BASE >= 1) x += base_scale * bs[u];
    if (HAS_BASE == 2) x += b2[u];
    if (post_relu) x = fmaxf(x, 0.f);
    if (cin && u < rmax) crow[(int64_t)u * ldc] = x;
  }
}

template <int CONV_GROUPS>
__global__ void __launch_bounds__(THREADS, 1) k_gemm_tc(const __grid_constant__ CUtensorMap tmapA, const Args a) {
  constexpr int GROUP_WARPS = CONV_WARPS / CONV_GROUPS;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_raw_full[MAX_RAW];
  __shared__ __align__(8) uint64_t bar_raw_empty[MAX_RAW];
  __shared__ __align__(8) uint64_t bar_aop_full[MAX_AOP];
  __shared__ __align__(8) uint64_t bar_b_full[MAX_B];
  __shared__ __align__(8) uint64_t bar_b_empty[MAX_B];
  __shared__ __align__(8) uint64_t bar_acc_full[2];
  __shared__ __align__(8) uint64_t bar_acc_empty[2];
  __shared__ uint32_t tmem_holder;

#ifdef TC_TRACE
  const long long t_kernel0 = clock64();
#endif
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  // hoist every parameter the role loops need (keeps constant-bank loads out of the hot loops)
  const int K = a.K, J = a.J, lda = a.lda, bn = a.bn, n_tiles = a.n_tiles;
  const int NR = a.n_raw, NA = a.n_aop, NB = a.n_b;
  int* const status = a.status;
  const int nkb = (K + BK - 1) / BK;
  const uint32_t lbo_b = (uint32_t)bn * 16u + 16u;
  const uint32_t b_plane = CHUNKS * lbo_b;
  const uint32_t RAW_BYTES = raw_bytes(J);
  const int RS = raw_stride(J);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t raw_off = 0;
  const uint32_t aop_off = raw_off + (uint32_t)NR * RAW_BYTES;
  const uint32_t b_off = aop_off + (uint32_t)NA * AOP_BYTES;
  const uint32_t staging_off = b_off + (uint32_t)NB * 2u * b_plane;
  const uint32_t acc_cols = ((uint32_t)bn + 31u) & ~31u;       // TMEM columns of one accumulator buffer
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2 * acc_cols) tmem_cols <<= 1;
  const int64_t total_tiles = a.m_tiles * n_tiles;
  const int64_t tile0 = blockIdx.x, tstep = gridDim.x;

  if (tid == 0) {
    // converter barriers count WARPS (lane 0 arrives after __syncwarp): 256 per-thread arrivals on one barrier word are
    // 256 serialised shared-memory atomics per K block and barrier
    for (int s = 0; s < NR; ++s) { mbar_init(smem_u32(&bar_raw_full[s]), 1); mbar_init(smem_u32(&bar_raw_empty[s]), GROUP_WARPS); }
    for (int s = 0; s < NA; ++s) { mbar_init(smem_u32(&bar_aop_full[s]), GROUP_WARPS); }
    for (int s = 0; s < NB; ++s) { mbar_init(smem_u32(&bar_b_full[s]), 1); mbar_init(smem_u32(&bar_b_empty[s]), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&bar_acc_full[b]), 1); mbar_init(smem_u32(&bar_acc_empty[b]), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_holder;

  if (warp == 0) {
    // =========================== A producer (raw fp32 tiles, TMA) ===========================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmapA)) : "memory");
      uint32_t s = 0, ph = 0;        // ring slot and phase bit (flips on wrap)
      bool first_lap = true, ok = true;
      const int box_rows = BM / J;
      const uint32_t box_bytes = RAW_BYTES / (uint32_t)J;
      for (int64_t t = tile0; t < total_tiles && ok; t += tstep) {
        const int c1 = (int)((t / n_tiles) * box_rows);
        for (int kb = 0; kb < nkb && ok; ++kb) {
          if (!first_lap) TW(7, ok = mbar_wait(smem_u32(&bar_raw_empty[s]), ph ^ 1u, status, 1));
          const uint32_t dst = smem_base + raw_off + s * RAW_BYTES;
          const uint32_t bar = smem_u32(&bar_raw_full[s]);
          mbar_expect_tx(bar, RAW_BYTES);
          for (int j = 0; j < J; ++j)   // box start rounded down to a 16-byte boundary of the super-row
            tma_load_2d(dst + (uint32_t)j * box_bytes, &tmapA, (j * lda + kb * BK) & ~3, c1, bar);
          if (++s == (uint32_t)NR) { s = 0; ph ^= 1u; first_lap = false; }
        }
      }
    }
  } else if (warp == WARP_BPROD) {
    // =========================== B producer (pre-packed weight images) ===========================
    if (lane == 0) {
      const float* Bimg = a.Bimg;
      uint32_t s = 0, ph = 0;
      bool first_lap = true, ok = true;
      const uint32_t img_bytes = 2 * b_plane;
      const float* const basep = a.base;
      const float* const base2p = a.base2;
      const int64_t ldbase = a.ldbase, ldbase2 = a.ldbase2, M = a.M;
      const bool base_pf = basep != nullptr && (reinterpret_cast<uint64_t>(basep) & 15) == 0;
      const bool base2_pf = base2p != nullptr && (reinterpret_cast<uint64_t>(base2p) & 15) == 0;
      for (int64_t t = tile0; t < total_tiles && ok; t += tstep) {
        if (base_pf && (t % n_tiles) == 0) {
          // the epilogue of this tile reads base[m0 : m0+128, :]: one contiguous span -> warm it in L2 now,
          // a whole main loop ahead of its use
          const int64_t m0 = (t / n_tiles) * BM;
          const int64_t rows = (M - m0 < BM) ? (M - m0) : BM;
          const uint32_t bytes = (uint32_t)((rows * ldbase * 4) & ~15ll);
          if (bytes > 0)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<uint64_t>(basep + m0 * ldbase)), "r"(bytes) : "memory");
        }
        if (base2_pf && (t % n_tiles) == 0) {
          const int64_t m0 = (t / n_tiles) * BM;
          const int64_t rows = (M - m0 < BM) ? (M - m0) : BM;
          const uint32_t bytes = (uint32_t)((rows * ldbase2 * 4) & ~15ll);
          if (bytes > 0)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<uint64_t>(base2p + m0 * ldbase2)), "r"(bytes) : "memory");
        }
        const float* src = Bimg + (size_t)(t % n_tiles) * nkb * (img_bytes / 4);
        for (int kb = 0; kb < nkb && ok; ++kb, src += img_bytes / 4) {
          if (!first_lap) TW(9, ok = mbar_wait(smem_u32(&bar_b_empty[s]), ph ^ 1u, status, 6));
          const uint32_t bar = smem_u32(&bar_b_full[s]);
          mbar_expect_tx(bar, img_bytes);
          bulk_load_1d(smem_base + b_off + s * img_bytes, src, img_bytes, bar);
          if (++s == (uint32_t)NB) { s = 0; ph ^= 1u; first_lap = false; }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = make_idesc(bn);
    // descriptor templates: everything but the 14-bit start address; K-step j advances the address by 2 chunks
    const uint64_t desc_a = make_desc(0, LBO_A), desc_b = make_desc(0, lbo_b);
    const uint32_t step_a = (2 * LBO_A) >> 4, step_b = (2 * lbo_b) >> 4;
    uint32_t sa = 0, pa = 0, sb = 0, pb = 0, tc = 0;
    bool ok = true;
    for (int64_t t = tile0; t < total_tiles && ok; t += tstep, ++tc) {
      const uint32_t ab = tc & 1, aph = (tc >> 1) & 1;
      if (tc >= 2) TW(2, ok = mbar_wait(smem_u32(&bar_acc_empty[ab]), aph ^ 1u, status, 4));   // epilogue drained this buffer
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tmem_d = tmem_base + ab * acc_cols;
      for (int kb = 0; kb < nkb && ok; ++kb) {
        TW(0, ok = mbar_wait(smem_u32(&bar_b_full[sb]), pb, status, 2));
        TW(1, ok = ok && mbar_wait(smem_u32(&bar_aop_full[sa]), pa, status, 2));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t a_hi = (smem_base + aop_off + sa * AOP_BYTES) >> 4, a_lo = a_hi + (A_PLANE >> 4);
          const uint32_t b_hi = (smem_base + b_off + sb * 2u * b_plane) >> 4, b_lo = b_hi + (b_plane >> 4);
#ifdef TC_TRACE
          const long long tm0 = clock64();
#endif
#pragma unroll
          for (int j = 0; j < BK / 8; ++j) {
            const uint64_t dah = desc_a | (uint64_t)(a_hi + j * step_a);
            const uint64_t dal = desc_a | (uint64_t)(a_lo + j * step_a);
            const uint64_t dbh = desc_b | (uint64_t)(b_hi + j * step_b);
            const uint64_t dbl = desc_b | (uint64_t)(b_lo + j * step_b);
            umma_tf32(tmem_d, dal, dbh, idesc, (kb > 0 || j > 0) ? 1u : 0u);   // small terms first
            umma_tf32(tmem_d, dah, dbl, idesc, 1u);
            umma_tf32(tmem_d, dah, dbh, idesc, 1u);
          }
#ifdef TC_TRACE
          const long long tm1 = clock64();
#endif
          // ONE commit releases the A-operand stage and the weight stage (same ring depth, same slot index; the
          // converters and the B producer both wait on bar_b_empty)
          umma_commit(smem_u32(&bar_b_empty[sb]));
          if (kb == nkb - 1) umma_commit(smem_u32(&bar_acc_full[ab])); // accumulator complete
#ifdef TC_TRACE
          if (blockIdx.x == 0) { g_tc_trace[10] += tm1 - tm0; g_tc_trace[11] += clock64() - tm1; }
#endif
        }
        __syncwarp();
        if (++sa == (uint32_t)NA) { sa = 0; pa ^= 1u; }
        if (++sb == (uint32_t)NB) { sb = 0; pb ^= 1u; }
      }
    }
  } else if (warp < WARP_EPI0) {
    // =========================== converters (raw -> tf32 hi/lo planes) ===========================
    // A warp instruction covers two rows (r, r+4) x 16 k.  Within its group, warp gw owns row phase p = gw & 3 of the 8-row
    // groups q in [QPW * (gw >> 2), QPW * (gw >> 2) + QPW):  row = 8q + p + 4*half.
    constexpr int QPW = 16 / (GROUP_WARPS / 4);  // 8-row groups per warp
    const int cw = warp - 2;
    const int grp = cw / GROUP_WARPS, gw = cw % GROUP_WARPS;
    const int p = gw & 3, q0 = (gw >> 2) * QPW;
    const int half = lane >> 4, kk = lane & 15;
    const uint32_t kc_off = (uint32_t)(kk >> 2) * LBO_A + (uint32_t)(kk & 3) * 4u;
    // per-thread constant source / destination offsets of its rows
    int src_off[QPW];
    uint32_t dst_off[QPW];
#pragma unroll
    for (int q = 0; q < QPW; ++q) {
      const int row = 8 * (q0 + q) + p + 4 * half;
      const int ridx = (J == 4) ? ((row & 3) * 32 + (row >> 2)) : row;   // position of the row inside the raw boxes
      const int shift = (J == 4) ? (((row & 3) * lda) & 3) : 0;          // element shift of its row phase
      src_off[q] = ridx * RS + shift + kk;
      dst_off[q] = kc_off + (uint32_t)row * 16u;
    }
    uint32_t it = 0;
    bool ok = true;
    for (int64_t t = tile0; t < total_tiles && ok; t += tstep) {
      for (int kb = 0; kb < nkb && ok; ++kb) {
        const uint32_t i = it++;
        if ((int)(i % CONV_GROUPS) != grp) continue;          // the other group's stage
        const uint32_t sr = i % (uint32_t)NR, pr = (i / (uint32_t)NR) & 1u;
        const uint32_t sa = i % (uint32_t)NA, pa = (i / (uint32_t)NA) & 1u;
        const bool first_lap_a = i < (uint32_t)NA;
        if (cw == 0) TW(4, ok = mbar_wait(smem_u32(&bar_raw_full[sr]), pr, status, 3)); else ok = mbar_wait(smem_u32(&bar_raw_full[sr]), pr, status, 3);
        const float* raw = reinterpret_cast<const float*>(smem + raw_off + (size_t)sr * RAW_BYTES);
        float v[QPW];
#pragma unroll
        for (int q = 0; q < QPW; ++q) v[q] = raw[src_off[q]];
        if (kb == nkb - 1 && (kb * BK + kk) >= K) {   // K tail: columns beyond K hold the next row's data (or TMA zero fill)
#pragma unroll
          for (int q = 0; q < QPW; ++q) v[q] = 0.f;
        }
        if (!first_lap_a) { if (cw == 0) TW(5, ok = ok && mbar_wait(smem_u32(&bar_b_empty[sa]), pa ^ 1u, status, 7)); else ok = ok && mbar_wait(smem_u32(&bar_b_empty[sa]), pa ^ 1u, status, 7); }
        uint8_t* a_hi = smem + aop_off + (size_t)sa * AOP_BYTES;
        uint8_t* a_lo = a_hi + A_PLANE;
#pragma unroll
        for (int q = 0; q < QPW; ++q) {
          // hi = x rounded to tf32 (round-half-away on the magnitude bits), lo = x - hi exactly; the tensor core
          // reads the top 19 bits of lo.  (Finite inputs assumed: no NaN/Inf special-casing in the hot loop.)
          const uint32_t hb = (__float_as_uint(v[q]) + 0x1000u) & 0xFFFFE000u;
          const float lo = v[q] - __uint_as_float(hb);
          *reinterpret_cast<uint32_t*>(a_hi + dst_off[q]) = hb;
          *reinterpret_cast<float*>(a_lo + dst_off[q]) = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA (async proxy)
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(smem_u32(&bar_raw_empty[sr]));                   // raw stage consumed (values are in registers)
          mbar_arrive(smem_u32(&bar_aop_full[sa]));
        }
      }
    }
  } else {
    // =========================== epilogue ===========================
    const int qd = warp & 3;   // TMEM lane quadrant this warp may access
    float* stg = reinterpret_cast<float*>(smem + staging_off) + qd * (32 * 33);
    float* const Cp = a.C;
    const float* const basep = a.base;
    const float* const base2p = a.base2;
    const float* const biasp = a.bias;
    const int64_t ldc = a.ldc, ldbase = a.ldbase, ldbase2 = a.ldbase2, M = a.M;
    const int N = a.N, act = a.act;
    const float scale = a.scale, bias_scale = a.bias_scale, base_scale = a.base_scale;
    const bool post_relu = a.post_relu != 0;
    uint32_t tc = 0;
    bool ok = true;
    for (int64_t t = tile0; t < total_tiles && ok; t += tstep, ++tc) {
      const int64_t m0 = (t / n_tiles) * BM + 32 * qd;
      const int n0 = (int)(t % n_tiles) * bn;
      const uint32_t ab = tc & 1, aph = (tc >> 1) & 1;
      if (qd == 0) TW(8, ok = mbar_wait(smem_u32(&bar_acc_full[ab]), aph, status, 5)); else ok = mbar_wait(smem_u32(&bar_acc_full[ab]), aph, status, 5);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ab * acc_cols + ((uint32_t)(32 * qd) << 16);
      const int rmax = (M - m0 < 32) ? (int)(M - m0) : 32;   // valid rows of this warp's slab (may be <= 0)
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
            : "r"(taddr + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(r[j]);   // row = lane, col = j
        __syncwarp();
        const int col = n0 + c0 + lane;
        const bool cin = (c0 + lane < bn) && (col < N);
        const float bv = (cin && biasp) ? bias_scale * __ldg(biasp + col) : 0.f;
        float* crow = Cp + m0 * ldc + col;
        const float* brow = basep ? basep + m0 * ldbase + col : nullptr;
        const float* b2row = base2p ? base2p + m0 * ldbase2 + col : nullptr;
#define GN_STORE(ACT_, HB_) store_rows<ACT_, HB_>(stg, lane, bv, scale, crow, ldc, brow, ldbase, base_scale, b2row, ldbase2, rmax, cin, post_relu)
        if (basep && base2p) {
          if (act == 0) GN_STORE(0, 2); else if (act == 1) GN_STORE(1, 2); else GN_STORE(2, 2);
        } else if (basep) {
          if (act == 0) GN_STORE(0, 1); else if (act == 1) GN_STORE(1, 1); else GN_STORE(2, 1);
        } else {
          if (act == 0) GN_STORE(0, 0); else if (act == 1) GN_STORE(1, 0); else GN_STORE(2, 0);
        }
#undef GN_STORE
        __syncwarp();
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(smem_u32(&bar_acc_empty[ab]));
    }
  }

#ifdef TC_TRACE
  if (blockIdx.x == 0 && tid == 0) g_tc_trace[15] += clock64() - t_kernel0;
#endif
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
  }
}

// Pack a row-major weight matrix W [rows, cols] into per-(n-tile, k-block) stage images:
//   image(nt, kb) = [hi plane | lo plane], plane = CHUNKS chunks of lbo_b bytes, element (n, k) of the tile at
//   chunk (k%16)/4, byte n*16 + (k%4)*4.  Out-of-range rows / columns are zero.
__global__ void k_pack_b(const float* __restrict__ W, int rows, int cols, int64_t ld, float* __restrict__ img, int bn,
                         int n_tiles, int nkb) {
  const int plane_f = CHUNKS * (bn * 16 + 16) / 4;
  const int64_t total = (int64_t)n_tiles * nkb * 2 * plane_f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int w = (int)(r % plane_f); r /= plane_f;
    const int plane = (int)(r % 2); r /= 2;
    const int kb = (int)(r % nkb);
    const int nt = (int)(r / nkb);
    const int chunk_f = (bn * 16 + 16) / 4;
    const int kc = w / chunk_f, rem = w % chunk_f;
    const int nl = rem >> 2, e = rem & 3;
    float v = 0.f;
    const int n = nt * bn + nl, k = kb * BK + kc * 4 + e;
    if (nl < bn && n < rows && k < cols) v = W[(int64_t)n * ld + k];
    const uint32_t h = to_tf32(v);
    img[i] = (plane == 0) ? __uint_as_float(h) : __uint_as_float(to_tf32(v - __uint_as_float(h)));
  }
}

inline int pick_bn(int N) {
  const int npad = (N + 15) & ~15;
  const int ntiles = (npad + 255) / 256;
  return ((npad + ntiles - 1) / ntiles + 15) & ~15;
}
inline int pick_ntiles(int N) {
  const int npad = (N + 15) & ~15;
  const int bn = pick_bn(N);
  return (npad + bn - 1) / bn;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

__device__ int g_tc_status_word = 0;   // barrier-timeout status of the tcgen05 kernels (0 = ok)
int* status_ptr() {
  // the symbol has one instance (and one address) per device: cache per device ordinal, not per process
  constexpr int kMaxDev = 64;
  static int* cache[kMaxDev] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  if (dev < 0 || dev >= kMaxDev) {
    int* p = nullptr;
    return cudaGetSymbolAddress(reinterpret_cast<void**>(&p), g_tc_status_word) == cudaSuccess ? p : nullptr;
  }
  if (!cache[dev]) {
    int* p = nullptr;
    if (cudaGetSymbolAddress(reinterpret_cast<void**>(&p), g_tc_status_word) == cudaSuccess) cache[dev] = p;
  }
  return cache[dev];
}

}  // namespace tc

size_t presplit_floats(int rows, int cols) {
  const int bn = tc::pick_bn(rows), nt = tc::pick_ntiles(rows), nkb = (cols + tc::BK - 1) / tc::BK;
  return (size_t)nt * nkb * 2 * (tc::CHUNKS * (bn * 16 + 16) / 4);
}

int presplit_weights(const float* W, int rows, int cols, int64_t ld, float* img, cudaStream_t s) {
  const int bn = tc::pick_bn(rows), nt = tc::pick_ntiles(rows), nkb = (cols + tc::BK - 1) / tc::BK;
  const int64_t total = (int64_t)presplit_floats(rows, cols);
  int64_t blocks = ceil_div64(total, 256);
  if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
  tc::k_pack_b<<<(unsigned)blocks, 256, 0, s>>>(W, rows, cols, ld, img, bn, nt, nkb);
  GN_LAUNCHED();
  return GNODE_OK;
}

// The TMA view needs either 16-byte aligned rows (lda % 4 == 0) or dense rows (K == lda, super-row trick).
bool gemm_nt_tc_supported(const GemmNT& g) {
  if (g.Bsplit == nullptr || g.M < 4 || g.N < 16 || g.K < 1) return false;
  if ((reinterpret_cast<uintptr_t>(g.A) & 15) != 0) return false;
  if (g.lda % 4 == 0) return true;
  return g.K == g.lda && g.lda < (1 << 20);
}

int gemm_nt_tc(const GemmNT& g, cudaStream_t s) {
  if (g.M == 0 || g.N == 0) return GNODE_OK;
  int* status_dev = tc::status_ptr();
  tc::EncodeTiledFn enc = tc::encode_fn();
  if (!status_dev || !enc) { set_error("gemm_nt_tc: driver entry point / status symbol unavailable"); return GNODE_ERR_CUDA; }
  const int J = (g.lda % 4 == 0) ? 1 : 4;
  const int64_t M_tc = (J == 4) ? (g.M / 4) * 4 : g.M;     // rows covered by whole super-rows
  tc::Args a;
  a.Bimg = g.Bsplit; a.C = g.C; a.ldc = g.ldc; a.M = M_tc; a.N = g.N; a.K = g.K; a.lda = (int)g.lda; a.J = J;
  a.bn = tc::pick_bn(g.N); a.n_tiles = tc::pick_ntiles(g.N);
  a.m_tiles = ceil_div64(M_tc, tc::BM);
  a.bias = g.bias; a.act = g.relu; a.base = g.base; a.ldbase = g.ldbase; a.scale = g.scale;
  a.bias_scale = g.bias_scale; a.base_scale = g.base_scale;
  a.base2 = g.base ? g.base2 : nullptr; a.ldbase2 = g.ldbase2; a.post_relu = g.post_relu;
  a.status = status_dev;
  // shared-memory plan: A-operand ring (3) + B ring (4, or 3 for wide tiles) + staging, the rest is the raw ring
  const size_t img = 2 * (size_t)tc::CHUNKS * ((size_t)a.bn * 16 + 16);
  const size_t budget = 224 * 1024;
  a.n_b = (a.bn > 128) ? 3 : tc::MAX_B;
  a.n_aop = a.n_b;      // one ring index and one release barrier for the A-operand stage and the weight stage
  const size_t fixed = (size_t)a.n_aop * tc::AOP_BYTES + (size_t)a.n_b * img + tc::STAGING_BYTES;
  int n_raw = (int)((budget - fixed) / tc::raw_bytes(J));
  if (n_raw > tc::MAX_RAW) n_raw = tc::MAX_RAW;
  if (n_raw < 2) { set_error("gemm_nt_tc: tile does not fit in shared memory"); return GNODE_ERR_ARG; }
  a.n_raw = n_raw;
  const size_t smem = fixed + (size_t)n_raw * tc::raw_bytes(J);

  CUtensorMap tmap;
  {
    cuuint64_t gdim[2], gstride[1];
    cuuint32_t box[2] = {(cuuint32_t)tc::raw_stride(J), (cuuint32_t)(tc::BM / J)};
    cuuint32_t estr[2] = {1, 1};
    if (J == 1) {
      gdim[0] = (cuuint64_t)g.K; gdim[1] = (cuuint64_t)M_tc; gstride[0] = (cuuint64_t)g.lda * 4;
    } else {
      gdim[0] = (cuuint64_t)g.lda * 4; gdim[1] = (cuuint64_t)(M_tc / 4); gstride[0] = (cuuint64_t)g.lda * 16;
    }
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(g.A), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("gemm_nt_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return GNODE_ERR_CUDA; }
  }
  if (first_use_on_device(reinterpret_cast<const void*>(&tc::k_gemm_tc<1>))) {
    GN_CUDA(cudaFuncSetAttribute(tc::k_gemm_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(225 * 1024)));
    GN_CUDA(cudaFuncSetAttribute(tc::k_gemm_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(225 * 1024)));
  }
  if (M_tc > 0) {
    const int64_t total = a.m_tiles * a.n_tiles;
    const unsigned grid = (unsigned)(total < kNumSMs ? total : kNumSMs);
    if (g.K >= g.N) tc::k_gemm_tc<2><<<grid, tc::THREADS, smem, s>>>(tmap, a);   // main-loop bound: two converter chains
    else tc::k_gemm_tc<1><<<grid, tc::THREADS, smem, s>>>(tmap, a);
    GN_LAUNCHED();
  }
  if (M_tc < g.M) {   // up to 3 trailing rows that do not fill a super-row: FFMA kernel
    GemmNT tail = g;
    tail.A = g.A + M_tc * g.lda; tail.C = g.C + M_tc * g.ldc; tail.M = g.M - M_tc;
    if (g.base) tail.base = g.base + M_tc * g.ldbase;
    if (g.base2) tail.base2 = g.base2 + M_tc * g.ldbase2;
    tail.Bsplit = nullptr;
    GN_TRY(gemm_nt_simt(tail, s));
  }
  return GNODE_OK;
}

// reads and clears the tcgen05 status word (0 = ok); synchronises the stream
int gemm_tc_status(cudaStream_t s, int* out) {
  *out = 0;
  int* status_dev = tc::status_ptr();
  if (!status_dev) return GNODE_OK;
  GN_CUDA(cudaMemcpyAsync(out, status_dev, sizeof(int), cudaMemcpyDeviceToHost, s));
  GN_CUDA(cudaStreamSynchronize(s));
  if (*out != 0) GN_CUDA(cudaMemsetAsync(status_dev, 0, sizeof(int), s));
  return GNODE_OK;
}

}  // namespace gnode

// Enqueues (no synchronisation) a copy of the tcgen05 status word of the current device into `host_word` (pinned host
// memory): the deferred counterpart of gnode_tc_status for hot paths that must not block.
extern "C" int gnode_tc_status_async(int32_t* host_word, gnode_stream_t stream) {
  GN_ARG(host_word != nullptr, "gnode_tc_status_async: host_word is null");
  int* status_dev = gnode::tc::status_ptr();
  if (!status_dev) { gnode::set_error("gnode_tc_status_async: status symbol unavailable"); return GNODE_ERR_CUDA; }
  GN_CUDA(cudaMemcpyAsync(host_word, status_dev, sizeof(int), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
  return GNODE_OK;
}

// Synchronises the stream and reports whether any tcgen05 kernel hit a barrier timeout since the last
// call (0 = healthy).  Meant for tests / debugging; the hot path never calls it.
extern "C" int gnode_tc_status(gnode_stream_t stream) {
  int st = 0;
  int rc = gnode::gemm_tc_status(static_cast<cudaStream_t>(stream), &st);
  if (rc != GNODE_OK) return rc;
  if (st != 0) {
    gnode::set_error("tcgen05 kernel barrier timeout (code %d: 1 = A producer waiting for a free raw stage, 2 = MMA waiting "
                     "for operands, 3 = converter waiting for TMA bytes, 4 = MMA waiting for a free accumulator, "
                     "5 = epilogue waiting for the accumulator, 6 = B producer waiting for a free stage, 7 = converter "
                     "waiting for a free A-operand stage)", st);
    return GNODE_ERR_CUDA;
  }
  return GNODE_OK;
}

#ifdef TC_TRACE
extern "C" int gnode_tc_trace(long long* out16, int reset) {
  int rc = (int)cudaMemcpyFromSymbol(out16, gnode::tc::g_tc_trace, sizeof(long long) * 16);
  if (reset) { long long z[16] = {0}; rc |= (int)cudaMemcpyToSymbol(gnode::tc::g_tc_trace, z, sizeof(z)); }
  return rc;
}
#endif
