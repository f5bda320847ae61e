// Graph-resident forward chain of the folded RK stages (sm_100a, tcgen05).
//
// After folding (fold.cu) the per-stage work of GraphODEFunc (scripts/train_gde.py:33-45) is 2H-wide only:
//
//   V_s  = dt * sum_{j<s} beta_sj cat2_j                Z_s = Z_0 + V_s @ M13^T + (dt sum_j beta_sj) c13
//   h1   = relu(A(Z_l) + Z_r + b1)                      cat1_s = [A(h1) | h1]
//   h2   = relu(cat1_s @ w2cat^T + b2)                  cat2_s = [A(h2) | h2]
//
// A batch is a disjoint union of small graphs (Batch.from_data_list, scripts/train_gde.py:367): no edge leaves its
// graph, so a tile of <= 128 consecutive rows holding WHOLE graphs carries every neighbour row its mean
// aggregations A(.) need.  One CTA owns a tile for ALL stages of the step: the [128 x 2H] fp32 tile lives in shared
// memory, the two small contractions run on tcgen05 (3xTF32, accumulators in TMEM), the aggregations read shared
// memory, and HBM only sees Z_0 once on the way in and V_s / cat1_s / cat2_s (what the backward pass needs) on the
// way out; the cat2_j re-reads of later stages hit L2 (the same CTA wrote them microseconds earlier).
//
//   warp 0      weight-image producer: M13 / w2cat K-block images (pre-split tf32 hi/lo, UMMA layout) by 1-D bulk copy
//   warp 1      TMEM alloc + MMA issue (kind::tf32, lo*hi + hi*lo + hi*hi per K step)
//   warps 2-9   workers: tile loads / combinations, fp32 -> hi/lo operand conversion, TMEM epilogues, aggregations,
//               tile stores; they advance in lockstep through named barrier 1.
#include "common.cuh"
#include "field.cuh"
#include "tc_common.cuh"

namespace gnode {
namespace chain {
using namespace tc;

constexpr int TM = 128;                 // rows per tile = TMEM lanes
constexpr int W2H = 128, WH = 64;       // 2H, H (this kernel is specialised for hidden_dim = 64)
constexpr int TP = 132;                 // tile row pitch in floats (16-byte aligned rows)
constexpr int BK = 8, CHUNKS = 2, NKB = W2H / BK;  // one tf32 K step per stage: small rings -> two CTAs per SM
constexpr int LBO_A = TM * 16 + 16;
constexpr int A_PLANE = CHUNKS * LBO_A;            // 4128
constexpr int AOP_BYTES = 2 * A_PLANE;             // 8256
// weight images come from presplit_weights (K blocks of 16: [hi plane: 4 chunks | lo plane: 4 chunks]); a stage of
// this kernel takes chunks (2h, 2h+1) of both planes of K-block kb/2
constexpr int LBO_B1 = W2H * 16 + 16, B1_PLANE16 = 4 * LBO_B1, B1_IMG16 = 2 * B1_PLANE16;   // N = 128
constexpr int LBO_B2 = WH * 16 + 16, B2_PLANE16 = 4 * LBO_B2, B2_IMG16 = 2 * B2_PLANE16;    // N = 64
constexpr int B_STAGE = 4 * LBO_B1;                // [hi: 2 chunks | lo: 2 chunks] of the wider image: 8256
constexpr int N_AOP = 2, N_B = 2;
constexpr int WORKERS = 256, THREADS = 64 + WORKERS;
constexpr int T_BYTES = TM * TP * 4;               // 67584
constexpr int SMEM_BYTES = T_BYTES + N_AOP * AOP_BYTES + N_B * B_STAGE;   // 100608: two CTAs per SM
constexpr int NBR_REG = 4;                         // neighbour ids per row kept in registers
constexpr int TMEM_COLS = 256;                     // acc1: cols 0..127, acc2: cols 128..191

struct Args {
  const float* z0;
  float* cat1[kMaxStages];
  float* cat2[kMaxStages];
  float* V[kMaxStages];
  float coef[kMaxStages][kMaxStages];   // dt * beta[s][j]
  float csol[kMaxStages];               // dt * c_sol[s]
  float* Cout;                          // optional: C = sum_s csol[s] cat2_s, written after the last stage
  float c13_scale[kMaxStages];          // dt * sum_j beta[s][j]
  const float *c13, *b1, *b2;
  const float *img13, *img2;            // weight images of M13 [2H x 2H] and w2cat [H x 2H]
  const int32_t *rowptr, *col;
  const int32_t* tiles;                 // [0] = number of tiles, [1 ..] = first row of every tile, then N
  int S;
  int* status;                          // barrier-timeout word shared with the other tcgen05 kernels
  int* err;                             // set to 1 when a neighbour lies outside its tile
};

__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(WORKERS) : "memory"); }

// bounded wait that gives up immediately once any wait of this CTA has timed out
__device__ __forceinline__ void wait_bar(uint32_t addr, uint32_t parity, volatile int* dead, int* status, int code) {
  if (mbar_try_wait(addr, parity)) return;            // fast path: no shared-memory flag read
  for (uint32_t i = 0; i < SPIN_LIMIT; ++i) {
    if (mbar_try_wait(addr, parity)) return;
    if ((i & 1023u) == 1023u && *dead) return;         // another wait of this CTA already timed out
  }
  *dead = 1;
  if (status) atomicExch(status, code);
}

__global__ void __launch_bounds__(THREADS, 2) k_chain_fwd(const Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_aop_full[N_AOP];
  __shared__ __align__(8) uint64_t bar_aop_empty[N_AOP];
  __shared__ __align__(8) uint64_t bar_b_full[N_B];
  __shared__ __align__(8) uint64_t bar_b_empty[N_B];
  __shared__ __align__(8) uint64_t bar_acc_full[2];
  __shared__ uint32_t tmem_holder;
  __shared__ int dead_flag;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int S = a.S;
  int* const status = a.status;
  volatile int* dead = &dead_flag;
  float* const T = reinterpret_cast<float*>(smem);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t aop_off = T_BYTES, b_off = T_BYTES + N_AOP * AOP_BYTES;
  const int n_tiles = a.tiles[0];

  if (tid == 0) {
    dead_flag = 0;
    for (int s = 0; s < N_AOP; ++s) { mbar_init(smem_u32(&bar_aop_full[s]), WORKERS); mbar_init(smem_u32(&bar_aop_empty[s]), 1); }
    for (int s = 0; s < N_B; ++s) { mbar_init(smem_u32(&bar_b_full[s]), 1); mbar_init(smem_u32(&bar_b_empty[s]), 1); }
    mbar_init(smem_u32(&bar_acc_full[0]), 1);
    mbar_init(smem_u32(&bar_acc_full[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"((uint32_t)TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_holder;

  if (warp == 0) {
    // =========================== weight-image producer ===========================
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      bool first_lap = true;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int st = 0; st < S; ++st) {
          for (int g = (st > 0 ? 0 : 1); g < 2; ++g) {           // g = 0: M13 (stage > 0 only), g = 1: w2cat
            const uint8_t* img = reinterpret_cast<const uint8_t*>(g == 0 ? a.img13 : a.img2);
            const uint32_t lbo_b = g == 0 ? LBO_B1 : LBO_B2;
            const uint32_t plane16 = g == 0 ? B1_PLANE16 : B2_PLANE16, img16 = g == 0 ? B1_IMG16 : B2_IMG16;
            const uint32_t half_bytes = 2 * lbo_b;                 // two K chunks of one plane
            for (int kb = 0; kb < NKB; ++kb) {
              if (!first_lap) wait_bar(smem_u32(&bar_b_empty[s]), ph ^ 1u, dead, status, 21);
              const uint32_t bar = smem_u32(&bar_b_full[s]);
              const uint8_t* src = img + (size_t)(kb >> 1) * img16 + (size_t)(kb & 1) * half_bytes;
              const uint32_t dst = smem_base + b_off + s * B_STAGE;
              mbar_expect_tx(bar, 2 * half_bytes);
              bulk_load_1d(dst, src, half_bytes, bar);                           // hi chunks
              bulk_load_1d(dst + half_bytes, src + plane16, half_bytes, bar);    // lo chunks
              if (++s == (uint32_t)N_B) { s = 0; ph ^= 1u; first_lap = false; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    const uint64_t desc_a = make_desc(0, LBO_A);
    uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      for (int st = 0; st < S; ++st) {
        for (int g = (st > 0 ? 0 : 1); g < 2; ++g) {
          const uint32_t lbo_b = g == 0 ? LBO_B1 : LBO_B2;
          const uint32_t idesc = make_idesc(g == 0 ? W2H : WH);
          const uint64_t desc_b = make_desc(0, lbo_b);
          const uint32_t tmem_d = tmem_base + (g == 0 ? 0u : (uint32_t)W2H);
          for (int kb = 0; kb < NKB; ++kb) {
            wait_bar(smem_u32(&bar_b_full[sb]), pb, dead, status, 22);
            wait_bar(smem_u32(&bar_aop_full[sa]), pa, dead, status, 23);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
              const uint32_t a_hi = (smem_base + aop_off + sa * AOP_BYTES) >> 4, a_lo = a_hi + (A_PLANE >> 4);
              const uint32_t b_hi = (smem_base + b_off + sb * B_STAGE) >> 4, b_lo = b_hi + ((2 * lbo_b) >> 4);
              const uint64_t dah = desc_a | (uint64_t)a_hi, dal = desc_a | (uint64_t)a_lo;
              const uint64_t dbh = desc_b | (uint64_t)b_hi, dbl = desc_b | (uint64_t)b_lo;
              umma_tf32(tmem_d, dal, dbh, idesc, kb > 0 ? 1u : 0u);   // small terms first
              umma_tf32(tmem_d, dah, dbl, idesc, 1u);
              umma_tf32(tmem_d, dah, dbh, idesc, 1u);
              umma_commit(smem_u32(&bar_aop_empty[sa]));
              umma_commit(smem_u32(&bar_b_empty[sb]));
              if (kb == NKB - 1) umma_commit(smem_u32(&bar_acc_full[g]));
            }
            __syncwarp();
            if (++sa == (uint32_t)N_AOP) { sa = 0; pa ^= 1u; }
            if (++sb == (uint32_t)N_B) { sb = 0; pb ^= 1u; }
          }
        }
      }
    }
  } else {
    // =========================== workers ===========================
    const int wt = tid - 64;                       // 0..255
    const int cw = warp - 2;                       // 0..7
    // converter mapping: a warp instruction covers 4 rows (two apart: conflict-free banks) x 8 k; warp cw owns rows
    // [16 cw, 16 cw + 16), four instructions per K block
    const int ckk = lane & 7, cr = lane >> 3;
    const uint32_t kc_off = (uint32_t)(ckk >> 2) * LBO_A + (uint32_t)(ckk & 3) * 4u;
    int csrc[4];
    uint32_t cdst[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int row = 16 * cw + (q & 1) + 2 * cr + 8 * (q >> 1);
      csrc[q] = row * TP + ckk;
      cdst[q] = kc_off + (uint32_t)row * 16u;
    }
    uint32_t sa = 0, pa = 0;
    bool first_lap_a = true;
    uint32_t ph_acc[2] = {0u, 0u};
    // epilogue mapping: TMEM lane quadrant of this warp, column half
    const int eq = warp & 3, ehf = cw >> 2;
    // aggregation mapping: two threads per row, 32 of the 64 channels each
    // (threads 0..127 take channels 0..31 of rows 0..127, threads 128..255 channels 32..63: the eight lanes of a
    // 128-bit shared-memory phase then touch eight different rows = eight different bank groups)
    const int arow = wt & (TM - 1), ac0 = (wt >> 7) * 32;

    auto convert_tile = [&]() {
      for (int kb = 0; kb < NKB; ++kb) {
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = T[csrc[q] + kb * BK];
        if (!first_lap_a) wait_bar(smem_u32(&bar_aop_empty[sa]), pa ^ 1u, dead, status, 24);
        uint8_t* a_hi = smem + aop_off + (size_t)sa * AOP_BYTES;
        uint8_t* a_lo = a_hi + A_PLANE;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t hb = (__float_as_uint(v[q]) + 0x1000u) & 0xFFFFE000u;
          *reinterpret_cast<uint32_t*>(a_hi + cdst[q]) = hb;
          *reinterpret_cast<float*>(a_lo + cdst[q]) = v[q] - __uint_as_float(hb);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(smem_u32(&bar_aop_full[sa]));
        if (++sa == (uint32_t)N_AOP) { sa = 0; pa ^= 1u; first_lap_a = false; }
      }
    };
    // 32 TMEM columns starting at `col` of this warp's lane quadrant -> r[]
    auto tmem_ld32 = [&](uint32_t col, uint32_t (&r)[32]) {
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * eq) << 16) + col;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
            "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
            "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    };

    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int r0 = a.tiles[1 + t];
      const int nr = a.tiles[2 + t] - r0;                 // 1..128 rows, whole graphs
      // CSR slice of this thread's row (aggregation phases); the first NBR_REG neighbour ids (tile-local) stay in
      // registers for the three aggregations of every stage
      int nb_b = 0, nb_e = 0;
      if (arow < nr) { nb_b = a.rowptr[r0 + arow]; nb_e = a.rowptr[r0 + arow + 1]; }
      const float inv_deg = 1.0f / (float)((nb_e - nb_b) > 1 ? (nb_e - nb_b) : 1);
      int nbr[NBR_REG];
#pragma unroll
      for (int q = 0; q < NBR_REG; ++q) {
        int v = -1;
        if (nb_b + q < nb_e) {
          v = a.col[nb_b + q] - r0;
          if (v < 0 || v >= nr) { *a.err = 1; v = -1; }       // neighbour outside the tile: not a disjoint-union batch
        }
        nbr[q] = v;
      }

      // mean over the in-neighbours of T[.][src_col0 + ac0 .. +32) of this thread's row
      auto aggregate = [&](int src_col0, float4 (&acc)[8]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < NBR_REG; ++q) {
          if (nbr[q] >= 0) {
            const float4* src = reinterpret_cast<const float4*>(T + nbr[q] * TP + src_col0 + ac0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 v = src[i];
              acc[i].x += v.x; acc[i].y += v.y; acc[i].z += v.z; acc[i].w += v.w;
            }
          }
        }
        for (int p = nb_b + NBR_REG; p < nb_e; ++p) {          // rows with more than NBR_REG neighbours
          const int nb = a.col[p] - r0;
          if (nb < 0 || nb >= nr) { *a.err = 1; continue; }
          const float4* src = reinterpret_cast<const float4*>(T + nb * TP + src_col0 + ac0);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 v = src[i];
            acc[i].x += v.x; acc[i].y += v.y; acc[i].z += v.z; acc[i].w += v.w;
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { acc[i].x *= inv_deg; acc[i].y *= inv_deg; acc[i].z *= inv_deg; acc[i].w *= inv_deg; }
      };
      // coalesced tile store: rows < nr of T -> dst[(r0 + r) * 2H + c]
      auto store_tile = [&](float* dst) {
        for (int idx = wt; idx < TM * (W2H / 4); idx += WORKERS) {
          const int r = idx >> 5, c4 = idx & 31;
          if (r < nr) *reinterpret_cast<float4*>(dst + (size_t)(r0 + r) * W2H + 4 * c4) = *reinterpret_cast<const float4*>(T + r * TP + 4 * c4);
        }
      };

      for (int st = 0; st < S; ++st) {
        // ---- tile input: Z_0 (stage 0) or V_st = sum_j coef * cat2_j (later stages; also written out) ----
        // Every thread owns 16 float4 slots of the tile; all their loads are issued before the first store (the
        // compiler cannot hoist loads over the V stores by itself).
        {
          constexpr int SLOTS = TM * (W2H / 4) / WORKERS / 2;   // 8 per half
          for (int hf = 0; hf < 2; ++hf) {
          const int ibase = wt + hf * SLOTS * WORKERS;
          float4 acc[SLOTS];
#pragma unroll
          for (int u = 0; u < SLOTS; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (st == 0) {
#pragma unroll
            for (int u = 0; u < SLOTS; ++u) {
              const int idx = ibase + u * WORKERS, r = idx >> 5, c4 = idx & 31;
              if (r < nr) acc[u] = __ldg(reinterpret_cast<const float4*>(a.z0 + (size_t)(r0 + r) * W2H + 4 * c4));
            }
          } else {
            for (int j = 0; j < st; ++j) {
              const float cf = a.coef[st][j];
              if (cf == 0.f) continue;
              const float* srcj = a.cat2[j];
              float4 v[SLOTS];
#pragma unroll
              for (int u = 0; u < SLOTS; ++u) {
                const int idx = ibase + u * WORKERS, r = idx >> 5, c4 = idx & 31;
                v[u] = (r < nr) ? *reinterpret_cast<const float4*>(srcj + (size_t)(r0 + r) * W2H + 4 * c4)   // written by this CTA: plain load
                                : make_float4(0.f, 0.f, 0.f, 0.f);
              }
#pragma unroll
              for (int u = 0; u < SLOTS; ++u) {
                acc[u].x = fmaf(cf, v[u].x, acc[u].x); acc[u].y = fmaf(cf, v[u].y, acc[u].y);
                acc[u].z = fmaf(cf, v[u].z, acc[u].z); acc[u].w = fmaf(cf, v[u].w, acc[u].w);
              }
            }
#pragma unroll
            for (int u = 0; u < SLOTS; ++u) {
              const int idx = ibase + u * WORKERS, r = idx >> 5, c4 = idx & 31;
              if (r < nr) *reinterpret_cast<float4*>(a.V[st] + (size_t)(r0 + r) * W2H + 4 * c4) = acc[u];
            }
          }
#pragma unroll
          for (int u = 0; u < SLOTS; ++u) {
            const int idx = ibase + u * WORKERS, r = idx >> 5, c4 = idx & 31;
            *reinterpret_cast<float4*>(T + r * TP + 4 * c4) = acc[u];
          }
          }
        }
        worker_sync();
        if (st > 0) {
          // ---- Z_st = Z_0 + V_st @ M13^T + scale * c13 ----
          convert_tile();
          wait_bar(smem_u32(&bar_acc_full[0]), ph_acc[0], dead, status, 25);
          ph_acc[0] ^= 1u;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          {
            const float cs = a.c13_scale[st];
            const int row = 32 * eq + lane;
            const bool rin = row < nr;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int c0 = 64 * ehf + 32 * h;
              // this lane's Z_0 row segment (32 floats = one 128-byte line) is requested before the TMEM load returns
              float4 z[8];
              const float4* zp = reinterpret_cast<const float4*>(a.z0 + (size_t)(r0 + (rin ? row : 0)) * W2H + c0);
#pragma unroll
              for (int i = 0; i < 8; ++i) z[i] = rin ? __ldg(zp + i) : make_float4(0.f, 0.f, 0.f, 0.f);
              uint32_t r[32];
              tmem_ld32((uint32_t)c0, r);
              float* trow = T + row * TP + c0;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 c = __ldg(reinterpret_cast<const float4*>(a.c13 + c0) + i);
                float4 o;
                o.x = __uint_as_float(r[4 * i + 0]) + z[i].x + cs * c.x;
                o.y = __uint_as_float(r[4 * i + 1]) + z[i].y + cs * c.y;
                o.z = __uint_as_float(r[4 * i + 2]) + z[i].z + cs * c.z;
                o.w = __uint_as_float(r[4 * i + 3]) + z[i].w + cs * c.w;
                *reinterpret_cast<float4*>(trow + 4 * i) = o;
              }
            }
          }
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          worker_sync();
        }
        // ---- h1 = relu(A(Z_l) + Z_r + b1) -> right half (in place) ----
        if (arow < nr) {
          float4 acc[8];
          aggregate(0, acc);
          float4* own = reinterpret_cast<float4*>(T + arow * TP + WH + ac0);
          const float4* bb = reinterpret_cast<const float4*>(a.b1 + ac0);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 z = own[i], b = __ldg(bb + i);
            float4 h;
            h.x = fmaxf(acc[i].x + z.x + b.x, 0.f); h.y = fmaxf(acc[i].y + z.y + b.y, 0.f);
            h.z = fmaxf(acc[i].z + z.z + b.z, 0.f); h.w = fmaxf(acc[i].w + z.w + b.w, 0.f);
            own[i] = h;
          }
        }
        worker_sync();
        // ---- A(h1) -> left half: the tile is now cat1 ----
        if (arow < nr) {
          float4 acc[8];
          aggregate(WH, acc);
          float4* dstp = reinterpret_cast<float4*>(T + arow * TP + ac0);
#pragma unroll
          for (int i = 0; i < 8; ++i) dstp[i] = acc[i];
        }
        worker_sync();
        store_tile(a.cat1[st]);
        // ---- h2 = relu(cat1 @ w2cat^T + b2) -> right half ----
        convert_tile();
        wait_bar(smem_u32(&bar_acc_full[1]), ph_acc[1], dead, status, 26);
        ph_acc[1] ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
          uint32_t r[32];
          const int c0 = 32 * ehf;
          tmem_ld32((uint32_t)(W2H + c0), r);
          float* trow = T + (32 * eq + lane) * TP + WH + c0;
#pragma unroll
          for (int j = 0; j < 32; ++j) trow[j] = fmaxf(__uint_as_float(r[j]) + __ldg(a.b2 + c0 + j), 0.f);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        worker_sync();
        // ---- A(h2) -> left half: the tile is now cat2 ----
        if (arow < nr) {
          float4 acc[8];
          aggregate(WH, acc);
          float4* dstp = reinterpret_cast<float4*>(T + arow * TP + ac0);
#pragma unroll
          for (int i = 0; i < 8; ++i) dstp[i] = acc[i];
        }
        worker_sync();
        store_tile(a.cat2[st]);
        if (st == S - 1 && a.Cout != nullptr) {
          // C = sum_s csol[s] cat2_s: the last cat2 tile is still on chip, the earlier ones come back from L2
          constexpr int SLOTS = TM * (W2H / 4) / WORKERS / 2;
          for (int hf = 0; hf < 2; ++hf) {
            const int ibase = wt + hf * SLOTS * WORKERS;
            float4 acc[SLOTS];
            const float cl = a.csol[st];
#pragma unroll
            for (int u = 0; u < SLOTS; ++u) {
              const int idx = ibase + u * WORKERS, r = idx >> 5, c4 = idx & 31;
              const float4 v = *reinterpret_cast<const float4*>(T + r * TP + 4 * c4);
              acc[u] = make_float4(cl * v.x, cl * v.y, cl * v.z, cl * v.w);
            }
            for (int j = 0; j < st; ++j) {
              const float cf = a.csol[j];
              if (cf == 0.f) continue;
              const float* srcj = a.cat2[j];
              float4 v[SLOTS];
#pragma unroll
              for (int u = 0; u < SLOTS; ++u) {
                const int idx = ibase + u * WORKERS, r = idx >> 5, c4 = idx & 31;
                v[u] = (r < nr) ? *reinterpret_cast<const float4*>(srcj + (size_t)(r0 + r) * W2H + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
              }
#pragma unroll
              for (int u = 0; u < SLOTS; ++u) {
                acc[u].x = fmaf(cf, v[u].x, acc[u].x); acc[u].y = fmaf(cf, v[u].y, acc[u].y);
                acc[u].z = fmaf(cf, v[u].z, acc[u].z); acc[u].w = fmaf(cf, v[u].w, acc[u].w);
              }
            }
#pragma unroll
            for (int u = 0; u < SLOTS; ++u) {
              const int idx = ibase + u * WORKERS, r = idx >> 5, c4 = idx & 31;
              if (r < nr) *reinterpret_cast<float4*>(a.Cout + (size_t)(r0 + r) * W2H + 4 * c4) = acc[u];
            }
          }
        }
        worker_sync();      // the tile buffer is reused by the next stage; its cat2 rows are visible to this CTA
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS));
  }
}

// tiles[0] = number of tiles, tiles[1 + t] = first row of tile t, tiles[1 + n_tiles] = N.  Greedy packing of whole
// graphs (graph_ptr: node offsets, n_graphs + 1 entries) into tiles of at most TM rows; a graph larger than TM rows
// cannot be tiled: tiles[0] = -1.  One block: the offsets are staged in shared memory chunk by chunk (coalesced), one
// thread walks the chunk (the packing is inherently sequential, ~10 cycles per graph out of shared memory) and the
// block flushes the tile starts it produced.
constexpr int TB_CHUNK = 4096;
__global__ void __launch_bounds__(1024) k_tiles_build(const int64_t* __restrict__ graph_ptr, int64_t n_graphs,
                                                      int32_t* __restrict__ tiles) {
  __shared__ int32_t sp[TB_CHUNK + 1];
  __shared__ int32_t st[TB_CHUNK + 1];
  __shared__ int s_nt, s_new, s_start, s_bad;
  if (threadIdx.x == 0) { s_nt = 0; s_start = (int)graph_ptr[0]; s_bad = 0; tiles[1] = (int32_t)graph_ptr[0]; }
  __syncthreads();
  for (int64_t g0 = 0; g0 < n_graphs; g0 += TB_CHUNK) {
    const int cnt = (int)((n_graphs - g0 < TB_CHUNK) ? (n_graphs - g0) : TB_CHUNK);
    for (int i = threadIdx.x; i <= cnt; i += blockDim.x) sp[i] = (int32_t)graph_ptr[g0 + i];
    __syncthreads();
    if (threadIdx.x == 0) {
      int start = s_start, made = 0;
      for (int i = 0; i < cnt; ++i) {
        const int b = sp[i], e = sp[i + 1];
        if (e - b > TM) s_bad = 1;
        if (e - start > TM) { st[made++] = b; start = b; }   // graph i does not fit: a new tile starts at it
      }
      s_start = start;
      s_new = made;
    }
    __syncthreads();
    const int base = s_nt;
    for (int i = threadIdx.x; i < s_new; i += blockDim.x) tiles[2 + base + i] = st[i];
    __syncthreads();
    if (threadIdx.x == 0) s_nt = base + s_new;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int nt = s_nt + 1;
    tiles[1 + nt] = (int32_t)graph_ptr[n_graphs];
    tiles[0] = s_bad ? -1 : nt;
  }
}

}  // namespace chain

namespace tc { int* status_ptr(); }

bool chain_fwd_supported(const Sage3Ctx& c) { return c.H == chain::WH && c.use_tc && c.g_tiles != nullptr; }

int chain_fwd(Sage3Ctx& c, FoldWs& f, const Tableau& tb, float dt, float* Cout, cudaStream_t s) {
  int* status_dev = tc::status_ptr();
  if (!status_dev) { set_error("chain_fwd: status symbol unavailable"); return GNODE_ERR_CUDA; }
  chain::Args a{};
  a.z0 = f.z0;
  for (int st = 0; st < tb.S; ++st) {
    a.cat1[st] = f.cat1[st]; a.cat2[st] = f.cat2[st]; a.V[st] = f.V[st];
    double bsum = 0.0;
    for (int j = 0; j < st; ++j) { a.coef[st][j] = (float)tb.beta[st][j] * dt; bsum += tb.beta[st][j]; }
    a.c13_scale[st] = (float)bsum * dt;
    a.csol[st] = (float)tb.c_sol[st] * dt;
  }
  a.Cout = Cout;
  a.c13 = f.c13; a.b1 = c.b1; a.b2 = c.b2;
  a.img13 = f.sM13; a.img2 = c.s2;
  a.rowptr = c.g.rowptr; a.col = c.g.col;
  a.tiles = c.g_tiles;
  a.S = tb.S;
  a.status = status_dev;
  a.err = c.g_tile_err;
  GN_PROF(s, (double)c.N * tb.S * (2.0 * 128 * 128 + 2.0 * 128 * 64), 4.0 * (double)c.N * 128 * (1 + 3.0 * tb.S - 1),
          "chain_fwd S=%d", tb.S);
  static bool attr_set = false;
  if (!attr_set) {
    GN_CUDA(cudaFuncSetAttribute(chain::k_chain_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, chain::SMEM_BYTES));
    GN_CUDA(cudaFuncSetAttribute(chain::k_chain_fwd, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_set = true;
  }
  chain::k_chain_fwd<<<2 * kNumSMs, chain::THREADS, chain::SMEM_BYTES, s>>>(a);
  GN_LAUNCHED();
  return GNODE_OK;
}

}  // namespace gnode

using namespace gnode;

// tiles: device int32 [n_graphs + 2].  graph_ptr: device int64 [n_graphs + 1] node offsets of the graphs of the batch.
extern "C" int gnode_tiles_build(const int64_t* graph_ptr, int64_t n_graphs, int32_t* tiles, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_ARG(graph_ptr && tiles && n_graphs > 0, "gnode_tiles_build: bad argument");
  chain::k_tiles_build<<<1, 1024, 0, s>>>(graph_ptr, n_graphs, tiles);
  GN_LAUNCHED();
  return GNODE_OK;
}
