// Graph-resident backward chain of the folded RK stages (sm_100a, tcgen05): the transpose of chain_fwd.cu.
//
// With gz_s = dL/dZ_s and G3 = (dL/dy_1) @ w3cat (fold.cu), stage s of the backward pass of a step is 2H-wide only:
//
//   U_s    = dt * sum_{i>s} beta_is gz_i                      gcat2 = dt c_s G3 + U_s @ M13
//   g_v2   = (A^T(gcat2_l) + gcat2_r) * [h2_s > 0]            gcat1 = g_v2 @ w2cat
//   g_u1   = (A^T(gcat1_l) + gcat1_r) * [h1_s > 0]            gz_s  = [A^T(g_u1) | g_u1]
//
// A^T is the transpose of the mean aggregation: A^T(g)[j] = sum_{i : j -> i} g[i] / max(deg_in(i), 1).  No edge leaves
// its graph, so a tile of whole graphs (the same tiles as the forward chain) carries every row these sums touch.  One
// CTA owns a tile for ALL stages, last to first: gz_{s+1} is still on chip when U_s is formed, the other gz_i come back
// from L2.  HBM sees G3 and the ReLU sign bits (8 bytes per row and layer, kept by the forward chain) on the way in and gz_s, U_s, g_v2_s on the
// way out (operands of the weight-gradient contractions that follow: dW2cat += g_v2_s^T cat1_s, R += U_s^T cat2_s),
// plus GZ = sum_s gz_s after the last stage.  Structure, tile layout and the three-term tensor-core product are those
// of chain_fwd.cu (chain_common.cuh).
#include <cstdlib>

#include "chain_common.cuh"
#include "field.cuh"

namespace gnode {
namespace chain {

#ifdef CHAIN_TRACE
__device__ long long g_chain_trace_b[128];
#define CTB(i) do { if (blockIdx.x == 0 && wt == 0 && t == blockIdx.x + 2 * (int)gridDim.x) g_chain_trace_b[(i)] = clock64(); } while (0)
#else
#define CTB(i) do { } while (0)
#endif

struct BwdArgs {
  const float* G3;                      // [N, 2H]
  const uint32_t* mask[kMaxStages];     // [N, 4] sign bits of h1 (words 0, 1) and h2 (words 2, 3), written by chain_fwd
  const float* gz[kMaxStages];          // out [N, 2H] (written through gz_out, read back as a source)
  float* gz_out[kMaxStages];
  float* U[kMaxStages];                 // out [N, 2H] (stages with has_u)
  float* gv2[kMaxStages];               // out [N, H]
  float* GZ;                            // out [N, 2H]
  float cu[kMaxStages][kMaxStages];     // cu[s][i] = dt * beta[i][s], i > s
  float cs_dt[kMaxStages];              // dt * c_sol[s]
  int has_u[kMaxStages];
  const float *img13T, *img2T;          // chain-format images of M13^T [2H x 2H] and w2cat^T [2H x H]
  const int32_t *rowptr, *t_rowptr, *t_col;
  const int32_t* tiles;
  int S;
  int tile_rows;
  int* status;
  int* err;
};

// TR / NB as in chain_fwd.cu: rows kept per chunk, 128-row blocks per tile.
template <int TR, int NB>
__device__ __forceinline__ void chain_bwd_body(const BwdArgs& a, uint8_t* smem, uint16_t* s_deg, uint64_t* bar_b_full,
                                               uint64_t* bar_b_empty, uint64_t* bar_a_ready_p, uint64_t* bar_acc_full_p,
                                               uint32_t tmem_base, int* dead_flag_p) {
  constexpr int lbo_t = lbo_t_of(TR);
  constexpr uint32_t b_off = (uint32_t)t_bytes_of(TR);
  constexpr uint32_t n_slots = (uint32_t)ring_slots(TR);
  uint64_t& bar_a_ready = *bar_a_ready_p;
  uint64_t& bar_acc_full = *bar_acc_full_p;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int S = a.S;
  int* const status = a.status;
  volatile int* dead = dead_flag_p;
  uint8_t* const T = smem;
  const uint32_t smem_base = smem_u32(smem);
  const int n_tiles = a.tiles[0];
  auto rows_of = [&](int t) { return a.tiles[2 + t] - a.tiles[1 + t]; };
  auto blocks_of = [&](int t) { return (NB == 2 && rows_of(t) > TM) ? 2 : 1; };

  if (warp == 0) {
    // =========================== weight-image producer ===========================
    {
      uint32_t s = 0, ph = 0;
      bool first_lap = true;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int nblk = blocks_of(t);
        for (int st = S - 1; st >= 0; --st) {
          for (int g = (a.has_u[st] ? 0 : 1); g < 2; ++g) {       // g = 0: M13^T (K = 128), g = 1: w2cat^T (K = 64)
            const uint8_t* img = reinterpret_cast<const uint8_t*>(g == 0 ? a.img13T : a.img2T);
            const int nkb = g == 0 ? W2H / KB16 : WH / KB16;
            constexpr uint32_t bytes = stage_bytes(W2H);
            for (int kbb = 0; kbb < nkb * nblk; ++kbb) {
              const int kb = kbb % nkb;
              if (!first_lap) wait_bar(smem_u32(&bar_b_empty[s]), ph ^ 1u, dead, status, 31);
              if (elect_one()) {
                const uint32_t bar = smem_u32(&bar_b_full[s]);
                mbar_expect_tx(bar, bytes);
                bulk_load_1d(smem_base + b_off + s * B_STAGE, img + (size_t)kb * bytes, bytes, bar);
              }
              __syncwarp();
              if (++s == n_slots) { s = 0; ph ^= 1u; first_lap = false; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    uint32_t pa = 0, sb = 0, pb = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int nblk = blocks_of(t);
      for (int st = S - 1; st >= 0; --st) {
        for (int g = (a.has_u[st] ? 0 : 1); g < 2; ++g) {
         for (int b = 0; b < nblk; ++b) {
          const int nkb = g == 0 ? W2H / KB16 : WH / KB16;
          const uint32_t a_addr = smem_base + (g == 0 ? 0u : (uint32_t)(16 * lbo_t)) + (uint32_t)(b * TM * 16);   // g = 1: right half of the tile
          const int mm = (b == 1 && rows_of(t) - TM <= SHORT_BLOCK_ROWS) ? 64 : 128;
          wait_bar(smem_u32(&bar_a_ready), pa, dead, status, 33);
          pa ^= 1u;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          for (int kb = 0; kb < nkb; ++kb) {
            wait_bar(smem_u32(&bar_b_full[sb]), pb, dead, status, 32);
            if (elect_one()) {
              issue_kblock(tmem_base + ACC_COL, tmem_base + ALO_COL, a_addr, (uint32_t)lbo_t, smem_base + b_off + sb * B_STAGE, W2H, kb, kb == 0, mm);
              umma_commit(smem_u32(&bar_b_empty[sb]));
              if (kb == nkb - 1) umma_commit(smem_u32(&bar_acc_full));
            }
            __syncwarp();
            if (++sb == n_slots) { sb = 0; pb ^= 1u; }
          }
         }
        }
      }
    }
  } else {
    // =========================== workers ===========================
    const int wt = tid - 64;                       // 0..255
    const int cw = warp - 2;                       // 0..7
    uint32_t ph_acc = 0u;
    const int eq = warp & 3, ehf = cw >> 2;        // TMEM mapping: lane quadrant of this warp, column half
    const int elane = 32 * eq + lane;              // row of a 128-row block in the TMEM mapping
    const int alane = wt & (TM - 1), ach = (wt >> 7) * 8;  // aggregation mapping: two threads per row of a block, 8 chunks each
    auto Tp = [&](int chunk, int row) { return reinterpret_cast<float4*>(T + (size_t)chunk * lbo_t + row * 16); };
    constexpr int n_slots_t = TR * NCHUNK;
    auto arrive_a = [&]() {
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_a_ready));
    };

    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int r0 = a.tiles[1 + t];
      const int nr = a.tiles[2 + t] - r0;
      // out-neighbours of this thread's row (transposed CSR), tile-local, and 1 / in-degree of every row of the tile
      const int nblk = (NB == 2 && nr > TM) ? 2 : 1;
      int nb_b = 0, nb_e = 0;
      if (alane < nr) { nb_b = a.t_rowptr[r0 + alane]; nb_e = a.t_rowptr[r0 + alane + 1]; }
      int nbr[NBR_REG];
#pragma unroll
      for (int q = 0; q < NBR_REG; ++q) {
        int v = -1;
        if (nb_b + q < nb_e) {
          v = a.t_col[nb_b + q] - r0;
          if (v < 0 || v >= nr) { *a.err = 1; v = -1; }
        }
        nbr[q] = v;
      }
      // (kept as 16-bit in-degrees: the float table did not fit next to a second CTA with the two-slot ring of 140-row tiles)
      for (int rr = wt; rr < NB * TM; rr += WORKERS) {
        int d = 1;
        if (rr < nr) { d = a.rowptr[r0 + rr + 1] - a.rowptr[r0 + rr]; if (d > 65535) { *a.err = 1; d = 65535; } }
        s_deg[rr] = (uint16_t)(d > 1 ? d : 1);
      }
      auto inv_deg_of = [&](int row) { return __frcp_rn((float)s_deg[row]); };   // == 1.0f / (float)deg, correctly rounded
      // A^T over chunks [c0, c0 + 8) of row 128 b + alane (block 1 walks its slice of the transposed CSR in global memory)
      auto aggregate_t = [&](int b, int c0, float4 (&acc)[8]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        int pb_ = nb_b + NBR_REG, pe_ = nb_e;
        if (b == 0) {
#pragma unroll
          for (int q = 0; q < NBR_REG; ++q) {
            if (nbr[q] >= 0) {
              const float w = inv_deg_of(nbr[q]);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 v = *Tp(c0 + i, nbr[q]);
                acc[i].x = fmaf(w, v.x, acc[i].x); acc[i].y = fmaf(w, v.y, acc[i].y);
                acc[i].z = fmaf(w, v.z, acc[i].z); acc[i].w = fmaf(w, v.w, acc[i].w);
              }
            }
          }
        } else {
          pb_ = a.t_rowptr[r0 + TM + alane]; pe_ = a.t_rowptr[r0 + TM + alane + 1];
        }
        for (int p = pb_; p < pe_; ++p) {
          const int nb = a.t_col[p] - r0;
          if (nb < 0 || nb >= nr) { *a.err = 1; continue; }
          const float w = inv_deg_of(nb);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 v = *Tp(c0 + i, nb);
            acc[i].x = fmaf(w, v.x, acc[i].x); acc[i].y = fmaf(w, v.y, acc[i].y);
            acc[i].z = fmaf(w, v.z, acc[i].z); acc[i].w = fmaf(w, v.w, acc[i].w);
          }
        }
      };
      // right half <- (A^T(left half) + right half) * [h > 0]; which = 0: h1, 1: h2 (sign bits kept by the forward chain)
      auto relu_back = [&](const uint32_t* mask, int which) {
       for (int b = 0; b < nblk; ++b) {
        const int arow = b * TM + alane;
        if (arow < nr) {
          const uint32_t m = __ldg(mask + (size_t)(r0 + arow) * 4 + 2 * which + (ach >> 3));
          float4 acc[8];
          aggregate_t(b, ach, acc);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4* own = Tp(16 + ach + i, arow);
            const float4 g = *own;
            float4 o;
            o.x = (m >> (4 * i)) & 1u ? acc[i].x + g.x : 0.f; o.y = (m >> (4 * i + 1)) & 1u ? acc[i].y + g.y : 0.f;
            o.z = (m >> (4 * i + 2)) & 1u ? acc[i].z + g.z : 0.f; o.w = (m >> (4 * i + 3)) & 1u ? acc[i].w + g.w : 0.f;
            *own = o;
          }
        }
       }
      };
      // tile-linear combination (rows < nr; zeros elsewhere), optionally also written to `out`:
      //   mode 0:  T <- c_self * T + sum_{i in [i0, i1)} cu[st_][i] * gz_i        (U_st)
      //   mode 1:  T <- cs_dt[st_] * G3
      //   mode 2:  T <- c_self * T + sum_{i in [i0, i1)} gz_i                       (GZ)
      auto combine = [&](int mode, int st_, float c_self, int i0, int i1, float* out) {
        constexpr int SLOTS = TM * NCHUNK / WORKERS / 2;   // 8 per half block
        for (int hf = 0; hf < 2 * nblk; ++hf) {
          const int ibase = wt + hf * SLOTS * WORKERS;
          float4 acc[SLOTS];
#pragma unroll
          for (int u = 0; u < SLOTS; ++u) {
            const int idx = ibase + u * WORKERS, r = idx >> 5, c4 = idx & 31;
            acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nr && c_self != 0.f) { const float4 v = *Tp(c4, r); acc[u] = make_float4(c_self * v.x, c_self * v.y, c_self * v.z, c_self * v.w); }
          }
          for (int q = i0; q < i1; ++q) {
            const float cf = mode == 0 ? a.cu[st_][q] : (mode == 1 ? a.cs_dt[st_] : 1.f);
            if (cf == 0.f) continue;
            const float* sq = mode == 1 ? a.G3 : a.gz[q];
            float4 v[SLOTS];
#pragma unroll
            for (int u = 0; u < SLOTS; ++u) {
              const int idx = ibase + u * WORKERS, r = idx >> 5, c4 = idx & 31;
              v[u] = (r < nr) ? *reinterpret_cast<const float4*>(sq + (size_t)(r0 + r) * W2H + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < SLOTS; ++u) {
              acc[u].x = fmaf(cf, v[u].x, acc[u].x); acc[u].y = fmaf(cf, v[u].y, acc[u].y);
              acc[u].z = fmaf(cf, v[u].z, acc[u].z); acc[u].w = fmaf(cf, v[u].w, acc[u].w);
            }
          }
          if (out != nullptr) {
#pragma unroll
            for (int u = 0; u < SLOTS; ++u) {
              const int idx = ibase + u * WORKERS, r = idx >> 5, c4 = idx & 31;
              if (r < nr) __stcs(reinterpret_cast<float4*>(out + (size_t)(r0 + r) * W2H + 4 * c4), acc[u]);
            }
          }
#pragma unroll
          for (int u = 0; u < SLOTS; ++u) {
            const int idx = ibase + u * WORKERS, r = idx >> 5, c4 = idx & 31;
            if (idx < n_slots_t) *Tp(c4, r) = acc[u];
          }
        }
      };

      for (int st = S - 1; st >= 0; --st) {
        const float csd = a.cs_dt[st];
        CTB(16 * st + 0);
        if (a.has_u[st]) {
          // ---- U_st = sum_{i>st} cu * gz_i, in place (gz_{st+1} is still on chip); kept for R += U_st^T cat2_st ----
          combine(0, st, a.cu[st][st + 1], st + 2, S, a.U[st]);
          worker_sync_w();
          CTB(16 * st + 1);
          // ---- gcat2 = dt c_st G3 + U_st @ M13, block after block ----
          for (int b = 0; b < nblk; ++b) {
            residual_to_tmem(T, lbo_t, TR, tmem_base, eq, lane, 16 * ehf, 0, ALO_COL, b * TM);
            residual_to_tmem(T, lbo_t, TR, tmem_base, eq, lane, 16 * ehf + 8, 0, ALO_COL, b * TM);
            arrive_a();
            CTB(16 * st + 2);
            // G3 comes back in the tile-linear mapping (coalesced), first half requested before the accumulator is ready
            constexpr int ZS = 8;
            const int zbase = wt + b * 2 * ZS * WORKERS;
            float4 z[ZS];
#pragma unroll
            for (int u = 0; u < ZS; ++u) {
              const int idx = zbase + u * WORKERS, r = idx >> 5, c4 = idx & 31;
              z[u] = (r < nr && csd != 0.f) ? __ldg(reinterpret_cast<const float4*>(a.G3 + (size_t)(r0 + r) * W2H + 4 * c4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            wait_bar(smem_u32(&bar_acc_full), ph_acc, dead, status, 35);
            ph_acc ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            CTB(16 * st + 3);
            const int erow = b * TM + elane;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t r[32];
              tmem_ld32(tmem_base, eq, (uint32_t)(ACC_COL + 64 * ehf + 32 * h), r);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if (erow < TR)
                  *Tp(16 * ehf + 8 * h + i, erow) = make_float4(__uint_as_float(r[4 * i + 0]), __uint_as_float(r[4 * i + 1]),
                                                                __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
              }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            float4 z2[ZS];
#pragma unroll
            for (int u = 0; u < ZS; ++u) {
              const int idx = zbase + (ZS + u) * WORKERS, r = idx >> 5, c4 = idx & 31;
              z2[u] = (r < nr && csd != 0.f) ? __ldg(reinterpret_cast<const float4*>(a.G3 + (size_t)(r0 + r) * W2H + 4 * c4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            worker_sync();
#pragma unroll
            for (int u = 0; u < ZS; ++u) {
              const int idx = zbase + u * WORKERS, r = idx >> 5, c4 = idx & 31;
              if (r < nr) { float4* tp = Tp(c4, r); float4 v = *tp; v.x = fmaf(csd, z[u].x, v.x); v.y = fmaf(csd, z[u].y, v.y); v.z = fmaf(csd, z[u].z, v.z); v.w = fmaf(csd, z[u].w, v.w); *tp = v; }
            }
#pragma unroll
            for (int u = 0; u < ZS; ++u) {
              const int idx = zbase + (ZS + u) * WORKERS, r = idx >> 5, c4 = idx & 31;
              if (r < nr) { float4* tp = Tp(c4, r); float4 v = *tp; v.x = fmaf(csd, z2[u].x, v.x); v.y = fmaf(csd, z2[u].y, v.y); v.z = fmaf(csd, z2[u].z, v.z); v.w = fmaf(csd, z2[u].w, v.w); *tp = v; }
            }
            worker_sync_w();
          }
        } else {
          // ---- no later stage feeds on this one: gcat2 = dt c_st G3 ----
          combine(1, st, 0.f, 0, 1, nullptr);
          worker_sync_w();
        }
        CTB(16 * st + 4);
        // ---- g_v2 = (A^T(gcat2_l) + gcat2_r) * [h2 > 0] -> right half ----
        relu_back(a.mask[st], 1);
        worker_sync_w();
        CTB(16 * st + 5);
        // ---- gcat1 = g_v2 @ w2cat (K = 64: the right half is the operand); g_v2 goes out while the contraction runs ----
        for (int b = 0; b < nblk; ++b) {
          residual_to_tmem(T, lbo_t, TR, tmem_base, eq, lane, 16 + 8 * ehf, 16, ALO_COL, b * TM);
          arrive_a();
          CTB(16 * st + 6);
          if (b == 0) {
            for (int idx = wt; idx < nr * (NCHUNK / 2); idx += WORKERS) {
              const int r = idx >> 4, c4 = idx & 15;
              __stcs(reinterpret_cast<float4*>(a.gv2[st] + (size_t)(r0 + r) * WH + 4 * c4), *Tp(16 + c4, r));
            }
          }
          CTB(16 * st + 7);
          wait_bar(smem_u32(&bar_acc_full), ph_acc, dead, status, 36);
          ph_acc ^= 1u;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          // every thread has finished reading the tile for the g_v2 store; the epilogue of a block writes both halves of
          // its OWN rows only (its operand rows, whose contraction has completed), never the other block's operand rows
          worker_sync();
          CTB(16 * st + 8);
          const int erow = b * TM + elane;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t r[32];
            tmem_ld32(tmem_base, eq, (uint32_t)(ACC_COL + 64 * ehf + 32 * h), r);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (erow < TR)
                *Tp(16 * ehf + 8 * h + i, erow) = make_float4(__uint_as_float(r[4 * i + 0]), __uint_as_float(r[4 * i + 1]),
                                                              __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
            }
          }
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          worker_sync_w();
        }
        CTB(16 * st + 9);
        // ---- g_u1 = (A^T(gcat1_l) + gcat1_r) * [h1 > 0] -> right half ----
        relu_back(a.mask[st], 0);
        worker_sync_w();
        CTB(16 * st + 10);
        // ---- A^T(g_u1) -> left half: the tile is now gz_st ----
        for (int b = 0; b < nblk; ++b) {
          const int arow = b * TM + alane;
          if (arow < nr) {
            float4 acc[8];
            aggregate_t(b, 16 + ach, acc);
#pragma unroll
            for (int i = 0; i < 8; ++i) *Tp(ach + i, arow) = acc[i];
          }
        }
        worker_sync_w();
        CTB(16 * st + 11);
        for (int idx = wt; idx < nr * NCHUNK; idx += WORKERS) {
          const int r = idx >> 5, c4 = idx & 31;
          *reinterpret_cast<float4*>(a.gz_out[st] + (size_t)(r0 + r) * W2H + 4 * c4) = *Tp(c4, r);
        }
        if (st == 0) {
          // GZ = sum_s gz_s: gz_0 is on chip, the others come back from L2 (own stores: worker_sync makes them visible)
          worker_sync();
          combine(2, 0, 1.f, 1, S, a.GZ);
        }
        CTB(16 * st + 12);
        worker_sync();      // the tile buffer is modified by the next stage / tile
        CTB(16 * st + 13);
      }
    }
  }
}

__global__ void __launch_bounds__(THREADS, 2) k_chain_bwd(const BwdArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_b_full[MAX_SLOTS];
  __shared__ __align__(8) uint64_t bar_b_empty[MAX_SLOTS];
  __shared__ int s_tr;
  __shared__ __align__(8) uint64_t bar_a_ready;
  __shared__ __align__(8) uint64_t bar_acc_full;
  __shared__ uint32_t tmem_holder;
  __shared__ int dead_flag;
  __shared__ uint16_t s_deg[2 * TM];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int n_tiles = a.tiles[0];

  // tiles[0] < 0: a graph of the batch exceeds the tile capacity the caller announced (max_graph_nodes understated).
  // No tile is processed; flag it so that the deferred check raises instead of returning uninitialised results.
  if (n_tiles < 0 && blockIdx.x == 0 && tid == 0 && a.err != nullptr) *a.err = 2;

  if (tid == 0) {
    dead_flag = 0;
    int mr = 8;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) { const int nr = a.tiles[2 + t] - a.tiles[1 + t]; mr = nr > mr ? nr : mr; }
    s_tr = (mr + 7) & ~7;
    for (int s = 0; s < MAX_SLOTS; ++s) { mbar_init(smem_u32(&bar_b_full[s]), 1); mbar_init(smem_u32(&bar_b_empty[s]), 1); }
    mbar_init(smem_u32(&bar_a_ready), WORKERS / 32);
    mbar_init(smem_u32(&bar_acc_full), 1);
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
  if (a.tile_rows <= TM) {
    if (s_tr <= 96) chain_bwd_body<96, 1>(a, smem, s_deg, bar_b_full, bar_b_empty, &bar_a_ready, &bar_acc_full, tmem_base, &dead_flag);
    else chain_bwd_body<128, 1>(a, smem, s_deg, bar_b_full, bar_b_empty, &bar_a_ready, &bar_acc_full, tmem_base, &dead_flag);
  } else if (a.tile_rows <= TR_MID) {
    chain_bwd_body<TR_MID, 2>(a, smem, s_deg, bar_b_full, bar_b_empty, &bar_a_ready, &bar_acc_full, tmem_base, &dead_flag);
  } else {
    chain_bwd_body<TR_BIG, 2>(a, smem, s_deg, bar_b_full, bar_b_empty, &bar_a_ready, &bar_acc_full, tmem_base, &dead_flag);
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS));
  }
}

}  // namespace chain

namespace tc { int* status_ptr(); }
#ifdef CHAIN_TRACE
extern "C" int gnode_chain_trace_b(long long* out128) {
  return (int)cudaMemcpyFromSymbol(out128, chain::g_chain_trace_b, sizeof(long long) * 128);
}
#endif

bool chain_bwd_supported(const Sage3Ctx& c, const FoldWs& f) {
  static const bool off = [] { const char* e = std::getenv("GNODE_NO_CHAIN_BWD"); return e && e[0] == '1'; }();
  if (off) return false;
  return chain_fwd_supported(c) && c.ci2T != nullptr && f.ci13T != nullptr && f.Us[0] != nullptr && f.gv2s[0] != nullptr && f.mask[0] != nullptr &&
         c.g.t_rowptr != nullptr && c.g.t_col != nullptr;
}

// All stages of the backward pass of one step: reads f.G3 and the stage slots f.cat1 / f.cat2, fills f.gzs[s], f.Us[s]
// (stages with a later dependant), f.gv2s[s] and f.GZ.  has_u[s] reports which U_s were written.
int chain_bwd(Sage3Ctx& c, FoldWs& f, const Tableau& tb, float dt, bool* has_u, cudaStream_t s) {
  int* status_dev = tc::status_ptr();
  if (!status_dev) { set_error("chain_bwd: status symbol unavailable"); return GNODE_ERR_CUDA; }
  if (c.pend_ci2T) { GN_TRY(chain_pack_image(c.w2catT, 2 * c.H, c.H, c.H, c.ci2T, s)); c.pend_ci2T = 0; }   // see chain_fwd
  if (f.pend_ci13T) { GN_TRY(chain_pack_image(f.M13T, 2 * c.H, 2 * c.H, 2 * c.H, f.ci13T, s)); f.pend_ci13T = 0; }
  chain::BwdArgs a{};
  a.G3 = f.G3; a.GZ = f.GZ;
  int n_u = 0;
  for (int st = 0; st < tb.S; ++st) {
    a.mask[st] = f.mask[st];
    a.gz[st] = f.gzs[st]; a.gz_out[st] = f.gzs[st]; a.U[st] = f.Us[st]; a.gv2[st] = f.gv2s[st];
    a.cs_dt[st] = (float)tb.c_sol[st] * dt;
    a.has_u[st] = 0;
    for (int i = st + 1; i < tb.S; ++i) {
      a.cu[st][i] = (float)tb.beta[i][st] * dt;
      if (tb.beta[i][st] != 0.0) a.has_u[st] = 1;
    }
    has_u[st] = a.has_u[st] != 0;
    n_u += a.has_u[st];
  }
  a.img13T = f.ci13T; a.img2T = c.ci2T;
  a.rowptr = c.g.rowptr; a.t_rowptr = c.g.t_rowptr; a.t_col = c.g.t_col;
  a.tiles = c.g_tiles;
  a.S = tb.S;
  a.tile_rows = c.g.tile_rows > 0 ? c.g.tile_rows : chain::TM;
  a.status = status_dev;
  a.err = c.g_tile_err;
  GN_PROF(s, (double)c.N * (n_u * 2.0 * 128 * 128 + tb.S * 2.0 * 128 * 64),
          4.0 * (double)c.N * (128.0 * (tb.S + n_u + 1 + 1) + 64.0 * tb.S + 4.0 * tb.S), "chain_bwd S=%d", tb.S);
  if (first_use_on_device(reinterpret_cast<const void*>(&chain::k_chain_bwd))) {
    GN_CUDA(cudaFuncSetAttribute(chain::k_chain_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, chain::SMEM_BYTES_BIG));
    GN_CUDA(cudaFuncSetAttribute(chain::k_chain_bwd, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  }
  const bool big = a.tile_rows > chain::TR_MID;       // tiles of 145 .. 256 rows: one CTA per SM
  chain::k_chain_bwd<<<(big ? 1 : 2) * kNumSMs, chain::THREADS, chain::smem_bytes_of(a.tile_rows), s>>>(a);
  GN_LAUNCHED();
  return GNODE_OK;
}

}  // namespace gnode
