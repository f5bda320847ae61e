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
// aggregations A(.) need.  One CTA owns a tile for ALL stages of the step.  The [128 x 2H] fp32 tile lives in shared
// memory in the UMMA operand layout (chain_common.cuh): the two small contractions read it in place (tf32 leading
// term + weight-residual term) and take the tile's own residual as a bf16 operand from tensor memory, so a
// contraction costs ONE hand-off between the worker warps and the MMA warp.  The aggregations read shared memory,
// and HBM only sees Z_0 once on the way in and cat1_s / cat2_s / C (what the backward pass needs) on the way out;
// the cat2_j re-reads of later stages hit L2 (the same CTA wrote them microseconds earlier).
//
//   warp 0      weight-image producer: one 1-D bulk copy per K block of 16 (M13 / w2cat images, L2 resident)
//   warp 1      TMEM alloc + MMA issue
//   warps 2-9   workers: tile loads / combinations, residual operand, TMEM epilogues, aggregations, tile stores; they
//               advance in lockstep through named barrier 1.
#include <cstdlib>

#include "chain_common.cuh"
#include "field.cuh"

namespace gnode {
namespace chain {

#ifdef CHAIN_TRACE
__device__ long long g_chain_trace[128];
#define CT(i) do { if (blockIdx.x == 0 && wt == 0 && t == blockIdx.x + 2 * (int)gridDim.x) g_chain_trace[(i)] = clock64(); } while (0)
#else
#define CT(i) do { } while (0)
#endif

struct Args {
  const float* z0;
  float* cat1[kMaxStages];
  float* cat2[kMaxStages];
  float coef[kMaxStages][kMaxStages];   // dt * beta[s][j]
  float csol[kMaxStages];               // dt * c_sol[s]
  uint32_t* mask[kMaxStages];           // optional [N, 4]: sign bits of h1 (words 0, 1) and h2 (words 2, 3) for the backward chain
  float* Cout;                          // optional: C = sum_s csol[s] cat2_s, written after the last stage
  float* Cout2;                         // optional second combination with csol2 (dopri5: the error-estimate weights)
  float csol2[kMaxStages];
  float* Zlast;                         // optional [N, 2H]: Z of the LAST stage.  For an FSAL tableau (dopri5: the input of stage 6 is
                                        // the step's solution) this is Z_0 of the next step: y_1 @ w1cat^T without touching y_1
  float c13_scale[kMaxStages];          // dt * sum_j beta[s][j]
  const float *c13, *b1, *b2;
  const float *img13, *img2;            // chain-format weight images of M13 [2H x 2H] and w2cat [H x 2H]
  const int32_t *rowptr, *col;
  const int32_t* tiles;                 // [0] = number of tiles, [1 ..] = first row of every tile, then N
  int S;
  int tile_rows;                        // most rows a tile may hold (gnode_graph.tile_rows): selects the kernel variant
  int* status;                          // barrier-timeout word shared with the other tcgen05 kernels
  int* err;                             // set to 1 when a neighbour lies outside its tile
};

// TR = rows kept per chunk of the tile; NB = 128-row blocks per tile (1: tiles of <= 128 rows; 2: tiles of <= 256 rows, e.g.
// the 140-node graphs of 19 AGVs + 9 pickers).  With NB = 2 the two contractions of a stage run block after block through
// the same accumulator / residual columns of tensor memory (the weight images are streamed once per block), every other
// phase covers all rows of the tile.
template <int TR, int NB>
__device__ __forceinline__ void chain_fwd_body(const Args& a, uint8_t* smem, uint64_t* bar_b_full, uint64_t* bar_b_empty,
                                               uint64_t* bar_a_ready_p, uint64_t* bar_acc_full_p, uint32_t tmem_base,
                                               int* dead_flag_p) {
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
    // (the whole warp walks the schedule with warp-uniform state; one elected lane issues the copy: tc_common.cuh, elect_one)
    {
      uint32_t s = 0, ph = 0;
      bool first_lap = true;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int nblk = blocks_of(t);
        for (int st = 0; st < S; ++st) {
          for (int g = (st > 0 ? 0 : 1); g < 2; ++g) {           // g = 0: M13 (stage > 0 only), g = 1: w2cat
            const uint8_t* img = reinterpret_cast<const uint8_t*>(g == 0 ? a.img13 : a.img2);
            // a ring slot / bulk copy carries one K block of the N = 128 image or two of the N = 64 image: fewer, fatter
            // ring stages (every stage costs a copy completion, a barrier wait and a commit on top of its MMAs)
            const int kps = g == 0 ? 1 : 2;
            const uint32_t bytes = g == 0 ? stage_bytes(W2H) : 2 * stage_bytes(WH);
            for (int b = 0; b < nblk; ++b) {
              for (int kb = 0; kb < W2H / KB16 / kps; ++kb) {
                if (!first_lap) wait_bar(smem_u32(&bar_b_empty[s]), ph ^ 1u, dead, status, 21);
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
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    uint32_t pa = 0, sb = 0, pb = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int nblk = blocks_of(t);
      for (int st = 0; st < S; ++st) {
        for (int g = (st > 0 ? 0 : 1); g < 2; ++g) {
          const int n = g == 0 ? W2H : WH;
          for (int b = 0; b < nblk; ++b) {
            wait_bar(smem_u32(&bar_a_ready), pa, dead, status, 23);   // tile + residual operand of this block ready
            pa ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int kps = g == 0 ? 1 : 2;               // K blocks per ring slot
            const int mm = (b == 1 && rows_of(t) - TM <= SHORT_BLOCK_ROWS) ? 64 : 128;
            for (int kb = 0; kb < W2H / KB16; kb += kps) {
              wait_bar(smem_u32(&bar_b_full[sb]), pb, dead, status, 22);
              if (elect_one()) {
                for (int j = 0; j < kps; ++j)
                  issue_kblock(tmem_base + ACC_COL, tmem_base + ALO_COL, smem_base + (uint32_t)(b * TM * 16), (uint32_t)lbo_t,
                               smem_base + b_off + sb * B_STAGE + (uint32_t)(j * stage_bytes(WH)), n, kb + j, kb + j == 0, mm);
                umma_commit(smem_u32(&bar_b_empty[sb]));
                if (kb + kps == W2H / KB16) umma_commit(smem_u32(&bar_acc_full));
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
    // TMEM mapping: lane quadrant of this warp, column half
    const int eq = warp & 3, ehf = cw >> 2;
    const int elane = 32 * eq + lane;              // row of a 128-row block this thread owns in the TMEM mapping
    // aggregation mapping: two threads per row of a block, 32 of the 64 channels (8 chunks) each
    const int alane = wt & (TM - 1), ach = (wt >> 7) * 8;
    auto Tp = [&](int chunk, int row) { return reinterpret_cast<float4*>(T + (size_t)chunk * lbo_t + row * 16); };
    constexpr int n_slots_t = TR * NCHUNK;         // float4 slots of the tile (tile-linear mapping: idx -> row idx >> 5, chunk idx & 31)
    constexpr int SLOTS = TM * NCHUNK / WORKERS / 2;   // 8 slots per thread and half block
    // residual operand of one block (K = 128) -> tensor memory, then hand the contraction to the MMA warp
    auto hand_off = [&](int b) {
      residual_to_tmem(T, lbo_t, TR, tmem_base, eq, lane, 16 * ehf, 0, ALO_COL, b * TM);
      residual_to_tmem(T, lbo_t, TR, tmem_base, eq, lane, 16 * ehf + 8, 0, ALO_COL, b * TM);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_a_ready));
    };

    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int r0 = a.tiles[1 + t];
      const int nr = a.tiles[2 + t] - r0;                 // rows of the tile: whole graphs
      const int nblk = (NB == 2 && nr > TM) ? 2 : 1;
      // CSR slice of this thread's row of block 0 (aggregation phases); its first NBR_REG neighbour ids (tile-local) stay
      // in registers for the three aggregations of every stage.
      int nb_b = 0, nb_e = 0;
      if (alane < nr) { nb_b = a.rowptr[r0 + alane]; nb_e = a.rowptr[r0 + alane + 1]; }
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
      // the same for this thread's row of block 1 (two-block tiles only)
      int nb1_b = 0, nb1_e = 0;
      int nbr1[NBR_REG];
#pragma unroll
      for (int q = 0; q < NBR_REG; ++q) nbr1[q] = -1;
      if (NB == 2 && TM + alane < nr) {
        nb1_b = a.rowptr[r0 + TM + alane]; nb1_e = a.rowptr[r0 + TM + alane + 1];
#pragma unroll
        for (int q = 0; q < NBR_REG; ++q) {
          if (nb1_b + q < nb1_e) {
            int v = a.col[nb1_b + q] - r0;
            if (v < 0 || v >= nr) { *a.err = 1; v = -1; }
            nbr1[q] = v;
          }
        }
      }
      const float inv_deg1 = 1.0f / (float)((nb1_e - nb1_b) > 1 ? (nb1_e - nb1_b) : 1);

      // mean over the in-neighbours of chunks [c0, c0 + 8) of row 128 b + alane
      auto aggregate = [&](int b, int c0, float4 (&acc)[8]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int pb_ = (b == 0 ? nb_b : nb1_b) + NBR_REG, pe_ = b == 0 ? nb_e : nb1_e;
        const float w = b == 0 ? inv_deg : inv_deg1;
#pragma unroll
        for (int q = 0; q < NBR_REG; ++q) {
          const int nq = b == 0 ? nbr[q] : nbr1[q];
          if (nq >= 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 v = *Tp(c0 + i, nq);
              acc[i].x += v.x; acc[i].y += v.y; acc[i].z += v.z; acc[i].w += v.w;
            }
          }
        }
        // neighbours beyond the NBR_REG kept in registers: ids fetched eight at a time (one dependent global / L2 load per
        // neighbour made rows of dense graphs a ~250-cycle latency chain per neighbour)
        // Tiles of one block (NB = 1, the medium warehouse) keep the scalar loop: the eight-wide id batch costs the forward
        // kernel 4 % there through register pressure (A/B in profiles/r2_ab_idbatch.txt); two-block tiles (graphs of 129 .. 256
        // nodes, where dense graphs live) fetch ids eight at a time: forward chain of the 256-agent complete graphs 4.68 -> 2.27
        // ms.  (The backward kernel keeps its scalar loop: every variant of the batch, and splitting it into two kernels as
        // here, cost its one-block path 2.5 - 3.6 %.)
        if (NB == 1) {
        for (int p = pb_; p < pe_; ++p) {
          const int nb = a.col[p] - r0;
          if (nb < 0 || nb >= nr) { *a.err = 1; continue; }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 v = *Tp(c0 + i, nb);
            acc[i].x += v.x; acc[i].y += v.y; acc[i].z += v.z; acc[i].w += v.w;
          }
        }
        } else {
        for (int p = pb_; p < pe_; p += 8) {
          int ids[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) ids[q] = (p + q < pe_) ? __ldg(a.col + p + q) - r0 : -1;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int nb = ids[q];
            if (p + q >= pe_) continue;
            if (nb < 0 || nb >= nr) { *a.err = 1; continue; }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 v = *Tp(c0 + i, nb);
              acc[i].x += v.x; acc[i].y += v.y; acc[i].z += v.z; acc[i].w += v.w;
            }
          }
        }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { acc[i].x *= w; acc[i].y *= w; acc[i].z *= w; acc[i].w *= w; }
      };
      // coalesced tile store: rows < nr of T -> dst[(r0 + r) * 2H + c]
      // (streaming = evict-first in L2: for outputs that this kernel does not read back)
      auto store_tile = [&](float* dst, bool streaming) {
#pragma unroll 4
        for (int idx = wt; idx < nr * NCHUNK; idx += WORKERS) {
          const int r = idx >> 5, c4 = idx & 31;
          float4* gp = reinterpret_cast<float4*>(dst + (size_t)(r0 + r) * W2H + 4 * c4);
          if (streaming) __stcs(gp, *Tp(c4, r)); else *gp = *Tp(c4, r);
        }
      };

      for (int st = 0; st < S; ++st) {
        CT(16 * st + 0);
        if (st + 1 < S) {   // Z_0 comes back in the epilogue of the next stage: keep its lines in L2
          for (int b = 0; b < nblk; ++b) {
            if (b * TM + elane < nr) {
              const float* zl = a.z0 + (size_t)(r0 + b * TM + elane) * W2H + 64 * ehf;
              prefetch_l2(zl); prefetch_l2(zl + 32);
            }
          }
        }
        // ---- tile input: Z_0 (stage 0) or V_st = sum_j coef * cat2_j, in place (cat2_{st-1} is still on chip) ----
        for (int hf = 0; hf < 2 * nblk; ++hf) {
          const int ibase = wt + hf * SLOTS * WORKERS;
          float4 acc[SLOTS];
          if (st == 0) {
#pragma unroll
            for (int u = 0; u < SLOTS; ++u) {
              const int idx = ibase + u * WORKERS, r = idx >> 5, c4 = idx & 31;
              acc[u] = (r < nr) ? __ldg(reinterpret_cast<const float4*>(a.z0 + (size_t)(r0 + r) * W2H + 4 * c4))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          } else {
            const float cl = a.coef[st][st - 1];
#pragma unroll
            for (int u = 0; u < SLOTS; ++u) {
              const int idx = ibase + u * WORKERS, r = idx >> 5, c4 = idx & 31;
              acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (r < nr) { const float4 v = *Tp(c4, r); acc[u] = make_float4(cl * v.x, cl * v.y, cl * v.z, cl * v.w); }
            }
            for (int j = 0; j < st - 1; ++j) {
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
          }
#pragma unroll
          for (int u = 0; u < SLOTS; ++u) {
            const int idx = ibase + u * WORKERS, r = idx >> 5, c4 = idx & 31;
            if (idx < n_slots_t) *Tp(c4, r) = acc[u];
          }
        }
        worker_sync_w();
        CT(16 * st + 1);
        if (st > 0) {
          // ---- Z_st = Z_0 + V_st @ M13^T + scale * c13, block after block ----
          const float cs = a.c13_scale[st];
          for (int b = 0; b < nblk; ++b) {
            hand_off(b);
            CT(16 * st + 2);
            // Z_0 comes back in the tile-linear mapping (coalesced: a warp reads one 512-byte row; a row-per-lane load
            // would touch 32 lines per instruction).  Its first half is requested before the accumulator is ready.
            const int zbase = wt + b * 2 * SLOTS * WORKERS;
            float4 z[SLOTS];
#pragma unroll
            for (int u = 0; u < SLOTS; ++u) {
              const int idx = zbase + u * WORKERS, r = idx >> 5, c4 = idx & 31;
              z[u] = (r < nr) ? __ldg(reinterpret_cast<const float4*>(a.z0 + (size_t)(r0 + r) * W2H + 4 * c4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            wait_bar(smem_u32(&bar_acc_full), ph_acc, dead, status, 25);
            ph_acc ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            CT(16 * st + 3);
            // accumulator + scale * c13 -> tile, lane = row
            // (every warp runs this, also one whose lane quadrant holds no tile rows: skipping it measured 20 % slower)
            const int erow = b * TM + elane;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int c0 = 64 * ehf + 32 * h;
              uint32_t r[32];
              tmem_ld32(tmem_base, eq, (uint32_t)(ACC_COL + c0), r);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 c = __ldg(reinterpret_cast<const float4*>(a.c13 + c0) + i);
                float4 o;
                o.x = fmaf(cs, c.x, __uint_as_float(r[4 * i + 0])); o.y = fmaf(cs, c.y, __uint_as_float(r[4 * i + 1]));
                o.z = fmaf(cs, c.z, __uint_as_float(r[4 * i + 2])); o.w = fmaf(cs, c.w, __uint_as_float(r[4 * i + 3]));
                if (erow < TR) *Tp(c0 / 4 + i, erow) = o;
              }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            // second half of Z_0 on its way while the workers meet
            float4 z2[SLOTS];
#pragma unroll
            for (int u = 0; u < SLOTS; ++u) {
              const int idx = zbase + (SLOTS + u) * WORKERS, r = idx >> 5, c4 = idx & 31;
              z2[u] = (r < nr) ? __ldg(reinterpret_cast<const float4*>(a.z0 + (size_t)(r0 + r) * W2H + 4 * c4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            worker_sync();
#pragma unroll
            for (int u = 0; u < SLOTS; ++u) {
              const int idx = zbase + u * WORKERS, r = idx >> 5, c4 = idx & 31;
              if (r < nr) { float4* tp = Tp(c4, r); float4 v = *tp; v.x += z[u].x; v.y += z[u].y; v.z += z[u].z; v.w += z[u].w; *tp = v; }
            }
#pragma unroll
            for (int u = 0; u < SLOTS; ++u) {
              const int idx = zbase + (SLOTS + u) * WORKERS, r = idx >> 5, c4 = idx & 31;
              if (r < nr) { float4* tp = Tp(c4, r); float4 v = *tp; v.x += z2[u].x; v.y += z2[u].y; v.z += z2[u].z; v.w += z2[u].w; *tp = v; }
            }
            worker_sync_w();
          }
          if (st == S - 1 && a.Zlast != nullptr) {
            // Z of the last stage goes out as it stands in the tile (a pass of its own, outside the loops above: the
            // fixed-grid solvers never take it and must not pay registers for it); the next phase rewrites the right
            // half in place, hence the barrier
            for (int idx = wt; idx < nr * 32; idx += WORKERS) {
              const int r = idx >> 5, c4 = idx & 31;
              __stcs(reinterpret_cast<float4*>(a.Zlast + (size_t)(r0 + r) * W2H + 4 * c4), *Tp(c4, r));
            }
            worker_sync_w();
          }
        }
        CT(16 * st + 4);
        // ---- h1 = relu(A(Z_l) + Z_r + b1) -> right half (in place) ----
        for (int b = 0; b < nblk; ++b) {
          const int arow = b * TM + alane;
          if (arow < nr) {
            float4 acc[8];
            aggregate(b, ach, acc);
            const float4* bb = reinterpret_cast<const float4*>(a.b1 + 4 * ach);
            uint32_t mbits = 0u;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4* own = Tp(16 + ach + i, arow);
              const float4 z = *own, bv = __ldg(bb + i);
              float4 h;
              h.x = fmaxf(acc[i].x + z.x + bv.x, 0.f); h.y = fmaxf(acc[i].y + z.y + bv.y, 0.f);
              h.z = fmaxf(acc[i].z + z.z + bv.z, 0.f); h.w = fmaxf(acc[i].w + z.w + bv.w, 0.f);
              *own = h;
              mbits |= (h.x > 0.f ? 1u : 0u) << (4 * i) | (h.y > 0.f ? 2u : 0u) << (4 * i) | (h.z > 0.f ? 4u : 0u) << (4 * i) |
                       (h.w > 0.f ? 8u : 0u) << (4 * i);
            }
            if (a.mask[st] != nullptr) a.mask[st][(size_t)(r0 + arow) * 4 + (ach >> 3)] = mbits;
          }
        }
        worker_sync_w();
        CT(16 * st + 5);
        // ---- A(h1) -> left half: the tile is now cat1 ----
        for (int b = 0; b < nblk; ++b) {
          const int arow = b * TM + alane;
          if (arow < nr) {
            float4 acc[8];
            aggregate(b, 16 + ach, acc);
#pragma unroll
            for (int i = 0; i < 8; ++i) *Tp(ach + i, arow) = acc[i];
          }
        }
        worker_sync_w();
        CT(16 * st + 6);
        // ---- h2 = relu(cat1 @ w2cat^T + b2) -> right half, block after block; cat1 goes out while the first contraction
        // runs ----
        for (int b = 0; b < nblk; ++b) {
          hand_off(b);
          CT(16 * st + 7);
          if (b == 0 && a.cat1[st] != nullptr) store_tile(a.cat1[st], true);   // (null: forward-only solve, nobody reads cat1)
          CT(16 * st + 8);
          wait_bar(smem_u32(&bar_acc_full), ph_acc, dead, status, 26);
          ph_acc ^= 1u;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          uint32_t r[32];
          tmem_ld32(tmem_base, eq, (uint32_t)(ACC_COL + 32 * ehf), r);
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          if (b == 0) worker_sync();           // every thread has finished reading the tile for the cat1 store
          CT(16 * st + 9);
          const int erow = b * TM + elane;
          uint32_t mbits = 0u;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(a.b2 + 32 * ehf) + i);
            float4 h;
            h.x = fmaxf(__uint_as_float(r[4 * i + 0]) + bv.x, 0.f); h.y = fmaxf(__uint_as_float(r[4 * i + 1]) + bv.y, 0.f);
            h.z = fmaxf(__uint_as_float(r[4 * i + 2]) + bv.z, 0.f); h.w = fmaxf(__uint_as_float(r[4 * i + 3]) + bv.w, 0.f);
            if (erow < TR) *Tp(16 + 8 * ehf + i, erow) = h;
            mbits |= (h.x > 0.f ? 1u : 0u) << (4 * i) | (h.y > 0.f ? 2u : 0u) << (4 * i) | (h.z > 0.f ? 4u : 0u) << (4 * i) |
                     (h.w > 0.f ? 8u : 0u) << (4 * i);
          }
          if (a.mask[st] != nullptr && erow < nr) a.mask[st][(size_t)(r0 + erow) * 4 + 2 + ehf] = mbits;
          if (NB == 2 && b + 1 < nblk) worker_sync_w();   // block 1's operand rows must not change under its residual pass
        }
        worker_sync_w();
        CT(16 * st + 10);
        // ---- A(h2) -> left half: the tile is now cat2 ----
        for (int b = 0; b < nblk; ++b) {
          const int arow = b * TM + alane;
          if (arow < nr) {
            float4 acc[8];
            aggregate(b, 16 + ach, acc);
#pragma unroll
            for (int i = 0; i < 8; ++i) *Tp(ach + i, arow) = acc[i];
          }
        }
        worker_sync_w();
        CT(16 * st + 11);
        store_tile(a.cat2[st], st == S - 1 && a.Cout == nullptr);
        CT(16 * st + 12);
        if (st == S - 1 && a.Cout != nullptr) {
          // C = sum_s csol[s] cat2_s (and, for dopri5, the same sum with the error-estimate weights): the last cat2 tile is
          // still on chip, the earlier ones come back from L2 once for both combinations
          constexpr int CS = SLOTS / 2;
          const bool two = a.Cout2 != nullptr;
          for (int q = 0; q < 4 * nblk; ++q) {
            const int ibase = wt + q * CS * WORKERS;
            float4 acc[CS], acc2[CS];
            const float cl = a.csol[st], cl2 = two ? a.csol2[st] : 0.f;
#pragma unroll
            for (int u = 0; u < CS; ++u) {
              const int idx = ibase + u * WORKERS, r = idx >> 5, c4 = idx & 31;
              acc[u] = make_float4(0.f, 0.f, 0.f, 0.f); acc2[u] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (r < nr) {
                const float4 v = *Tp(c4, r);
                acc[u] = make_float4(cl * v.x, cl * v.y, cl * v.z, cl * v.w);
                acc2[u] = make_float4(cl2 * v.x, cl2 * v.y, cl2 * v.z, cl2 * v.w);
              }
            }
            for (int j = 0; j < st; ++j) {
              const float cf = a.csol[j], cf2 = two ? a.csol2[j] : 0.f;
              if (cf == 0.f && cf2 == 0.f) continue;
              const float* srcj = a.cat2[j];
              float4 v[CS];
#pragma unroll
              for (int u = 0; u < CS; ++u) {
                const int idx = ibase + u * WORKERS, r = idx >> 5, c4 = idx & 31;
                v[u] = (r < nr) ? *reinterpret_cast<const float4*>(srcj + (size_t)(r0 + r) * W2H + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
              }
#pragma unroll
              for (int u = 0; u < CS; ++u) {
                acc[u].x = fmaf(cf, v[u].x, acc[u].x); acc[u].y = fmaf(cf, v[u].y, acc[u].y);
                acc[u].z = fmaf(cf, v[u].z, acc[u].z); acc[u].w = fmaf(cf, v[u].w, acc[u].w);
                acc2[u].x = fmaf(cf2, v[u].x, acc2[u].x); acc2[u].y = fmaf(cf2, v[u].y, acc2[u].y);
                acc2[u].z = fmaf(cf2, v[u].z, acc2[u].z); acc2[u].w = fmaf(cf2, v[u].w, acc2[u].w);
              }
            }
#pragma unroll
            for (int u = 0; u < CS; ++u) {
              const int idx = ibase + u * WORKERS, r = idx >> 5, c4 = idx & 31;
              if (r < nr) {
                __stcs(reinterpret_cast<float4*>(a.Cout + (size_t)(r0 + r) * W2H + 4 * c4), acc[u]);
                if (two) __stcs(reinterpret_cast<float4*>(a.Cout2 + (size_t)(r0 + r) * W2H + 4 * c4), acc2[u]);
              }
            }
          }
        }
        worker_sync();      // the tile buffer is modified by the next stage; its cat2 rows are visible to this CTA
        CT(16 * st + 13);
      }
    }
  }

}

// TWO_BLOCK selects the family of tile variants compiled into the kernel: tiles of <= 128 rows, or tiles of two 128-row
// blocks (graphs of 129 .. 256 nodes).  Two kernels instead of one: register allocation of the hot one-block variants is
// then independent of the two-block code (a shared kernel cost the one-block path 4 % when the two-block path changed).
template <bool TWO_BLOCK>
__global__ void __launch_bounds__(THREADS, 2) k_chain_fwd(const Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_b_full[MAX_SLOTS];
  __shared__ __align__(8) uint64_t bar_b_empty[MAX_SLOTS];
  __shared__ int s_tr;
  __shared__ __align__(8) uint64_t bar_a_ready;
  __shared__ __align__(8) uint64_t bar_acc_full;
  __shared__ uint32_t tmem_holder;
  __shared__ int dead_flag;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int n_tiles = a.tiles[0];

  // tiles[0] < 0: a graph of the batch exceeds the tile capacity the caller announced (max_graph_nodes understated).
  // No tile is processed; flag it so that the deferred check raises instead of returning uninitialised results.
  if (n_tiles < 0 && blockIdx.x == 0 && tid == 0 && a.err != nullptr) *a.err = 2;

  if (tid == 0) {
    dead_flag = 0;
    int mr = 8;                                                  // most rows of any tile of this CTA
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
  // rows kept per chunk: compile-time variants so that the tile addressing folds into immediates
#ifndef CHAIN_TR_SMALL
#define CHAIN_TR_SMALL 96
#endif
  // (the host launches with the shared memory of the variant the batch's largest graph needs: a.tile_rows)
  if (!TWO_BLOCK) {
    if (s_tr <= CHAIN_TR_SMALL) chain_fwd_body<CHAIN_TR_SMALL, 1>(a, smem, bar_b_full, bar_b_empty, &bar_a_ready, &bar_acc_full, tmem_base, &dead_flag);
    else chain_fwd_body<128, 1>(a, smem, bar_b_full, bar_b_empty, &bar_a_ready, &bar_acc_full, tmem_base, &dead_flag);
  } else if (a.tile_rows <= TR_MID) {
    chain_fwd_body<TR_MID, 2>(a, smem, bar_b_full, bar_b_empty, &bar_a_ready, &bar_acc_full, tmem_base, &dead_flag);
  } else {
    chain_fwd_body<TR_BIG, 2>(a, smem, bar_b_full, bar_b_empty, &bar_a_ready, &bar_acc_full, tmem_base, &dead_flag);
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS));
  }
}

// chain-format weight image (chain_common.cuh): one 16-byte unit per thread
__global__ void k_pack_image(const float* __restrict__ W, int n, int k, int64_t ld, uint4* __restrict__ img) {
  const int units_per_chunk = n + 1, units_per_stage = 10 * units_per_chunk;
  const int64_t total = (int64_t)(k / KB16) * units_per_stage;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int kb = (int)(i / units_per_stage), u = (int)(i % units_per_stage);
    const int chunk = u / units_per_chunk, row = u % units_per_chunk;
    uint4 out = make_uint4(0u, 0u, 0u, 0u);
    if (row < n) {
      const float* src = W + (size_t)row * ld + kb * KB16;
      if (chunk < 8) {
        const int c = chunk & 3;
        uint32_t v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float w = src[4 * c + e];
          const uint32_t hb = (__float_as_uint(w) + 0x1000u) & 0xFFFFE000u;
          v[e] = chunk < 4 ? hb : __float_as_uint(w - __uint_as_float(hb));
        }
        out = make_uint4(v[0], v[1], v[2], v[3]);
      } else {
        const int c = chunk - 8;
        uint32_t v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(v[e]) : "f"(src[8 * c + 2 * e + 1]), "f"(src[8 * c + 2 * e]));
        }
        out = make_uint4(v[0], v[1], v[2], v[3]);
      }
    }
    img[i] = out;
  }
}

int pack_image(const float* W, int n, int k, int64_t ld, float* img, cudaStream_t s) {
  const int64_t total = (int64_t)(k / KB16) * 10 * (n + 1);
  k_pack_image<<<(unsigned)ceil_div64(total, 256), 256, 0, s>>>(W, n, k, ld, reinterpret_cast<uint4*>(img));
  GN_LAUNCHED();
  return GNODE_OK;
}

// tiles[0] = number of tiles, tiles[1 + t] = first row of tile t, tiles[1 + n_tiles] = N.  Greedy packing of whole
// graphs (graph_ptr: node offsets, n_graphs + 1 entries) into tiles of at most TM rows; a graph larger than TM rows
// cannot be tiled: tiles[0] = -1.  One block: the offsets are staged in shared memory chunk by chunk (coalesced), one
// thread walks the chunk (the packing is inherently sequential, ~10 cycles per graph out of shared memory) and the
// block flushes the tile starts it produced.
constexpr int TB_CHUNK = 4096;
__global__ void __launch_bounds__(1024) k_tiles_build(const int64_t* __restrict__ graph_ptr, int64_t n_graphs,
                                                      int32_t* __restrict__ tiles, int max_rows) {
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
        if (e - b > max_rows) s_bad = 1;
        if (e - start > max_rows) { st[made++] = b; start = b; }   // graph i does not fit: a new tile starts at it
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

// Parallel form of the same greedy packing for batches whose offsets fit in shared memory.  The packing is a chain
// 0 -> nxt(0) -> nxt(nxt(0)) ... with nxt(g) = first graph that no longer fits into a tile starting at g (binary
// search over the offsets, one per graph).  The orbit of 0 is marked by pointer doubling (log2 rounds: graphs already
// marked mark their 2^k-th successor), a block-wide prefix count turns marks into tile indices.
constexpr int TBP_MAX = 16384;           // graphs: 13 bytes of shared memory each
__global__ void __launch_bounds__(1024) k_tiles_build_par(const int64_t* __restrict__ graph_ptr, int n, int32_t* __restrict__ tiles,
                                                          int max_rows) {
  extern __shared__ __align__(16) uint8_t tb_smem[];
  int32_t* sp = reinterpret_cast<int32_t*>(tb_smem);          // [n + 1] offsets
  int32_t* ja = sp + (n + 1);                                 // [n] jump (double buffered)
  int32_t* jb = ja + n;
  uint8_t* mk = reinterpret_cast<uint8_t*>(jb + n);           // [n] marks
  __shared__ int s_bad, s_warp[32];
  const int tid = threadIdx.x;
  if (tid == 0) s_bad = 0;
  for (int i = tid; i <= n; i += 1024) sp[i] = (int32_t)graph_ptr[i];
  __syncthreads();
  for (int g = tid; g < n; g += 1024) {
    if (sp[g + 1] - sp[g] > max_rows) s_bad = 1;
    // largest h in (g, n] with sp[h] - sp[g] <= max_rows: graphs g .. h-1 share the tile
    int lo = g + 1, hi = n;
    const int lim = sp[g] + max_rows;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (sp[mid] <= lim) lo = mid; else hi = mid - 1; }
    ja[g] = lo;                                               // == n: the tile runs to the end of the batch
    mk[g] = g == 0;
  }
  __syncthreads();
  int32_t *jc = ja, *jn = jb;
  for (int span = 1; span < n; span <<= 1) {
    for (int g = tid; g < n; g += 1024) {
      const int j = jc[g];
      if (j < n && mk[g]) mk[j] = 1;
      jn[g] = j < n ? jc[j] : n;
    }
    __syncthreads();
    int32_t* t = jc; jc = jn; jn = t;
  }
  // tile index of a marked graph = number of marked graphs before it
  const int per = (n + 1023) / 1024, g0 = tid * per, g1 = min(n, g0 + per);
  int cnt = 0;
  for (int g = g0; g < g1; ++g) cnt += mk[g];
  int incl = cnt;
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += v; }
  if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    int w = s_warp[tid];
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, w, o); if (tid >= o) w += v; }
    s_warp[tid] = w;
  }
  __syncthreads();
  int idx = incl - cnt + ((tid >> 5) > 0 ? s_warp[(tid >> 5) - 1] : 0);
  for (int g = g0; g < g1; ++g) if (mk[g]) tiles[1 + idx++] = sp[g];
  if (tid == 0) {
    const int nt = s_warp[31];
    tiles[1 + nt] = sp[n];
    tiles[0] = s_bad ? -1 : nt;
  }
}

}  // namespace chain

namespace tc { int* status_ptr(); }

size_t chain_image_floats(int n, int k) { return chain::image_floats(n, k); }
int chain_pack_image(const float* W, int n, int k, int64_t ld, float* img, cudaStream_t s) { return chain::pack_image(W, n, k, ld, img, s); }
bool chain_shape_ok(int H) { return H == chain::WH; }
bool chain_fwd_supported(const Sage3Ctx& c) { return c.H == chain::WH && c.use_tc && c.g_tiles != nullptr && c.ci2 != nullptr; }

int chain_fwd(Sage3Ctx& c, FoldWs& f, const Tableau& tb, float dt, float* Cout, cudaStream_t s, float* Cout2, const double* coef2) {
  int* status_dev = tc::status_ptr();
  if (!status_dev) { set_error("chain_fwd: status symbol unavailable"); return GNODE_ERR_CUDA; }
  // weight images of this kernel: packed by the first launch that reads them (Sage3Ctx::pack / FoldWs::prepare only mark them)
  if (c.pend_ci2) { GN_TRY(chain_pack_image(c.w2cat, c.H, 2 * c.H, 2 * c.H, c.ci2, s)); c.pend_ci2 = 0; }
  if (f.pend_ci13) { GN_TRY(chain_pack_image(f.M13, 2 * c.H, 2 * c.H, 2 * c.H, f.ci13, s)); f.pend_ci13 = 0; }
  chain::Args a{};
  a.z0 = f.z0;
  for (int st = 0; st < tb.S; ++st) {
    // cat1 and the ReLU sign bits exist for the backward pass only
    a.cat1[st] = f.forward_only ? nullptr : f.cat1[st]; a.cat2[st] = f.cat2[st]; a.mask[st] = f.forward_only ? nullptr : f.mask[st];
    double bsum = 0.0;
    for (int j = 0; j < st; ++j) { a.coef[st][j] = (float)tb.beta[st][j] * dt; bsum += tb.beta[st][j]; }
    a.c13_scale[st] = (float)bsum * dt;
    a.csol[st] = (float)tb.c_sol[st] * dt;
  }
  a.Cout = Cout;
  a.Zlast = (f.z_next && tb.S > 1) ? f.z_next : nullptr;
  a.Cout2 = (Cout && Cout2 && coef2) ? Cout2 : nullptr;
  if (a.Cout2) for (int st = 0; st < tb.S; ++st) a.csol2[st] = (float)coef2[st] * dt;
  a.c13 = f.c13; a.b1 = c.b1; a.b2 = c.b2;
  a.img13 = f.ci13; a.img2 = c.ci2;
  a.rowptr = c.g.rowptr; a.col = c.g.col;
  a.tiles = c.g_tiles;
  a.S = tb.S;
  a.tile_rows = c.g.tile_rows > 0 ? c.g.tile_rows : chain::TM;
  a.status = status_dev;
  a.err = c.g_tile_err;
  GN_PROF(s, (double)c.N * tb.S * (2.0 * 128 * 128 + 2.0 * 128 * 64), 4.0 * (double)c.N * 128 * (2 + 2.0 * tb.S),
          "chain_fwd S=%d", tb.S);
  if (first_use_on_device(reinterpret_cast<const void*>(&chain::k_chain_fwd<false>))) {
    GN_CUDA(cudaFuncSetAttribute(chain::k_chain_fwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, chain::SMEM_BYTES));
    GN_CUDA(cudaFuncSetAttribute(chain::k_chain_fwd<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    GN_CUDA(cudaFuncSetAttribute(chain::k_chain_fwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, chain::SMEM_BYTES_BIG));
    GN_CUDA(cudaFuncSetAttribute(chain::k_chain_fwd<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  }
  const bool big = a.tile_rows > chain::TR_MID;       // tiles of 145 .. 256 rows: one CTA per SM
  unsigned grid = (big ? 1 : 2) * kNumSMs;
#ifdef CHAIN_TRACE
  { const char* e = std::getenv("CHAIN_GRID"); if (e && atoi(e) > 0) grid = (unsigned)atoi(e); }
#endif
  if (a.tile_rows <= chain::TM) chain::k_chain_fwd<false><<<grid, chain::THREADS, chain::smem_bytes_of(a.tile_rows), s>>>(a);
  else chain::k_chain_fwd<true><<<grid, chain::THREADS, chain::smem_bytes_of(a.tile_rows), s>>>(a);
  GN_LAUNCHED();
  return GNODE_OK;
}

}  // namespace gnode

using namespace gnode;

#ifdef CHAIN_TRACE
extern "C" int gnode_chain_trace(long long* out128) {
  return (int)cudaMemcpyFromSymbol(out128, chain::g_chain_trace, sizeof(long long) * 128);
}
#endif

// tiles: device int32 [n_graphs + 2].  graph_ptr: device int64 [n_graphs + 1] node offsets of the graphs of the batch.
// max_rows: most rows a tile may hold (128, or up to 256 for batches with graphs of 129 .. 256 nodes).
extern "C" int gnode_tiles_build_rows(const int64_t* graph_ptr, int64_t n_graphs, int32_t max_rows, int32_t* tiles,
                                      gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_ARG(graph_ptr && tiles && n_graphs > 0, "gnode_tiles_build: bad argument");
  GN_ARG(max_rows >= 1 && max_rows <= chain::TR_BIG, "gnode_tiles_build: max_rows must be in [1, %d]", chain::TR_BIG);
  if (n_graphs <= chain::TBP_MAX) {
    const size_t smem = (size_t)(n_graphs + 1) * 4 + (size_t)n_graphs * 9 + 16;
    if (first_use_on_device(reinterpret_cast<const void*>(&chain::k_tiles_build_par))) {
      GN_CUDA(cudaFuncSetAttribute(chain::k_tiles_build_par, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)(chain::TBP_MAX + 1) * 4 + (size_t)chain::TBP_MAX * 9 + 16)));
    }
    chain::k_tiles_build_par<<<1, 1024, smem, s>>>(graph_ptr, (int)n_graphs, tiles, max_rows);
  } else {
    chain::k_tiles_build<<<1, 1024, 0, s>>>(graph_ptr, n_graphs, tiles, max_rows);
  }
  GN_LAUNCHED();
  return GNODE_OK;
}

extern "C" int gnode_tiles_build(const int64_t* graph_ptr, int64_t n_graphs, int32_t* tiles, gnode_stream_t stream) {
  return gnode_tiles_build_rows(graph_ptr, n_graphs, chain::TM, tiles, stream);
}
