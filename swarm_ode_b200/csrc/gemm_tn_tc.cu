// tcgen05 engine for the weight-gradient contractions of the GNODE backward pass ("TN": the reduction runs
// over node rows), sm_100a only.  Stands for autograd's  dW = grad_out^T @ input  of every Linear inside
// SAGEConv (loss.backward(), scripts/train_gde.py:493).
//
//   C[p, q] += scale * sum_r A[r, p] * B[r, q]          A: [R, P], B: [R, Q], both dense row-major fp32
//
// One of the two operands is at most 128 wide (2H for the GNODE layers): it becomes the UMMA M operand, the
// other (up to 512 wide, D = 399 -> two N tiles of 208) the N operand, so ONE accumulator [128 x wn] in TMEM
// holds the whole product and every CTA streams its own contiguous range of node rows exactly once:
//   warp 0      producer: per stage (8 node rows = one tf32 K step) two 1-D bulk copies of the dense row
//               slabs (8 * w * 4 bytes each; 16-byte aligned for any w) into a deep raw ring.
//   warps 2-9   converters: transpose + 3xTF32 split.  Thread task = (column c, K chunk kc): four LDS of
//               raw[4kc..4kc+3][c] (conflict-free: consecutive lanes, consecutive columns), hi/lo split,
//               two 16-byte stores into the K-major no-swizzle operand planes (conflict-free).
//   warp 1      TMEM alloc + MMA issue: per stage, per N tile: lo*hi + hi*lo + hi*hi (kind::tf32, K = 8).
//   warps 2-5   (after their last stage) epilogue: TMEM -> per-CTA partial [wn][128] in global memory.
// A second tiny kernel reduces the per-CTA partials in fixed order (deterministic, no atomics).
#include "common.cuh"
#include "tc_common.cuh"

namespace gnode {
namespace tctn {
using namespace tc;

constexpr int BKR = 8;                   // node rows per K step (tf32 UMMA K = 8); a stage holds KS K steps
constexpr int MW = 128;                  // M operand width after zero padding
constexpr int LBO_M = MW * 16 + 16;      // bytes between consecutive 16-byte K chunks of an M-operand plane
constexpr int CONV_WARPS = 8;
constexpr int CONV_THREADS = CONV_WARPS * 32;
constexpr int THREADS = (2 + CONV_WARPS) * 32;   // 320
constexpr int MAX_RAW = 8, N_OP = 3;
constexpr int N_TASKS = 5;               // converter task budget: 2 * KS * (wm + wn) <= N_TASKS * CONV_THREADS
// The converter chain of a stage (wait raw -> 4 loads, split -> wait operand slot -> 2 stores -> proxy fence -> arrive) is a
// serial latency of ~1300 cycles per thread and stage however little work it holds, so the eight converter warps form
// TN_GROUPS independent groups that take alternate stages: two chains in flight per CTA (narrow operands 0.42 -> 0.37 ms
// over 1.56 M rows; four groups of two warps spill and lose, a fourth operand stage changes nothing: profiles/r2_ab_tn.txt).
#ifndef TN_GROUPS
#define TN_GROUPS 2
#endif
constexpr int GROUP_THREADS = CONV_THREADS / TN_GROUPS;
constexpr int G_TASKS = N_TASKS * TN_GROUPS;   // tasks per thread of a group

struct Args {
  const float* Am; int wm;     // M operand [rows, wm]
  const float* Bn; int wn;     // N operand [rows, wn]
  int bn, nt;                  // N tile width (multiple of 16, <= 256) and count; nt * bn >= wn
  int ks;                      // K steps (8 node rows each) per stage: 1 for wide operands, 2 / 4 for narrow ones
  int64_t n_kb;                // stages (8 * ks row blocks) in total
  int n_raw;                   // raw ring depth
  float* partials;             // [gridDim.x][wn][MW]
  float* colpart_m;            // optional [gridDim.x][TN_GROUPS][2 * ks][wm]: per-CTA, per-group, per-K-chunk column sums of the M operand
  float* colpart_n;            // optional [gridDim.x][TN_GROUPS][2 * ks][wn]
  int* status;
};

__global__ void __launch_bounds__(THREADS, 1) k_gemm_tn_tc(const Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_raw_full[MAX_RAW];
  __shared__ __align__(8) uint64_t bar_raw_empty[MAX_RAW];
  __shared__ __align__(8) uint64_t bar_op_full[N_OP];
  __shared__ __align__(8) uint64_t bar_op_empty[N_OP];
  __shared__ __align__(8) uint64_t bar_acc_full;
  __shared__ uint32_t tmem_holder;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int wm = a.wm, wn = a.wn, bn = a.bn, nt = a.nt, NR = a.n_raw, KS = a.ks;
  int* const status = a.status;
  const int wn_pad = nt * bn;
  const uint32_t lbo_n = (uint32_t)wn_pad * 16u + 16u;
  const uint32_t M_PLANE = 2u * (uint32_t)KS * LBO_M;
  const uint32_t n_plane = 2u * (uint32_t)KS * lbo_n;
  const uint32_t raw_m_bytes = 32u * (uint32_t)(KS * wm), raw_n_bytes = 32u * (uint32_t)(KS * wn);
  const uint32_t raw_stage = raw_m_bytes + raw_n_bytes;
  const uint32_t op_stage = 2u * M_PLANE + 2u * n_plane;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t op_off = (uint32_t)NR * raw_stage;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)wn_pad) tmem_cols <<= 1;
  // this CTA's contiguous range of 8-row blocks
  const int64_t kb0 = a.n_kb * blockIdx.x / gridDim.x, kb1 = a.n_kb * (blockIdx.x + 1) / gridDim.x;
  const int nkb = (int)(kb1 - kb0);

  if (tid == 0) {
    for (int s = 0; s < NR; ++s) { mbar_init(smem_u32(&bar_raw_full[s]), 1); mbar_init(smem_u32(&bar_raw_empty[s]), GROUP_THREADS); }
    for (int s = 0; s < N_OP; ++s) { mbar_init(smem_u32(&bar_op_full[s]), GROUP_THREADS); mbar_init(smem_u32(&bar_op_empty[s]), 1); }
    mbar_init(smem_u32(&bar_acc_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  // operand planes start as zeros: padding rows (m >= wm, n >= wn) are never written afterwards
  {
    uint4* z = reinterpret_cast<uint4*>(smem + op_off);
    const int n16 = (int)(N_OP * op_stage / 16);
    for (int i = tid; i < n16; i += THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_holder;

  if (warp == 0) {
    // =========================== producer ===========================
    if (lane == 0) {
      const int rows_m = BKR * KS * wm, rows_n = BKR * KS * wn;   // floats per stage
      const float* pm = a.Am + kb0 * (int64_t)rows_m;
      const float* pn = a.Bn + kb0 * (int64_t)rows_n;
      uint32_t s = 0, ph = 0;
      bool first_lap = true, ok = true;
      for (int i = 0; i < nkb && ok; ++i, pm += rows_m, pn += rows_n) {
        if (!first_lap) ok = mbar_wait(smem_u32(&bar_raw_empty[s]), ph ^ 1u, status, 11);
        const uint32_t dst = smem_base + s * raw_stage;
        const uint32_t bar = smem_u32(&bar_raw_full[s]);
        mbar_expect_tx(bar, raw_stage);
        bulk_load_1d(dst, pm, raw_m_bytes, bar);
        bulk_load_1d(dst + raw_m_bytes, pn, raw_n_bytes, bar);
        if (++s == (uint32_t)NR) { s = 0; ph ^= 1u; first_lap = false; }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = make_idesc(bn);
    const uint64_t desc_m = make_desc(0, LBO_M), desc_n = make_desc(0, lbo_n);
    uint32_t so = 0, po = 0;
    bool ok = true;
    for (int i = 0; i < nkb && ok; ++i) {
      ok = mbar_wait(smem_u32(&bar_op_full[so]), po, status, 12);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t m_hi = (smem_base + op_off + so * op_stage) >> 4, m_lo = m_hi + (M_PLANE >> 4);
        const uint32_t n_hi0 = m_hi + ((2u * M_PLANE) >> 4), n_lo0 = n_hi0 + (n_plane >> 4);
        for (int ks = 0; ks < KS; ++ks) {                          // K step ks uses chunks 2ks, 2ks + 1 of every plane
          const uint32_t am = (uint32_t)ks * ((2u * LBO_M) >> 4), an = (uint32_t)ks * ((2u * lbo_n) >> 4);
          const uint64_t dmh = desc_m | (uint64_t)(m_hi + am), dml = desc_m | (uint64_t)(m_lo + am);
          for (int j = 0; j < nt; ++j) {
            const uint32_t toff = (uint32_t)(j * bn) + an;         // 16 bytes per N row: bn rows = bn 16-byte units
            const uint64_t dnh = desc_n | (uint64_t)(n_hi0 + toff);
            const uint64_t dnl = desc_n | (uint64_t)(n_lo0 + toff);
            const uint32_t d = tmem_base + (uint32_t)(j * bn);
            umma_tf32(d, dml, dnh, idesc, (i > 0 || ks > 0) ? 1u : 0u);   // small terms first
            umma_tf32(d, dmh, dnl, idesc, 1u);
            umma_tf32(d, dmh, dnh, idesc, 1u);
          }
        }
        umma_commit(smem_u32(&bar_op_empty[so]));
        if (i == nkb - 1) umma_commit(smem_u32(&bar_acc_full));
      }
      __syncwarp();
      if (++so == (uint32_t)N_OP) { so = 0; po ^= 1u; }
    }
  } else {
    // =========================== converters ===========================
    const int t = (tid - 64) % GROUP_THREADS, grp = (tid - 64) / GROUP_THREADS;
    // Task tau = kc * (wm + wn) + column: K chunk kc (4 node rows) of one column of the M operand (column < wm) or of
    // the N operand.  src in floats from the raw stage base, dst in bytes from the operand-stage base (hi plane);
    // the lo plane sits lo_off bytes further.
    int src[G_TASKS], pitch[G_TASKS];
    uint32_t dst[G_TASKS], lo_off[G_TASKS];
    uint32_t valid = 0;
    const int wsum = wm + wn;
#pragma unroll
    for (int q = 0; q < G_TASKS; ++q) {
      const int task = t + q * GROUP_THREADS;
      src[q] = 0; pitch[q] = 0; dst[q] = 0; lo_off[q] = 0;
      if (task < 2 * KS * wsum) {
        const int kc = task / wsum, col = task - kc * wsum;
        if (col < wm) {
          src[q] = 4 * kc * wm + col; pitch[q] = wm;
          dst[q] = (uint32_t)kc * LBO_M + (uint32_t)col * 16u; lo_off[q] = M_PLANE;
        } else {
          const int n = col - wm;
          src[q] = BKR * KS * wm + 4 * kc * wn + n; pitch[q] = wn;
          dst[q] = 2u * M_PLANE + (uint32_t)kc * lbo_n + (uint32_t)n * 16u; lo_off[q] = n_plane;
        }
        valid |= 1u << q;
      }
    }
    float csum[G_TASKS];
#pragma unroll
    for (int q = 0; q < G_TASKS; ++q) csum[q] = 0.f;
    bool ok = true;
    for (int i = grp; i < nkb && ok; i += TN_GROUPS) {
      const uint32_t sr = (uint32_t)(i % NR), pr = (uint32_t)((i / NR) & 1);
      const uint32_t so = (uint32_t)(i % N_OP), po = (uint32_t)((i / N_OP) & 1);
      ok = mbar_wait(smem_u32(&bar_raw_full[sr]), pr, status, 13);
      const float* raw = reinterpret_cast<const float*>(smem + (size_t)sr * raw_stage);
      float v[G_TASKS][4];
#pragma unroll
      for (int q = 0; q < G_TASKS; ++q) {
        if (valid & (1u << q)) {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[q][e] = raw[src[q] + e * pitch[q]];
        }
      }
      if (i >= N_OP) ok = ok && mbar_wait(smem_u32(&bar_op_empty[so]), po ^ 1u, status, 14);
      uint8_t* op = smem + op_off + (size_t)so * op_stage;
#pragma unroll
      for (int q = 0; q < G_TASKS; ++q) {
        if (valid & (1u << q)) {
          uint4 h;
          float4 l;
          h.x = (__float_as_uint(v[q][0]) + 0x1000u) & 0xFFFFE000u; l.x = v[q][0] - __uint_as_float(h.x);
          h.y = (__float_as_uint(v[q][1]) + 0x1000u) & 0xFFFFE000u; l.y = v[q][1] - __uint_as_float(h.y);
          h.z = (__float_as_uint(v[q][2]) + 0x1000u) & 0xFFFFE000u; l.z = v[q][2] - __uint_as_float(h.z);
          h.w = (__float_as_uint(v[q][3]) + 0x1000u) & 0xFFFFE000u; l.w = v[q][3] - __uint_as_float(h.w);
          *reinterpret_cast<uint4*>(op + dst[q]) = h;
          *reinterpret_cast<float4*>(op + dst[q] + lo_off[q]) = l;
          csum[q] += (v[q][0] + v[q][1]) + (v[q][2] + v[q][3]);   // fused column sums (bias gradients)
        }
      }
      mbar_arrive(smem_u32(&bar_raw_empty[sr]));
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(smem_u32(&bar_op_full[so]));
    }
    // per-CTA column sums: task (kc, column) of group grp -> partial [cta][grp][kc][column] of its operand
#pragma unroll
    for (int q = 0; q < G_TASKS; ++q) {
      if (valid & (1u << q)) {
        const int task = t + q * GROUP_THREADS;
        const int kc = task / wsum, col = task - kc * wsum;
        const size_t slot = ((size_t)blockIdx.x * TN_GROUPS + grp) * 2 * KS + kc;
        if (col < wm) { if (a.colpart_m) a.colpart_m[slot * wm + col] = csum[q]; }
        else if (a.colpart_n) a.colpart_n[slot * wn + (col - wm)] = csum[q];
      }
    }
    // =========================== epilogue (warps 2..5: TMEM lane quadrants 2, 3, 0, 1) ===========================
    if (warp < 6 && nkb > 0 && ok) {
      const int qd = warp & 3;
      ok = mbar_wait(smem_u32(&bar_acc_full), 0, status, 15);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float* P = a.partials + (size_t)blockIdx.x * (size_t)wn * MW + 32 * qd + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * qd) << 16);
      for (int c0 = 0; c0 < wn && ok; c0 += 32) {
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
        for (int j = 0; j < 32; ++j)
          if (c0 + j < wn) P[(size_t)(c0 + j) * MW] = __uint_as_float(r[j]);   // lanes -> consecutive m: coalesced
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
  }
}

// C (+)= scale * sum_c partials[c][n][m]   with C[m, n] (n_major = 0) or C[n, m] (n_major = 1); m < wm
// followed (same launch) by the column sums:  out_m[c] += scale_m * sum over [S][2 ks] of colpart_m, same for n.
// A block of 256 threads owns 64 consecutive output elements; its four 64-thread groups each sum a contiguous quarter of
// the partials (coalesced 256-byte rows, four independent loads in flight), the four sums are combined through shared memory
// in a fixed order -> deterministic.
__global__ void __launch_bounds__(256) k_reduce_tn(const float* __restrict__ partials, int S, int wn, int wm, float* __restrict__ C,
                                                   int64_t ldc, int n_major, float scale, const float* __restrict__ colpart_m,
                                                   float* __restrict__ out_m, float scale_m, const float* __restrict__ colpart_n,
                                                   float* __restrict__ out_n, float scale_n, int ks) {
  __shared__ float part[4][64];
  const int e = threadIdx.x & 63, q = threadIdx.x >> 6;
  int64_t i = blockIdx.x * 64LL + e;
  const int64_t n_mat = (int64_t)wn * MW;
  const float* src = nullptr;     // partial r of this element at src[r * stride]
  int64_t stride = 0;
  int count = 0;
  float* dst = nullptr;
  float sc = 0.f;
  if (i < n_mat) {
    const int n = (int)(i / MW), m = (int)(i % MW);
    if (m < wm) {
      src = partials + i; stride = n_mat; count = S;
      dst = n_major ? C + (int64_t)n * ldc + m : C + (int64_t)m * ldc + n;
      sc = scale;
    }
  } else {
    i -= n_mat;
    if (i < wm) {
      if (out_m) { src = colpart_m + i; stride = wm; count = 2 * ks * S; dst = out_m + i; sc = scale_m; }
    } else if (i - wm < wn) {
      i -= wm;
      if (out_n) { src = colpart_n + i; stride = wn; count = 2 * ks * S; dst = out_n + i; sc = scale_n; }
    }
  }
  const int per = (count + 3) / 4;
  const int r0 = q * per, r1 = (r0 + per < count) ? r0 + per : count;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int r = r0;
  for (; r + 4 <= r1; r += 4) {
    s0 += __ldg(src + (size_t)r * stride);
    s1 += __ldg(src + (size_t)(r + 1) * stride);
    s2 += __ldg(src + (size_t)(r + 2) * stride);
    s3 += __ldg(src + (size_t)(r + 3) * stride);
  }
  for (; r < r1; ++r) s0 += __ldg(src + (size_t)r * stride);
  part[q][e] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (q == 0 && dst) *dst += sc * ((part[0][e] + part[1][e]) + (part[2][e] + part[3][e]));
}

struct Plan { bool ok; bool m_is_b; int wm, wn, bn, nt, ks; int64_t n_kb; int grid; };

Plan plan_for(const GemmTN& g) {
  Plan p{};
  p.ok = false;
  if (g.P < 1 || g.Q < 1 || g.Nrows < BKR) return p;
  if (g.lda != g.P || g.ldb != g.Q) return p;                      // dense row slabs only (1-D bulk copies)
  if ((reinterpret_cast<uintptr_t>(g.A) & 15) || (reinterpret_cast<uintptr_t>(g.B) & 15)) return p;
  if (g.Q <= MW && (g.P > MW || g.Q >= g.P)) { p.m_is_b = true; p.wm = g.Q; p.wn = g.P; }
  else if (g.P <= MW) { p.m_is_b = false; p.wm = g.P; p.wn = g.Q; }
  else return p;
  if (p.wn > 512) return p;
  const int npad = (p.wn + 15) & ~15;
  p.nt = (npad + 255) / 256;
  p.bn = ((npad + p.nt - 1) / p.nt + 15) & ~15;
  if (p.nt * p.bn > 512) return p;
  // rows per stage: as many K steps as the converter task budget allows (narrow operands -> fewer, fatter stages)
  p.ks = 1;
  while (p.ks < 4 && 2 * (2 * p.ks) * (p.wm + p.wn) <= N_TASKS * CONV_THREADS && g.Nrows >= (int64_t)BKR * 2 * p.ks * kNumSMs)
    p.ks *= 2;
  p.n_kb = g.Nrows / (BKR * p.ks);
  p.grid = (int)(p.n_kb < kNumSMs ? p.n_kb : kNumSMs);
  p.ok = true;
  return p;
}

}  // namespace tctn

namespace tc { int* status_ptr(); }

bool gemm_tn_tc_supported(const GemmTN& g) { return tctn::plan_for(g).ok; }

size_t gemm_tn_tc_workspace_floats(int P, int Q) {
  // worst case over both operand roles: kNumSMs partials of [wn][128], plus the column-sum partials [kNumSMs][TN_GROUPS][2 ks <= 8][P + Q]
  const int wn = (Q <= tctn::MW && (P > tctn::MW || Q >= P)) ? P : Q;
  return (size_t)kNumSMs * (size_t)wn * tctn::MW + (size_t)kNumSMs * 8 * TN_GROUPS * (size_t)(P + Q);
}

int gemm_tn_simt(const GemmTN& g, float* partials, cudaStream_t s);

int gemm_tn_tc(const GemmTN& g, float* partials, cudaStream_t s) {
  const tctn::Plan p = tctn::plan_for(g);
  if (!p.ok) { set_error("gemm_tn_tc: unsupported shape P=%d Q=%d rows=%lld", g.P, g.Q, (long long)g.Nrows); return GNODE_ERR_ARG; }
  int* status_dev = tc::status_ptr();
  if (!status_dev) { set_error("gemm_tn_tc: status symbol unavailable"); return GNODE_ERR_CUDA; }
  tctn::Args a;
  a.Am = p.m_is_b ? g.B : g.A; a.wm = p.wm;
  a.Bn = p.m_is_b ? g.A : g.B; a.wn = p.wn;
  a.bn = p.bn; a.nt = p.nt; a.ks = p.ks; a.n_kb = p.n_kb; a.partials = partials; a.status = status_dev;
  float* const out_m = p.m_is_b ? g.colsumB : g.colsumA;
  float* const out_n = p.m_is_b ? g.colsumA : g.colsumB;
  const float scale_m = p.m_is_b ? g.colsumB_scale : g.colsumA_scale;
  const float scale_n = p.m_is_b ? g.colsumA_scale : g.colsumB_scale;
  float* cp = partials + (size_t)p.grid * (size_t)p.wn * tctn::MW;
  a.colpart_m = out_m ? cp : nullptr;
  a.colpart_n = out_n ? cp + (size_t)p.grid * TN_GROUPS * 2 * p.ks * p.wm : nullptr;
  const size_t raw_stage = 32 * (size_t)p.ks * (size_t)(p.wm + p.wn);
  const size_t op_stage = 4 * (size_t)p.ks * ((size_t)tctn::LBO_M + ((size_t)p.nt * p.bn * 16 + 16));
  const size_t budget = 220 * 1024;
  int n_raw = (int)((budget - tctn::N_OP * op_stage) / raw_stage);
  if (n_raw > tctn::MAX_RAW) n_raw = tctn::MAX_RAW;
  if (n_raw < 2) { set_error("gemm_tn_tc: tile does not fit in shared memory"); return GNODE_ERR_ARG; }
  a.n_raw = n_raw;
  const size_t smem = n_raw * raw_stage + tctn::N_OP * op_stage;
  if (first_use_on_device(reinterpret_cast<const void*>(&tctn::k_gemm_tn_tc))) {
    GN_CUDA(cudaFuncSetAttribute(tctn::k_gemm_tn_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(225 * 1024)));
  }
  tctn::k_gemm_tn_tc<<<p.grid, tctn::THREADS, smem, s>>>(a);
  GN_LAUNCHED();
  const int64_t cnt = (int64_t)p.wn * tctn::MW + p.wm + p.wn;
  // C is [P, Q]: with the M operand = B (q = m) the partial index n is p -> rows of C are n
  tctn::k_reduce_tn<<<(unsigned)ceil_div64(cnt * 4, 256), 256, 0, s>>>(partials, p.grid, p.wn, p.wm, g.C, g.ldc,
                                                                    p.m_is_b ? 1 : 0, g.scale, a.colpart_m, out_m, scale_m,
                                                                    a.colpart_n, out_n, scale_n, p.ks * TN_GROUPS);
  GN_LAUNCHED();
  const int64_t done = p.n_kb * tctn::BKR * p.ks;
  if (done < g.Nrows) {   // trailing rows that do not fill a stage: FFMA kernel, accumulated on top
    GemmTN tail = g;
    tail.A = g.A + done * g.lda; tail.B = g.B + done * g.ldb; tail.Nrows = g.Nrows - done;
    GN_TRY(gemm_tn_simt(tail, partials, s));
  }
  return GNODE_OK;
}

}  // namespace gnode
