// Shared pieces of the graph-resident chain kernels (chain_fwd.cu, chain_bwd.cu; sm_100a, tcgen05).
//
// Tile layout.  A tile of <= 128 consecutive rows (whole graphs) x 2H = 128 fp32 columns lives in shared memory in the
// UMMA K-major no-swizzle layout: chunk c (columns 4c .. 4c+3) holds the 16-byte row pieces of all 128 rows back to
// back (8-row core matrices of 128 bytes), chunks LBO_T bytes apart.  The tile therefore IS the A operand of
// tcgen05.mma kind::tf32: the tensor core reads the fp32 containers and ignores the 13 low mantissa bits (truncation;
// probed on B200 with scripts/dev/probe_umma.cu), so no conversion pass and no operand copy is needed for the leading
// term.  fp32-grade accuracy comes from two correction terms:
//
//     A.B  ~=  trunc(A).B_hi + trunc(A).B_lo + bf16(A - trunc(A)).bf16(B)
//
// B_hi = rna_tf32(B), B_lo = B - B_hi (tf32 planes of the weights), and the residual of A (< 2^-10 |A|) as a bf16 A
// operand held in TENSOR MEMORY (TS form of kind::f16: lane = row, 32-bit column c = k 2c | 2c+1), multiplied with a
// bf16 copy of B; the terms dropped are O(2^-19) relative.  Weight images are pre-packed per K block of 16 as
// [hi: 4 chunks | lo: 4 chunks | bf16: 2 chunks], each chunk = N rows x 16 bytes + 16 bytes of padding, so that one
// 1-D bulk copy fetches a whole pipeline stage.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace gnode {
namespace chain {
using namespace tc;

constexpr int TM = 128;                       // rows per tile = TMEM lanes
constexpr int W2H = 128, WH = 64;             // 2H, H (the chain kernels are specialised for hidden_dim = 64)
constexpr int NCHUNK = W2H / 4;               // 32
// The tile keeps only TR = round_up(max rows of the CTA's tiles, 8) rows per chunk: chunk pitch lbo_t = 16 TR + 16.
// The MMA still reads 128 rows per chunk (rows >= TR alias the following chunks: garbage in accumulator lanes that
// nobody reads).  95-node graphs (12 AGVs + 7 pickers x 5 snapshots) need 49.7 KB instead of 66 KB, which buys a third
// slot of the weight ring.
__host__ __device__ constexpr int lbo_t_of(int tr) { return tr * 16 + 16; }
__host__ __device__ constexpr int t_bytes_of(int tr) { return NCHUNK * lbo_t_of(tr); }
constexpr int KB16 = 16;                      // K per pipeline stage of the weight stream
constexpr int WORKERS = 256, THREADS = 64 + WORKERS;
constexpr int NBR_REG = 4;                    // neighbour ids per row kept in registers
constexpr int ACC_COL = 0, ALO_COL = 128;     // TMEM columns: accumulator (<= 128), bf16 residual operand (<= 64)
constexpr int TMEM_COLS = 256;

__host__ __device__ constexpr int lbo_b(int n) { return n * 16 + 16; }
__host__ __device__ constexpr int stage_bytes(int n) { return 10 * lbo_b(n); }     // hi 4 + lo 4 + bf16 2 chunks
constexpr int B_STAGE = 2 * stage_bytes(WH);  // 20800: ring slot = one K block of an N = 128 image (20640) or TWO of an N = 64 image
constexpr int MAX_SLOTS = 4;
constexpr int SMEM_BYTES = t_bytes_of(96) + 3 * B_STAGE;   // 112064: two CTAs per SM; 2 slots for 128-row tiles, 3 for <= 96 rows
// Tiles of more than 128 rows (graphs of 129 .. 256 nodes) are processed as two 128-row blocks.  Up to TR_MID = 140 rows
// (28 agents x 5 snapshots: the 19 AGV + 9 picker warehouse) the tile keeps exactly TR_MID rows per chunk (chunk pitch
// 141 x 16 bytes: still an odd number of 16-byte units, so the 128-bit accesses stay conflict free), which leaves room for
// a TWO-slot weight ring next to a second CTA on the SM (a one-slot ring serialises every bulk copy with the MMAs that
// consume it: 7 us per block and contraction instead of 3.7).  Beyond that one CTA per SM with the full tile.
constexpr int TR_MID = 140, TR_BIG = 256;
constexpr int SMEM_BYTES_MID = t_bytes_of(TR_MID) + 2 * B_STAGE;   // 113792 (+ ~1.2 KB static: two CTAs need <= 115712 each)
constexpr int SMEM_BYTES_BIG = t_bytes_of(TR_BIG) + 2 * B_STAGE;   // 173184
__host__ __device__ constexpr int smem_bytes_of(int tr) { return tr <= TM ? SMEM_BYTES : (tr <= TR_MID ? SMEM_BYTES_MID : SMEM_BYTES_BIG); }
__host__ __device__ constexpr int ring_slots(int tr) {
  return (smem_bytes_of(tr) - t_bytes_of(tr)) / B_STAGE > MAX_SLOTS ? MAX_SLOTS : (smem_bytes_of(tr) - t_bytes_of(tr)) / B_STAGE;
}
static_assert(B_STAGE >= stage_bytes(W2H), "slot too small");
static_assert(ring_slots(128) >= 2 && ring_slots(TR_MID) >= 2 && ring_slots(TR_BIG) >= 2, "ring too small");
static_assert((lbo_t_of(TR_MID) / 16) % 2 == 1 && (lbo_t_of(96) / 16) % 2 == 1, "chunk pitch must be an odd number of 16-byte units");
// A second block of at most 16 rows runs as M = 64 MMAs: rows 0 .. 15 of an M = 64 accumulator sit in lanes 0 .. 15 exactly
// as for M = 128 (scripts/dev/probe_umma_m64.cu), and the TS form reads the residual row of lane L from lane L
// (probe_umma_m64_ts.cu), so only the instruction descriptor changes; the tensor core reads half the A rows.
constexpr int SHORT_BLOCK_ROWS = 16;

inline size_t image_floats(int n, int k) { return (size_t)(k / KB16) * stage_bytes(n) / 4; }
// img <- chain-format image of row-major W [n x k] (row stride ld); n % 8 == 0, k % 16 == 0
int pack_image(const float* W, int n, int k, int64_t ld, float* img, cudaStream_t s);

__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(WORKERS) : "memory"); }
// after a phase that wrote the tile: make the writes visible to the async proxy (tcgen05.mma reads the tile)
__device__ __forceinline__ void worker_sync_w() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  worker_sync();
}

// bounded wait that gives up immediately once any wait of this CTA has timed out
__device__ __forceinline__ void wait_bar(uint32_t addr, uint32_t parity, volatile int* dead, int* status, int code) {
  for (uint32_t i = 0; i < FAST_POLLS; ++i)            // fast path: no shared-memory flag read, no clock read
    if (mbar_try_wait(addr, parity)) return;
  const uint64_t t0 = globaltimer_ns();
  do {
    for (uint32_t i = 0; i < 256; ++i)
      if (mbar_try_wait(addr, parity)) return;
    if (*dead) return;                                 // another wait of this CTA already timed out
  } while (globaltimer_ns() - t0 < WAIT_LIMIT_NS);
  *dead = 1;
  if (status) atomicExch(status, code);
}

// instruction descriptor of kind::f16 with bf16 operands, fp32 accumulation, both K-major, M = 128
__device__ __forceinline__ uint32_t make_idesc_bf16(int bn, int bm = 128) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(bm >> 4) << 24);
}
// D[tmem] (+)= A[tmem] . B[smem]^T, A = bf16 in tensor memory (TS form)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}

// One K = 16 block of the three-term product: tile chunks [4 kb, 4 kb + 4) x weight stage at `bstage`.
__device__ __forceinline__ void issue_kblock(uint32_t tmem_acc, uint32_t tmem_alo, uint32_t tile_addr, uint32_t lbo_t,
                                             uint32_t bstage, int n, int kb, bool first, int m = 128) {
  const uint32_t lb = (uint32_t)lbo_b(n);
  const uint64_t dT = make_desc(0, lbo_t), dB = make_desc(0, lb);
  const uint32_t idf = make_idesc(n, m), idb = make_idesc_bf16(n, m);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const uint64_t da = dT | (uint64_t)(((tile_addr + (uint32_t)(4 * kb + 2 * h) * lbo_t) >> 4) & 0x3FFF);
    const uint64_t dlo = dB | (uint64_t)(((bstage + (uint32_t)(4 + 2 * h) * lb) >> 4) & 0x3FFF);
    const uint64_t dhi = dB | (uint64_t)(((bstage + (uint32_t)(2 * h) * lb) >> 4) & 0x3FFF);
    umma_tf32(tmem_acc, da, dlo, idf, (first && h == 0) ? 0u : 1u);    // small term first
    umma_tf32(tmem_acc, da, dhi, idf, 1u);
  }
  const uint64_t dbf = dB | (uint64_t)(((bstage + 8u * lb) >> 4) & 0x3FFF);
  umma_bf16_ts(tmem_acc, tmem_alo + (uint32_t)(8 * kb), dbf, idb, 1u);
}

// L2 prefetch of the 128-byte line at p (no register cost: hides the HBM latency of a load issued a stage later)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// 32 TMEM columns starting at `col` of lane quadrant `eq` -> r[]
__device__ __forceinline__ void tmem_ld32(uint32_t tmem_base, int eq, uint32_t col, uint32_t (&r)[32]) {
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
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&p)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(p[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]), "r"(p[4]), "r"(p[5]), "r"(p[6]), "r"(p[7]), "r"(p[8]), "r"(p[9]),
      "r"(p[10]), "r"(p[11]), "r"(p[12]), "r"(p[13]), "r"(p[14]), "r"(p[15])
      : "memory");
}

// bf16x2 of the parts of (x, y) that the tf32 truncation drops: low half = x, high half = y
__device__ __forceinline__ uint32_t residual_pair(float x, float y) {
  const float rx = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  const float ry = y - __uint_as_float(__float_as_uint(y) & 0xFFFFE000u);
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(ry), "f"(rx));
  return d;
}

// Residual operand of the tile -> tensor memory.  Thread (row = 32 eq + lane, part): chunks [c_first, c_first + 8) of
// its row = 32 k values = 16 TMEM columns starting at ALO_COL + 2 (c_first - c_base).
// Rows >= tr do not exist in the tile: zeros (warp-uniform skip when the whole quadrant is out of range).
// row_off: first row of the 128-row block whose residual is taken (tiles of two blocks).
__device__ __forceinline__ void residual_to_tmem(const uint8_t* tile, int lbo_t, int tr, uint32_t tmem_base, int eq, int lane,
                                                 int c_first, int c_base, int alo_col = ALO_COL, int row_off = 0) {
  if (row_off + 32 * eq >= tr) return;
  const int row = row_off + 32 * eq + lane;
  uint32_t p[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 x = row < tr ? *reinterpret_cast<const float4*>(tile + (size_t)(c_first + i) * lbo_t + row * 16)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
    p[2 * i] = residual_pair(x.x, x.y);
    p[2 * i + 1] = residual_pair(x.z, x.w);
  }
  tmem_st16(tmem_base + ((uint32_t)(32 * eq) << 16) + (uint32_t)(alo_col + 2 * (c_first - c_base)), p);
}

}  // namespace chain
}  // namespace gnode
