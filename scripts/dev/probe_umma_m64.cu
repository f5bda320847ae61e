// Hardware probe (sm_100a) for round 2 of the row-major K = 128 engine: where does a cta_group::1 tcgen05.mma with M = 64
// put its 64 accumulator rows in tensor memory, and does the lane field of the D address move them (lane offsets 16 and 64)?
// If two M = 64 accumulators can live in disjoint lane sets, a [128-lane x 400-column] TMEM region can be double-buffered
// by rows.   build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I swarm_ode_b200/csrc scripts/dev/probe_umma_m64.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "tc_common.cuh"
using namespace gnode::tc;

constexpr int M = 64, N = 16;
constexpr int LBO_A = M * 16, LBO_B = N * 16;

// out[3][128][16]: columns 0..15 after MMA #1 (D lane field 0), 16..31 after MMA #2 (lane field 16), 32..47 after MMA #3 (lane field 64)
__global__ void k_probe(float* out) {
  __shared__ __align__(128) uint8_t sA[2 * LBO_A];
  __shared__ __align__(128) uint8_t sB[2 * LBO_B];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t holder;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&holder)), "r"(64u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid < M)   // A[r][k] = 100 r + k + 1 (tf32-exact), K = 8 -> two 16-byte chunks per row
    for (int k = 0; k < 8; ++k)
      *reinterpret_cast<float*>(sA + (k >> 2) * LBO_A + (tid >> 3) * 128 + (tid & 7) * 16 + (k & 3) * 4) = (float)(100 * tid + k + 1);
  if (tid < N)   // B = identity on the first 8 columns
    for (int k = 0; k < 8; ++k)
      *reinterpret_cast<float*>(sB + (k >> 2) * LBO_B + (tid >> 3) * 128 + (tid & 7) * 16 + (k & 3) * 4) = (tid == k) ? 1.f : 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = holder;
  {  // sentinel -1 in columns 0..47 of every lane
    const uint32_t s = __float_as_uint(-1.f);
    for (int c0 = 0; c0 < 48; c0 += 8) {
      const uint32_t taddr = tm + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(s), "r"(s), "r"(s),
                   "r"(s), "r"(s), "r"(s), "r"(s), "r"(s)
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    const uint64_t da = make_desc(smem_u32(sA), LBO_A), db = make_desc(smem_u32(sB), LBO_B);
    const uint32_t idesc = make_idesc(N, M);
    umma_tf32(tm, da, db, idesc, 0u);
    umma_tf32(tm + (16u << 16) + 16u, da, db, idesc, 0u);
    umma_tf32(tm + (64u << 16) + 32u, da, db, idesc, 0u);
    umma_commit(smem_u32(&bar));
  }
  while (!mbar_try_wait(smem_u32(&bar), 0)) {}
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int part = 0; part < 3; ++part) {
    uint32_t r[16];
    const uint32_t taddr = tm + ((uint32_t)(32 * warp) << 16) + 16u * part;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 16; ++c) out[(part * 128 + tid) * 16 + c] = __uint_as_float(r[c]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64u));
}

int main() {
  static float h[3 * 128 * 16];
  float* d;
  cudaMalloc(&d, sizeof h);
  k_probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  const char* names[3] = {"D lane field 0 ", "D lane field 16", "D lane field 64"};
  for (int part = 0; part < 3; ++part) {
    printf("%s: lane -> accumulator row (from column 0; '.' = untouched, '?' = other)\n  ", names[part]);
    int rows_seen = 0;
    for (int lane = 0; lane < 128; ++lane) {
      const float v = h[(part * 128 + lane) * 16 + 0], v1 = h[(part * 128 + lane) * 16 + 1];
      if (v == -1.f) printf(" .");
      else {
        const int r = (int)((v - 1.f) / 100.f);
        if (v == 100.f * r + 1.f && v1 == 100.f * r + 2.f) { printf(" %d", r); ++rows_seen; } else printf(" ?(%g)", v);
      }
      if (lane % 32 == 31) printf("\n  ");
    }
    printf("rows placed: %d\n", rows_seen);
  }
  return 0;
}
