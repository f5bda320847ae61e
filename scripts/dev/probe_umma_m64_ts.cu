// Hardware probe (sm_100a), companion of probe_umma_m64.cu: with M = 64 and the A operand in TENSOR MEMORY (TS form of
// kind::f16), which TMEM lane feeds accumulator row r?  Every lane L holds the bf16 vector (L, 0, 0, ...); B is the identity,
// so D[r][0] = the lane that row r was read from.  Run with the A address lane field 0 and 16.
#include <cstdio>
#include <cuda_bf16.h>
#include "tc_common.cuh"
using namespace gnode::tc;

constexpr int M = 64, N = 16;
constexpr int LBO_B = N * 16;

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}

__global__ void k_probe(float* out) {
  __shared__ __align__(128) uint8_t sB[2 * LBO_B];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t holder;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&holder)), "r"(64u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid < N)
    for (int k = 0; k < 16; ++k)
      *reinterpret_cast<__nv_bfloat16*>(sB + (k >> 3) * LBO_B + (tid >> 3) * 128 + (tid & 7) * 16 + (k & 7) * 2) =
          __float2bfloat16((tid == k) ? 1.f : 0.f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = holder;
  {
    // D columns 0..31 <- -1; A operand columns 32..39: (lane, 0) in column 32, zeros elsewhere
    const uint32_t sent = __float_as_uint(-1.f);
    for (int c0 = 0; c0 < 32; c0 += 8) {
      const uint32_t taddr = tm + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(sent), "r"(sent),
                   "r"(sent), "r"(sent), "r"(sent), "r"(sent), "r"(sent), "r"(sent)
                   : "memory");
    }
    __nv_bfloat162 p = __floats2bfloat162_rn((float)tid, 0.f);
    const uint32_t a0 = *reinterpret_cast<uint32_t*>(&p);
    const uint32_t taddr = tm + ((uint32_t)(32 * warp) << 16) + 32u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(a0), "r"(0u), "r"(0u),
                 "r"(0u), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    const uint64_t db = make_desc(smem_u32(sB), LBO_B);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    umma_bf16_ts(tm, tm + 32u, db, idesc, 0u);                                  // D lanes 0.., A lane field 0
    umma_bf16_ts(tm + (16u << 16) + 16u, tm + (16u << 16) + 32u, db, idesc, 0u);  // D and A lane field 16
    umma_commit(smem_u32(&bar));
  }
  while (!mbar_try_wait(smem_u32(&bar), 0)) {}
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int part = 0; part < 2; ++part) {
    uint32_t r[16];
    const uint32_t taddr = tm + ((uint32_t)(32 * warp) << 16) + 16u * part;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    out[(part * 128 + tid) * 2 + 0] = __uint_as_float(r[0]);
    out[(part * 128 + tid) * 2 + 1] = __uint_as_float(r[1]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64u));
}

int main() {
  static float h[2 * 128 * 2];
  float* d;
  cudaMalloc(&d, sizeof h);
  k_probe<<<1, 128>>>(d);
  printf("kernel: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  for (int part = 0; part < 2; ++part) {
    printf("lane field %d: TMEM lane of D -> TMEM lane its A row was read from ('.' = D lane untouched)\n  ", part * 16);
    for (int lane = 0; lane < 128; ++lane) {
      const float v = h[(part * 128 + lane) * 2];
      if (v == -1.f) printf(" ."); else printf(" %g", v);
      if (lane % 32 == 31) printf("\n  ");
    }
    printf("\n");
  }
  return 0;
}
