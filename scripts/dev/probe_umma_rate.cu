// Hardware probe (sm_100a): issue rate of tcgen05.mma 128 x N x 8 (kind::tf32) / 128 x N x 16 (kind::f16) as a function of
// where and how the operands live -- the question behind the chain kernels' 75 .. 95 ns per MMA:
//   mode 0  A, B in shared memory, K-major NO swizzle (the chain kernels' layout: 16-byte chunks, LBO between chunks)
//   mode 1  A, B in shared memory, K-major SWIZZLE_128B
//   mode 2  A in TENSOR MEMORY (TS form), B shared no swizzle
//   mode 3  A in tensor memory, B shared SWIZZLE_128B
//   mode 4  A, B shared, MN-major no swizzle (the tile read "transposed": the weight-gradient form)
// The issuing lane is chosen by elect.sync (a lane test makes the compiler wrap every MMA in a ~180-cycle waterfall, which is
// what the first version of this probe measured: 185 .. 230 cycles whatever the mode or N).
// Operand contents are zeros: only time is measured (clock64 around `iters` back-to-back MMAs + commit + wait).
// build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I swarm_ode_b200/csrc scripts/dev/probe_umma_rate.cu -o gpurun_out/probe_umma_rate
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"
using namespace gnode::tc;

constexpr int SMEM = 160 * 1024;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {          // K-major, 128-byte swizzle: SBO = 1024 (8 rows x 128 B)
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                                                  // LBO unused for swizzled K-major
  d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                                                  // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint64_t desc_mn_none(uint32_t saddr, uint32_t sbo_bytes, uint32_t lbo_bytes) {   // MN-major, no swizzle
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) k_rate(int mode, int n, int iters, int kblocks, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t holder;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < SMEM / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&holder)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = holder;
  const uint32_t a_base = smem_u32(smem), b_base = a_base + 80 * 1024;
  if (warp == 1 && elect_one()) {
    const uint32_t lbo_a = 96 * 16 + 16, lbo_b = (uint32_t)n * 16 + 16;          // the chain kernels' chunk pitches
    const uint32_t idesc = make_idesc(n) | (mode == 4 ? ((1u << 15) | (1u << 16)) : 0u);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      #pragma unroll
      for (int kb = 0; kb < 16; ++kb) {                                       // one MMA per K = 8 step, walking the operands
        uint64_t da, db;
        if (mode == 0 || mode == 2) {
          da = make_desc(a_base + (uint32_t)(2 * kb) * lbo_a, lbo_a);
          db = make_desc(b_base + (uint32_t)(2 * kb) * lbo_b, lbo_b);
        } else if (mode == 1 || mode == 3) {
          da = desc_sw128(a_base + (uint32_t)(kb >> 2) * 16384 + (uint32_t)(kb & 3) * 32);   // 128 rows x 128 B per K = 32 group
          db = desc_sw128(b_base + (uint32_t)(kb >> 2) * (uint32_t)n * 128 + (uint32_t)(kb & 3) * 32);
        } else {
          da = desc_mn_none(a_base + (uint32_t)kb * 128, lbo_a, 128);                // M chunks lbo_a apart, K groups of 8 rows 128 B apart
          db = desc_mn_none(b_base + (uint32_t)kb * 128, lbo_a, 128);
        }
        if (mode == 2 || mode == 3) umma_tf32_ts(tm, tm + 256u + (uint32_t)(8 * kb), db, idesc, (it | kb) ? 1u : 0u);
        else umma_tf32(tm, da, db, idesc, (it | kb) ? 1u : 0u);
      }
    }
    umma_commit(smem_u32(&bar));
    while (!mbar_try_wait(smem_u32(&bar), 0)) {}
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u));
}

int main() {
  long long* d; long long h;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  const char* names[5] = {"SS K-major no swizzle", "SS K-major SWIZZLE_128B", "TS (A in TMEM) + B no swizzle", "TS + B SWIZZLE_128B", "SS MN-major no swizzle"};
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  for (int grid : {1, 148}) {
    for (int n : {128, 64}) {
      for (int mode = 0; mode < 5; ++mode) {
        const int iters = 64, kblocks = 16;
        k_rate<<<grid, 128, SMEM>>>(mode, n, 4, kblocks, d);      // warm
        k_rate<<<grid, 128, SMEM>>>(mode, n, iters, kblocks, d);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        const double cyc = (double)h / (iters * kblocks);
        printf("grid %3d  N=%3d  %-32s %7.1f cycles / MMA (128 x %d x 8 tf32; floor %d)  [%s]\n", grid, n, names[mode], cyc, n, 128 * n / 256,
               cudaGetErrorString(e));
      }
    }
  }
  return 0;
}
