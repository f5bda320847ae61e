// Hardware probe for the chain kernels (sm_100a): (1) how tcgen05.mma kind::tf32 treats the 13 low mantissa bits of
// an fp32 container (truncate vs round), (2) the TMEM layout of a bf16 A operand (TS form of kind::f16).
// build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I swarm_ode_b200/csrc scripts/dev/probe_umma.cu -o gpurun_out/probe_umma
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_bf16.h>
#include "tc_common.cuh"
using namespace gnode::tc;

constexpr int M = 128, N = 16;
constexpr int LBO_A = M * 16, LBO_B = N * 16;

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}

// out1[M][N]: A(fp32 containers, K = 8) x B(tf32 "identity" rows: B[n][k] = (n == k)), i.e. out1[m][n] = tf32(A[m][n]) for n < 8
// out2[M][N]: A_bf16 (TMEM, K = 16) x B_bf16 identity(16): out2[m][n] = A2[m][n]
__global__ void k_probe(const float* A, const float* A2, float* out1, float* out2) {
  __shared__ __align__(128) uint8_t sA[2 * LBO_A];
  __shared__ __align__(128) uint8_t sB[2 * LBO_B];
  __shared__ __align__(128) uint8_t sB2[2 * LBO_B];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t holder;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&holder)), "r"(64u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  // A: row tid, 8 k -> two chunks
  for (int k = 0; k < 8; ++k)
    *reinterpret_cast<float*>(sA + (k >> 2) * LBO_A + (tid >> 3) * 128 + (tid & 7) * 16 + (k & 3) * 4) = A[tid * 8 + k];
  if (tid < N) {
    for (int k = 0; k < 8; ++k)
      *reinterpret_cast<float*>(sB + (k >> 2) * LBO_B + (tid >> 3) * 128 + (tid & 7) * 16 + (k & 3) * 4) = (tid == k) ? 1.f : 0.f;
    for (int k = 0; k < 16; ++k)
      *reinterpret_cast<__nv_bfloat16*>(sB2 + (k >> 3) * LBO_B + (tid >> 3) * 128 + (tid & 7) * 16 + (k & 7) * 2) =
          __float2bfloat16((tid == k) ? 1.f : 0.f);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = holder;
  // bf16 A operand -> TMEM columns 32..39 of this thread's lane: column c holds k = 2c (low half), 2c + 1 (high half)
  {
    uint32_t r[8];
    for (int c = 0; c < 8; ++c) {
      __nv_bfloat162 p = __floats2bfloat162_rn(A2[tid * 16 + 2 * c], A2[tid * 16 + 2 * c + 1]);
      r[c] = *reinterpret_cast<uint32_t*>(&p);
    }
    const uint32_t taddr = tm + ((uint32_t)(32 * warp) << 16) + 32u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    const uint64_t da = make_desc(smem_u32(sA), LBO_A), db = make_desc(smem_u32(sB), LBO_B), db2 = make_desc(smem_u32(sB2), LBO_B);
    umma_tf32(tm, da, db, make_idesc(N), 0u);
    const uint32_t idesc_bf = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    umma_bf16_ts(tm + 16u, tm + 32u, db2, idesc_bf, 0u);
    umma_commit(smem_u32(&bar));
  }
  while (!mbar_try_wait(smem_u32(&bar), 0)) {}
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int half = 0; half < 2; ++half) {
    uint32_t r[16];
    const uint32_t taddr = tm + ((uint32_t)(32 * warp) << 16) + 16u * half;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    float* o = half == 0 ? out1 : out2;
    for (int c = 0; c < 16; ++c) o[tid * 16 + c] = __uint_as_float(r[c]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64u));
}

int main() {
  float hA[M * 8], hA2[M * 16], h1[M * N], h2[M * N];
  srand(1);
  for (int i = 0; i < M * 8; ++i) hA[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (int i = 0; i < M * 16; ++i) hA2[i] = (float)((i * 7) % 251 - 125);   // bf16-exact integers
  float *dA, *dA2, *d1, *d2;
  cudaMalloc(&dA, sizeof hA); cudaMalloc(&dA2, sizeof hA2); cudaMalloc(&d1, sizeof h1); cudaMalloc(&d2, sizeof h2);
  cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dA2, hA2, sizeof hA2, cudaMemcpyHostToDevice);
  k_probe<<<1, 128>>>(dA, dA2, d1, d2);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  cudaMemcpy(h1, d1, sizeof h1, cudaMemcpyDeviceToHost); cudaMemcpy(h2, d2, sizeof h2, cudaMemcpyDeviceToHost);
  int n_trunc = 0, n_rna = 0, n_other = 0, n_diff = 0;
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < 8; ++k) {
      uint32_t u; memcpy(&u, &hA[m * 8 + k], 4);
      uint32_t t = u & 0xFFFFE000u, r = (u + 0x1000u) & 0xFFFFE000u;
      uint32_t g; memcpy(&g, &h1[m * N + k], 4);
      if (t != r) { ++n_diff; if (g == t) ++n_trunc; else if (g == r) ++n_rna; else ++n_other; }
      else if (g != t) ++n_other;
    }
  printf("tf32 operand: of %d containers where truncation and rounding differ: trunc=%d rna=%d; other=%d\n", n_diff, n_trunc, n_rna, n_other);
  int bad = 0;
  for (int m = 0; m < M; ++m) for (int k = 0; k < 16; ++k) if (h2[m * N + k] != hA2[m * 16 + k]) ++bad;
  printf("bf16 TS operand (lane = row, column c = k 2c | 2c+1): mismatches=%d  sample row1: %g %g %g %g (want %g %g %g %g)\n", bad,
         h2[16], h2[17], h2[18], h2[19], hA2[16], hA2[17], hA2[18], hA2[19]);
  return 0;
}
