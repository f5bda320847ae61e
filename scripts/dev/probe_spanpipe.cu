// Fourth copy probe: the memory side of a ROW-MAJOR epilogue for y_1 = y + C w3cat^T (DESIGN.md "what is next" (i)).
// One CTA per SM; per 128-row block the [R x 399] spans of `in` are pulled into shared memory by cp.async.bulk (contiguous,
// 16-byte aligned), eight warps add 1 in place (lane = row, pitch 399 words is odd: conflict free), and the span is pushed
// back by a bulk store.  Optionally a second thread streams 40 x 12960 B of an L2-resident weight image per block, the
// L2 -> SM traffic the tensor-core main loop of the real kernel adds.
//   build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I../../swarm_ode_b200/csrc probe_spanpipe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace gnode::tc;

constexpr int D = 399, TM = 128;

__device__ __forceinline__ void bulk_store_1d(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

template <int R, int NS, int WSTREAM>
__global__ void __launch_bounds__(320, 1) k_spanpipe(const float* __restrict__ in, float* __restrict__ out, long M,
                                                      const uint8_t* __restrict__ wimg, int* status) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full[NS], bar_done[NS], bar_free[NS], bar_w[3];
  constexpr int SLOT = R * D * 4;
  constexpr int WST = 12960;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sb = smem_u32(smem);
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_done[s]), 8); mbar_init(smem_u32(&bar_free[s]), 1); }
    for (int s = 0; s < 3; ++s) mbar_init(smem_u32(&bar_w[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long m_tiles = (M + TM - 1) / TM;
  constexpr int GPT = TM / R;   // row groups per tile
  if (warp == 0) {
    if (lane == 0) {            // loader
      uint32_t k = 0;
      for (long t = blockIdx.x; t < m_tiles; t += gridDim.x)
        for (int g = 0; g < GPT; ++g, ++k) {
          const uint32_t s = k % NS, lap = k / NS;
          if (lap > 0 && !mbar_wait(smem_u32(&bar_free[s]), (lap - 1) & 1u, status, 1)) return;
          const long r0 = t * TM + g * R;
          long rows = M - r0; if (rows > R) rows = R; if (rows <= 0) rows = 0;
          const uint32_t bytes = (uint32_t)(rows * D * 4);
          mbar_expect_tx(smem_u32(&bar_full[s]), bytes);
          if (bytes) bulk_load_1d(sb + s * SLOT, in + r0 * D, bytes, smem_u32(&bar_full[s]));
        }
    } else if (lane == 1 && WSTREAM) {   // weight stream: 40 stages per tile through a 3-slot ring
      uint32_t k = 0;
      for (long t = blockIdx.x; t < m_tiles; t += gridDim.x)
        for (int i = 0; i < 40; ++i, ++k) {
          const uint32_t s = k % 3, lap = k / 3;
          if (lap > 0 && !mbar_wait(smem_u32(&bar_w[s]), (lap - 1) & 1u, status, 2)) return;
          mbar_expect_tx(smem_u32(&bar_w[s]), WST);
          bulk_load_1d(sb + NS * SLOT + s * WST, wimg + (size_t)i * WST, WST, smem_u32(&bar_w[s]));
        }
    }
  } else if (warp == 1) {
    if (lane == 0) {            // storer
      uint32_t k = 0;
      for (long t = blockIdx.x; t < m_tiles; t += gridDim.x)
        for (int g = 0; g < GPT; ++g, ++k) {
          const uint32_t s = k % NS, lap = k / NS;
          if (!mbar_wait(smem_u32(&bar_done[s]), lap & 1u, status, 3)) return;
          const long r0 = t * TM + g * R;
          long rows = M - r0; if (rows > R) rows = R; if (rows <= 0) rows = 0;
          const uint32_t bytes = (uint32_t)(rows * D * 4);
          if (bytes) bulk_store_1d(out + r0 * D, sb + s * SLOT, bytes);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          if (k > 0) {          // the previous store has finished reading its slot
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            mbar_arrive(smem_u32(&bar_free[(k - 1) % NS]));
          }
        }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else {
    const int w = warp - 2;     // 8 workers: R rows x 399 columns, lane = row (R = 32) , warps split the columns
    uint32_t k = 0;
    for (long t = blockIdx.x; t < m_tiles; t += gridDim.x)
      for (int g = 0; g < GPT; ++g, ++k) {
        const uint32_t s = k % NS, lap = k / NS;
        if (!mbar_wait(smem_u32(&bar_full[s]), lap & 1u, status, 4)) return;
        float* slot = reinterpret_cast<float*>(smem + s * SLOT);
        const int row = lane % R;
        if (lane < R) {
          for (int c = w; c < D; c += 8) slot[row * D + c] += 1.f;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_done[s]));
      }
  }
}

int main() {
  const long M = 389120;
  float *in, *out; uint8_t* wimg; int* status;
  cudaMalloc(&in, M * D * 4); cudaMalloc(&out, M * D * 4); cudaMalloc(&wimg, 40 * 12960); cudaMalloc(&status, 4);
  cudaMemset(in, 0, M * D * 4); cudaMemset(wimg, 0, 40 * 12960); cudaMemset(status, 0, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double gb = 2.0 * M * D * 4 / 1e9;
  auto run = [&](const char* name, auto kern, int smem) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int it = 0; it < 3; ++it) kern<<<148, 320, smem>>>(in, out, M, wimg, status);
    cudaEventRecord(e0);
    for (int it = 0; it < 10; ++it) kern<<<148, 320, smem>>>(in, out, M, wimg, status);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int st; cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost);
    float chk; cudaMemcpy(&chk, out + (M - 1) * D + 398, 4, cudaMemcpyDeviceToHost);
    printf("%-58s %.3f ms  %5.0f GB/s  status %d  out[last]=%g  %s\n", name, ms / 10, gb / (ms / 10 / 1e3), st, chk, cudaGetErrorString(cudaGetLastError()));
  };
  run("32-row spans (51 KB) x 2 slots", k_spanpipe<32, 2, 0>, 2 * 32 * D * 4 + 3 * 12960);
  run("32-row spans x 2 slots + weight stream", k_spanpipe<32, 2, 1>, 2 * 32 * D * 4 + 3 * 12960);
  run("32-row spans x 3 slots", k_spanpipe<32, 3, 0>, 3 * 32 * D * 4 + 3 * 12960);
  run("32-row spans x 4 slots", k_spanpipe<32, 4, 0>, 4 * 32 * D * 4 + 3 * 12960);
  run("16-row spans (25 KB) x 4 slots", k_spanpipe<16, 4, 0>, 4 * 16 * D * 4 + 3 * 12960);
  run("16-row spans x 4 slots + weight stream", k_spanpipe<16, 4, 1>, 4 * 16 * D * 4 + 3 * 12960);
  run("16-row spans x 8 slots", k_spanpipe<16, 8, 0>, 8 * 16 * D * 4 + 3 * 12960);
  run("8-row spans (12.8 KB) x 8 slots", k_spanpipe<8, 8, 0>, 8 * 8 * D * 4 + 3 * 12960);
  run("8-row spans x 16 slots", k_spanpipe<8, 16, 0>, 16 * 8 * D * 4 + 3 * 12960);
  return 0;
}
