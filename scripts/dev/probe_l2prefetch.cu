// Follow-up to probe_rowpieces.cu: does a bulk L2 prefetch of the contiguous [128 x 399] span (issued once per block, one
// block ahead) lift the ~2.2 TB/s ceiling of column-chunked access to a 1596-byte-pitch matrix?  Also separates the read
// side from the write side.   build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a probe_l2prefetch.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int D = 399, ROWS = 128;
constexpr int SPAN_BYTES = ROWS * D * 4;   // 204288, a multiple of 16

__device__ __forceinline__ void bulk_prefetch_l2(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// MODE bit0: read, bit1: write, bit2: prefetch the input span one block ahead, bit3: prefetch the OUTPUT span too (so the
// stores land on resident lines)
template <int piece, int MODE>
__global__ void __launch_bounds__(256, 4) k_pieces(const float* __restrict__ in, float* __restrict__ out, long M) {
  float sink = 0.f;
  const long nblk = (M + ROWS - 1) / ROWS;
  if ((MODE & 4) && threadIdx.x == 0 && blockIdx.x < nblk) {
    const long m0 = (long)blockIdx.x * ROWS;
    const long nr = (M - m0 < ROWS) ? (M - m0) : ROWS;
    bulk_prefetch_l2(in + m0 * D, (unsigned)(nr * D * 4) & ~15u);
  }
  for (long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const long m0 = blk * ROWS;
    const int nr = (int)((M - m0 < ROWS) ? (M - m0) : ROWS);
    if ((MODE & 4) && threadIdx.x == 0 && blk + gridDim.x < nblk) {
      const long m1 = (blk + gridDim.x) * ROWS;
      const long nr1 = (M - m1 < ROWS) ? (M - m1) : ROWS;
      bulk_prefetch_l2(in + m1 * D, (unsigned)(nr1 * D * 4) & ~15u);
    }
    if ((MODE & 8) && threadIdx.x == 32) bulk_prefetch_l2(out + m0 * D, (unsigned)((long)nr * D * 4) & ~15u);
    for (int n0 = 0; n0 < D; n0 += piece) {
      const int total = nr * piece;
      for (int base = 0; base < total; base += 20 * 256) {
        float v[20];
#pragma unroll
        for (int u = 0; u < 20; ++u) {
          const int idx = base + threadIdx.x + u * 256, row = idx / piece, col = n0 + idx % piece;
          if (MODE & 1) v[u] = (idx < total && col < D) ? __ldg(in + (m0 + row) * D + col) : 0.f;
          else v[u] = (float)idx;
        }
#pragma unroll
        for (int u = 0; u < 20; ++u) {
          const int idx = base + threadIdx.x + u * 256, row = idx / piece, col = n0 + idx % piece;
          if (MODE & 2) { if (idx < total && col < D) out[(m0 + row) * D + col] = v[u] + 1.f; }
          else sink += v[u];
        }
      }
      __syncthreads();
    }
  }
  if (!(MODE & 2) && sink == 12345.678f) out[threadIdx.x] = sink;
}

template <int MODE>
__global__ void __launch_bounds__(256, 4) k_span(const float* __restrict__ in, float* __restrict__ out, long M) {
  float sink = 0.f;
  for (long blk = blockIdx.x; blk * ROWS < M; blk += gridDim.x) {
    const long e0 = blk * ROWS * D;
    const long e1 = (blk * ROWS + ROWS < M ? blk * ROWS + ROWS : M) * D;
    for (long i = e0 + threadIdx.x * 4L; i < e1; i += 256 * 4 * 5) {
      float4 v[5];
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const long j = i + u * 1024L;
        v[u] = ((MODE & 1) && j + 3 < e1) ? *reinterpret_cast<const float4*>(in + j) : make_float4(0, 0, 0, (float)j);
      }
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const long j = i + u * 1024L;
        if (MODE & 2) { if (j + 3 < e1) { v[u].x += 1.f; v[u].y += 1.f; v[u].z += 1.f; v[u].w += 1.f; *reinterpret_cast<float4*>(out + j) = v[u]; } }
        else sink += v[u].x + v[u].y + v[u].z + v[u].w;
      }
    }
  }
  if (!(MODE & 2) && sink == 12345.678f) out[threadIdx.x] = sink;
}

static float *in_, *out_;
static cudaEvent_t e0, e1;
static const long M = 389120;

template <typename F>
static void timeit(const char* name, double gb, F launch) {
  for (int it = 0; it < 3; ++it) launch();
  cudaEventRecord(e0);
  for (int it = 0; it < 10; ++it) launch();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-64s %.3f ms  %5.0f GB/s  %s\n", name, ms / 10, gb / (ms / 10 / 1e3), cudaGetErrorString(cudaGetLastError()));
}

int main() {
  cudaMalloc(&in_, M * D * 4); cudaMalloc(&out_, M * D * 4);
  cudaMemset(in_, 0, M * D * 4);
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double one = (double)M * D * 4 / 1e9;
#define P(piece, mode, gbs, label) timeit(label, gbs, [&] { k_pieces<piece, mode><<<148 * 4, 256>>>(in_, out_, M); })
  P(80, 3, 2 * one, "80-col pieces  read+write");
  P(80, 7, 2 * one, "80-col pieces  read+write, input span prefetched to L2");
  P(80, 15, 2 * one, "80-col pieces  read+write, input and output spans prefetched");
  P(80, 1, one, "80-col pieces  read only");
  P(80, 5, one, "80-col pieces  read only, input span prefetched to L2");
  P(80, 2, one, "80-col pieces  write only");
  P(80, 10, one, "80-col pieces  write only, output span prefetched");
  P(16, 1, one, "16-col pieces (64 B, the K-block of an A operand) read only");
  P(16, 5, one, "16-col pieces  read only, input span prefetched to L2");
  P(32, 1, one, "32-col pieces  read only");
  P(32, 5, one, "32-col pieces  read only, input span prefetched to L2");
  P(399, 3, 2 * one, "whole rows     read+write");
  P(399, 1, one, "whole rows     read only");
  P(399, 2, one, "whole rows     write only");
  timeit("contiguous float4 spans read+write", 2 * one, [&] { k_span<3><<<148 * 4, 256>>>(in_, out_, M); });
  timeit("contiguous float4 spans read only", one, [&] { k_span<1><<<148 * 4, 256>>>(in_, out_, M); });
  timeit("contiguous float4 spans write only", one, [&] { k_span<2><<<148 * 4, 256>>>(in_, out_, M); });
  return 0;
}
