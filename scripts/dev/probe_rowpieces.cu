// Bandwidth probe for DESIGN.md 4.4 / "what is next" (i): out = in + 1 over a [M x 399] fp32 matrix (1596-byte rows),
// (A) in 40-column pieces per 128-row block, all rows of the block per pass (the access pattern of a column-chunked GEMM
// epilogue), (B) the same block as one contiguous span.  build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a probe_rowpieces.cu
#include <cstdio>
#include <type_traits>
#include <cuda_runtime.h>

constexpr int D = 399, ROWS = 128;

template <int piece>
__global__ void __launch_bounds__(256, 4) k_pieces(const float* __restrict__ in, float* __restrict__ out, long M) {
  for (long blk = blockIdx.x; blk * ROWS < M; blk += gridDim.x) {
    const long m0 = blk * ROWS;
    const int nr = (int)((M - m0 < ROWS) ? (M - m0) : ROWS);
    for (int n0 = 0; n0 < D; n0 += piece) {
      const int total = nr * piece;
      for (int base = 0; base < total; base += 20 * 256) {
        float v[20];
#pragma unroll
        for (int u = 0; u < 20; ++u) {
          const int idx = base + threadIdx.x + u * 256, row = idx / piece, col = n0 + idx % piece;
          v[u] = (idx < total && col < D) ? __ldg(in + (m0 + row) * D + col) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 20; ++u) {
          const int idx = base + threadIdx.x + u * 256, row = idx / piece, col = n0 + idx % piece;
          if (idx < total && col < D) out[(m0 + row) * D + col] = v[u] + 1.f;
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(256, 4) k_span(const float* __restrict__ in, float* __restrict__ out, long M) {
  for (long blk = blockIdx.x; blk * ROWS < M; blk += gridDim.x) {
    const long e0 = blk * ROWS * D;
    const long e1 = (blk * ROWS + ROWS < M ? blk * ROWS + ROWS : M) * D;
    for (long i = e0 + threadIdx.x * 4L; i < e1; i += 256 * 4 * 5) {
      float4 v[5];
#pragma unroll
      for (int u = 0; u < 5; ++u) { const long j = i + u * 1024L; v[u] = j + 3 < e1 ? *reinterpret_cast<const float4*>(in + j) : make_float4(0, 0, 0, 0); }
#pragma unroll
      for (int u = 0; u < 5; ++u) { const long j = i + u * 1024L; if (j + 3 < e1) { v[u].x += 1.f; v[u].y += 1.f; v[u].z += 1.f; v[u].w += 1.f; *reinterpret_cast<float4*>(out + j) = v[u]; } }
    }
  }
}

int main() {
  const long M = 389120;
  float *in, *out;
  cudaMalloc(&in, M * D * 4); cudaMalloc(&out, M * D * 4);
  cudaMemset(in, 0, M * D * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double gb = 2.0 * M * D * 4 / 1e9;
  auto run = [&](auto tag) {
    constexpr int piece = decltype(tag)::value;
    for (int it = 0; it < 3; ++it) k_pieces<piece><<<148 * 4, 256>>>(in, out, M);
    cudaEventRecord(e0);
    for (int it = 0; it < 10; ++it) k_pieces<piece><<<148 * 4, 256>>>(in, out, M);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("pieces of %3d columns (%4d bytes at a 1596-byte pitch): %.3f ms  %.0f GB/s\n", piece, piece * 4, ms / 10, gb / (ms / 10 / 1e3));
  };
  run(std::integral_constant<int, 32>{}); run(std::integral_constant<int, 40>{}); run(std::integral_constant<int, 80>{});
  run(std::integral_constant<int, 200>{}); run(std::integral_constant<int, 399>{});
  for (int it = 0; it < 3; ++it) k_span<<<148 * 4, 256>>>(in, out, M);
  cudaEventRecord(e0);
  for (int it = 0; it < 10; ++it) k_span<<<148 * 4, 256>>>(in, out, M);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("contiguous 128-row spans (float4):                         %.3f ms  %.0f GB/s   %s\n", ms / 10, gb / (ms / 10 / 1e3), cudaGetErrorString(cudaGetLastError()));
  return 0;
}
