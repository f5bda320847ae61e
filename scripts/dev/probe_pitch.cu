// Third copy probe: is the slow column-chunked WRITE to a [M x 399] fp32 matrix (probe_l2prefetch.cu: 2.5 TB/s against
// 6 TB/s for whole rows) a matter of sector / line alignment of the pieces (pitch 1596 B is only 4-byte aligned) or of the
// access pattern itself?  Same kernel, pitch and piece width as template parameters.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ROWS = 128;

template <int piece, int PITCH, int NCOL, int MODE>
__global__ void __launch_bounds__(256, 4) k_pieces(const float* __restrict__ in, float* __restrict__ out, long M) {
  float sink = 0.f;
  const long nblk = (M + ROWS - 1) / ROWS;
  for (long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const long m0 = blk * ROWS;
    const int nr = (int)((M - m0 < ROWS) ? (M - m0) : ROWS);
    for (int n0 = 0; n0 < NCOL; n0 += piece) {
      const int total = nr * piece;
      for (int base = 0; base < total; base += 20 * 256) {
        float v[20];
#pragma unroll
        for (int u = 0; u < 20; ++u) {
          const int idx = base + threadIdx.x + u * 256, row = idx / piece, col = n0 + idx % piece;
          if (MODE & 1) v[u] = (idx < total && col < NCOL) ? __ldg(in + (m0 + row) * PITCH + col) : 0.f;
          else v[u] = (float)idx;
        }
#pragma unroll
        for (int u = 0; u < 20; ++u) {
          const int idx = base + threadIdx.x + u * 256, row = idx / piece, col = n0 + idx % piece;
          if (MODE & 2) { if (idx < total && col < NCOL) out[(m0 + row) * PITCH + col] = v[u] + 1.f; }
          else sink += v[u];
        }
      }
      __syncthreads();
    }
  }
  if (!(MODE & 2) && sink == 12345.678f) out[threadIdx.x] = sink;
}

// the epilogue shape of a tensor-core GEMM: thread = row, 32 consecutive columns per pass (TMEM lane = row), float4 when aligned
template <int PITCH, int NCOL, int CH>
__global__ void __launch_bounds__(128, 8) k_lane_rows(float* __restrict__ out, long M) {
  const long nblk = (M + ROWS - 1) / ROWS;
  for (long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const long r = blk * ROWS + threadIdx.x;
    if (r >= M) continue;
    float* o = out + r * PITCH;
    for (int n0 = 0; n0 < NCOL; n0 += CH) {
      if (PITCH % 4 == 0) {
#pragma unroll
        for (int j = 0; j < CH; j += 4) if (n0 + j + 3 < NCOL) *reinterpret_cast<float4*>(o + n0 + j) = make_float4(1.f, 2.f, 3.f, (float)j);
      } else {
#pragma unroll
        for (int j = 0; j < CH; ++j) if (n0 + j < NCOL) o[n0 + j] = (float)j;
      }
    }
  }
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
  printf("%-70s %.3f ms  %5.0f GB/s  %s\n", name, ms / 10, gb / (ms / 10 / 1e3), cudaGetErrorString(cudaGetLastError()));
}

int main() {
  cudaMalloc(&in_, M * 512 * 4); cudaMalloc(&out_, M * 512 * 4);
  cudaMemset(in_, 0, M * 512 * 4);
  cudaEventCreate(&e0); cudaEventCreate(&e1);
#define P(piece, pitch, ncol, mode, label) \
  timeit(label, (double)M * ncol * 4 / 1e9 * ((mode) == 3 ? 2 : 1), [&] { k_pieces<piece, pitch, ncol, mode><<<148 * 4, 256>>>(in_, out_, M); })
  P(80, 399, 399, 2, "write  pitch 399  80-col pieces");
  P(80, 400, 399, 2, "write  pitch 400  80-col pieces (399 cols used)");
  P(80, 400, 400, 2, "write  pitch 400  80-col pieces (400 cols: rows abut)");
  P(100, 400, 400, 2, "write  pitch 400  100-col pieces");
  P(200, 400, 400, 2, "write  pitch 400  200-col pieces");
  P(400, 400, 400, 2, "write  pitch 400  whole rows");
  P(32, 416, 399, 2, "write  pitch 416  32-col pieces (line aligned)");
  P(96, 416, 399, 2, "write  pitch 416  96-col pieces (line aligned)");
  P(128, 416, 399, 2, "write  pitch 416  128-col pieces (line aligned)");
  P(128, 512, 512, 2, "write  pitch 512  128-col pieces, 512 cols");
  P(512, 512, 512, 2, "write  pitch 512  whole rows");
  P(128, 128, 128, 2, "write  pitch 128  whole rows (the 2H-wide tensors)");
  P(64, 128, 128, 2, "write  pitch 128  64-col pieces");
  P(80, 400, 400, 3, "r+w    pitch 400  80-col pieces");
  P(80, 399, 399, 3, "r+w    pitch 399  80-col pieces");
  P(16, 400, 400, 1, "read   pitch 400  16-col pieces");
  P(16, 399, 399, 1, "read   pitch 399  16-col pieces");
  P(32, 400, 400, 1, "read   pitch 400  32-col pieces");
  timeit("lane=row  pitch 399  32-col passes (scalar)", (double)M * 399 * 4 / 1e9, [&] { k_lane_rows<399, 399, 32><<<148 * 8, 128>>>(out_, M); });
  timeit("lane=row  pitch 400  32-col passes (float4)", (double)M * 400 * 4 / 1e9, [&] { k_lane_rows<400, 400, 32><<<148 * 8, 128>>>(out_, M); });
  timeit("lane=row  pitch 400  80-col passes (float4)", (double)M * 400 * 4 / 1e9, [&] { k_lane_rows<400, 400, 80><<<148 * 8, 128>>>(out_, M); });
  timeit("lane=row  pitch 400  400-col pass  (float4)", (double)M * 400 * 4 / 1e9, [&] { k_lane_rows<400, 400, 400><<<148 * 8, 128>>>(out_, M); });
  return 0;
}
