// Lossless narrow transport of a host batch (the reference moves a collated fp32 / int64 PyG Batch with
// `batch.to(device)`, scripts/train_gde.py:475).  Warehouse observations are small integers (flags, grid coordinates,
// SURVEY 8-d2), so a batch travels over PCIe in the narrowest type that reproduces every value exactly (checked on the
// host when the batch is packed, swarm_ode_b200/data.py:PackedBatch) and is widened here, on the device, to the fp32 /
// int64 tensors the reference contract names: bit-exact `x`, `edge_index`, `batch`.
#include "common.cuh"

#include <cuda_fp16.h>

namespace gnode {
namespace {

template <int KIND>
__device__ __forceinline__ float widen_one(const void* src, int64_t i) {
  if (KIND == GNODE_PACK_U8) return (float)static_cast<const uint8_t*>(src)[i];
  if (KIND == GNODE_PACK_I16) return (float)static_cast<const int16_t*>(src)[i];
  return __half2float(static_cast<const __half*>(src)[i]);
}

// 16 elements per thread and iteration: one 16-byte (u8) or two 16-byte (i16 / f16) loads, four 16-byte stores
template <int KIND>
__global__ void __launch_bounds__(256) k_unpack_features(const void* __restrict__ src, int64_t n, float* __restrict__ dst) {
  const int64_t n16 = n / 16;
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n16; v += (int64_t)gridDim.x * blockDim.x) {
    float o[16];
    if (KIND == GNODE_PACK_U8) {
      const uint4 q = __ldcs(static_cast<const uint4*>(src) + v);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        o[4 * j + 0] = (float)(w[j] & 0xFFu); o[4 * j + 1] = (float)((w[j] >> 8) & 0xFFu);
        o[4 * j + 2] = (float)((w[j] >> 16) & 0xFFu); o[4 * j + 3] = (float)(w[j] >> 24);
      }
    } else {
      const uint4 q0 = __ldcs(static_cast<const uint4*>(src) + 2 * v), q1 = __ldcs(static_cast<const uint4*>(src) + 2 * v + 1);
      const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (KIND == GNODE_PACK_I16) {
          o[2 * j] = (float)(int16_t)(w[j] & 0xFFFFu); o[2 * j + 1] = (float)(int16_t)(w[j] >> 16);
        } else {
          const __half2 h = *reinterpret_cast<const __half2*>(&w[j]);
          o[2 * j] = __low2float(h); o[2 * j + 1] = __high2float(h);
        }
      }
    }
    float4* d = reinterpret_cast<float4*>(dst) + 4 * v;
#pragma unroll
    for (int j = 0; j < 4; ++j) d[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
  }
  // ragged tail (< 16 elements)
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - 16 * n16)) {
    const int64_t i = 16 * n16 + threadIdx.x;
    dst[i] = widen_one<KIND>(src, i);
  }
}

__global__ void __launch_bounds__(256) k_unpack_edges(const int32_t* __restrict__ src, int64_t n, int64_t* __restrict__ dst) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = (int64_t)src[i];
}

// batch[i] = graph of node i: the last g with ptr[g] <= i (PyG `Batch.batch`; empty graphs own no node)
__global__ void __launch_bounds__(256) k_batch_vector(const int64_t* __restrict__ ptr, int64_t n_graphs, int64_t n_nodes,
                                                      int64_t* __restrict__ batch) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_nodes; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t lo = 0, hi = n_graphs;            // invariant: ptr[lo] <= i < ptr[hi]
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (__ldg(ptr + mid) <= i) lo = mid; else hi = mid;
    }
    batch[i] = lo;
  }
}

// Column-wise bit packing (GNODE_PACK_BITS): every column c of the [rows, cols] matrix holds non-negative integers below
// 2^w_c (w_c <= 8, chosen per column when the batch is packed); a row is the little-endian bit string of its columns,
// column c at bits [off[c], off[c + 1]), padded to row_bytes (a multiple of 16, at least one byte beyond the last bit).
// Warehouse observations are mostly flags (1 bit) and grid coordinates (5 bits): 105 bytes per 399-column row instead of
// 399 (u8) or 1596 (fp32).  A block stages UNPACK_ROWS rows in shared memory (16-byte loads), then decodes (row, column)
// pairs with consecutive threads on consecutive columns: coalesced fp32 stores.
constexpr int UNPACK_ROWS = 32, UNPACK_MAX_COLS = 2048, UNPACK_MAX_ROW_BYTES = 1024;
__global__ void __launch_bounds__(256) k_unpack_bits(const uint8_t* __restrict__ src, int64_t rows, int cols, int row_bytes,
                                                     const int32_t* __restrict__ off, float* __restrict__ dst) {
  extern __shared__ __align__(16) uint8_t ub_smem[];
  int32_t* s_off = reinterpret_cast<int32_t*>(ub_smem);                       // [cols + 1]
  uint8_t* s_rows = ub_smem + (((size_t)(cols + 1) * 4 + 15) & ~(size_t)15);  // [UNPACK_ROWS][row_bytes]
  for (int i = threadIdx.x; i <= cols; i += blockDim.x) s_off[i] = off[i];
  const int64_t n_chunks = (rows + UNPACK_ROWS - 1) / UNPACK_ROWS;
  for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const int64_t r0 = ch * UNPACK_ROWS;
    const int nr = (int)((rows - r0 < UNPACK_ROWS) ? (rows - r0) : UNPACK_ROWS);
    __syncthreads();                                                           // table loaded / previous chunk decoded
    const uint4* g = reinterpret_cast<const uint4*>(src + (size_t)r0 * row_bytes);
    const int n16 = nr * (row_bytes / 16);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) reinterpret_cast<uint4*>(s_rows)[i] = __ldcs(g + i);
    __syncthreads();
    const int total = nr * cols;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int r = i / cols, c = i - r * cols;
      const int o = s_off[c], w = s_off[c + 1] - o;
      const uint8_t* p = s_rows + (size_t)r * row_bytes + (o >> 3);
      const uint32_t v = (((uint32_t)p[0] | ((uint32_t)p[1] << 8)) >> (o & 7)) & ((1u << w) - 1u);
      dst[(size_t)(r0 + r) * cols + c] = (float)v;
    }
  }
}

unsigned grid_for(int64_t work) {
  int64_t b = ceil_div64(work, 256);
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace
}  // namespace gnode

using namespace gnode;

extern "C" int gnode_unpack_features(const void* src, int32_t kind, int64_t n, float* dst, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_ARG(n >= 0 && (n == 0 || (src && dst)), "gnode_unpack_features: bad argument");
  GN_ARG(kind >= GNODE_PACK_U8 && kind <= GNODE_PACK_F32, "gnode_unpack_features: kind must be 0 (u8), 1 (i16), 2 (f16) or 3 (f32)");
  GN_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
         "gnode_unpack_features: buffers must be 16-byte aligned");
  if (n == 0) return GNODE_OK;
  if (kind == GNODE_PACK_F32) {
    if (src != dst) GN_CUDA(cudaMemcpyAsync(dst, src, sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, s));
    return GNODE_OK;
  }
  const double in_bytes = (kind == GNODE_PACK_U8 ? 1.0 : 2.0) * (double)n;
  GN_PROF(s, 0.0, in_bytes + 4.0 * (double)n, "unpack_features kind=%d", kind);
  const unsigned grid = grid_for(n / 16 + 1);
  if (kind == GNODE_PACK_U8) k_unpack_features<GNODE_PACK_U8><<<grid, 256, 0, s>>>(src, n, dst);
  else if (kind == GNODE_PACK_I16) k_unpack_features<GNODE_PACK_I16><<<grid, 256, 0, s>>>(src, n, dst);
  else k_unpack_features<GNODE_PACK_F16><<<grid, 256, 0, s>>>(src, n, dst);
  GN_LAUNCHED();
  return GNODE_OK;
}

extern "C" int gnode_unpack_bits(const void* src, int64_t rows, int32_t cols, int32_t row_bytes, const int32_t* bit_offsets,
                                 float* dst, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_ARG(rows >= 0 && cols >= 1 && cols <= UNPACK_MAX_COLS && (rows == 0 || (src && dst && bit_offsets)), "gnode_unpack_bits: bad argument");
  GN_ARG(row_bytes >= 16 && row_bytes % 16 == 0 && row_bytes <= UNPACK_MAX_ROW_BYTES, "gnode_unpack_bits: row_bytes must be a multiple of 16 in [16, %d]",
         UNPACK_MAX_ROW_BYTES);
  GN_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0, "gnode_unpack_bits: the packed buffer must be 16-byte aligned");
  if (rows == 0) return GNODE_OK;
  GN_PROF(s, 0.0, (double)rows * row_bytes + 4.0 * (double)rows * cols, "unpack_features kind=bits");
  const size_t smem = (((size_t)(cols + 1) * 4 + 15) & ~(size_t)15) + (size_t)UNPACK_ROWS * row_bytes;
  if (first_use_on_device(reinterpret_cast<const void*>(&k_unpack_bits))) {
    GN_CUDA(cudaFuncSetAttribute(k_unpack_bits, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)((((size_t)(UNPACK_MAX_COLS + 1) * 4 + 15) & ~(size_t)15) + (size_t)UNPACK_ROWS * UNPACK_MAX_ROW_BYTES)));
  }
  const int64_t chunks = (rows + UNPACK_ROWS - 1) / UNPACK_ROWS;
  const int64_t cap = (int64_t)kNumSMs * 8;
  k_unpack_bits<<<(unsigned)(chunks < cap ? chunks : cap), 256, smem, s>>>(static_cast<const uint8_t*>(src), rows, cols, row_bytes,
                                                                          bit_offsets, dst);
  GN_LAUNCHED();
  return GNODE_OK;
}

extern "C" int gnode_unpack_edges(const int32_t* src, int64_t n, int64_t* dst, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_ARG(n >= 0 && (n == 0 || (src && dst)), "gnode_unpack_edges: bad argument");
  if (n == 0) return GNODE_OK;
  GN_PROF(s, 0.0, 12.0 * (double)n, "unpack_edges");
  k_unpack_edges<<<grid_for(n), 256, 0, s>>>(src, n, dst);
  GN_LAUNCHED();
  return GNODE_OK;
}

extern "C" int gnode_batch_vector(const int64_t* ptr, int64_t n_graphs, int64_t n_nodes, int64_t* batch, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_ARG(n_graphs >= 1 && n_nodes >= 0 && ptr && (n_nodes == 0 || batch), "gnode_batch_vector: bad argument");
  if (n_nodes == 0) return GNODE_OK;
  GN_PROF(s, 0.0, 8.0 * (double)n_nodes, "batch_vector");
  k_batch_vector<<<grid_for(n_nodes), 256, 0, s>>>(ptr, n_graphs, n_nodes, batch);
  GN_LAUNCHED();
  return GNODE_OK;
}
