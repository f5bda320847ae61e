// CSR construction from a PyG-style edge_index (int64 [2, E], unsorted).
// Replaces the scatter-by-destination inside SAGEConv.propagate (reference call sites
// scripts/train_gde.py:36,39,43): instead of atomically scattering messages on every layer of every
// RK stage, edges are bucketed once per batch into destination-sorted CSR (forward gather) and
// source-sorted CSR (transposed gather for backward).  Rows are sorted by neighbour id so that all
// later floating-point reductions have a fixed order.
#include "common.cuh"

namespace gnode {
namespace {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanChunk = kScanThreads * kScanItems;  // 2048

__global__ void k_count(const int64_t* __restrict__ ei, int64_t E, int64_t N,
                        int32_t* __restrict__ cnt_in, int32_t* __restrict__ cnt_out,
                        int32_t* __restrict__ flag) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; e < E; e += stride) {
    int64_t s = ei[e], d = ei[E + e];
    if (s < 0 || s >= N || d < 0 || d >= N) {
      if (!(s == -1 && d == -1)) *flag = 1;    // (-1, -1) is the padding edge of fixed-capacity edge buffers: skipped silently
      continue;
    }
    atomicAdd(&cnt_in[d], 1);
    atomicAdd(&cnt_out[s], 1);
  }
}

// block-level inclusive scan helper (256 threads)
__device__ __forceinline__ int block_excl_scan(int v, int* total) {
  __shared__ int warp_sums[kScanThreads / 32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_sums[w] = x;
  __syncthreads();
  if (w == 0) {
    int s = (lane < kScanThreads / 32) ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += y;
    }
    if (lane < kScanThreads / 32) warp_sums[lane] = s;
  }
  __syncthreads();
  int prefix = (w > 0) ? warp_sums[w - 1] : 0;
  if (total) *total = warp_sums[kScanThreads / 32 - 1];
  __syncthreads();
  return prefix + x - v;
}

// (blockIdx.y = 0 / 1: the in-degree and the out-degree scan of a CSR build run in the same launches)
__global__ void k_chunk_sums(const int32_t* __restrict__ cnt0, const int32_t* __restrict__ cnt1, int64_t N,
                             int32_t* __restrict__ sums0, int32_t* __restrict__ sums1) {
  const int32_t* __restrict__ cnt = blockIdx.y ? cnt1 : cnt0;
  int32_t* __restrict__ sums = blockIdx.y ? sums1 : sums0;
  int64_t base = (int64_t)blockIdx.x * kScanChunk;
  int local = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t idx = base + (int64_t)threadIdx.x * kScanItems + i;
    if (idx < N) local += cnt[idx];
  }
  int total;
  block_excl_scan(local, &total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// single block: exclusive scan of the chunk sums in place
__global__ void k_scan_sums(int32_t* __restrict__ sums0, int32_t* __restrict__ sums1, int64_t nb) {
  int32_t* __restrict__ sums = blockIdx.y ? sums1 : sums0;
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < nb; base += kScanThreads) {
    int64_t idx = base + threadIdx.x;
    int v = (idx < nb) ? sums[idx] : 0;
    int total;
    int ex = block_excl_scan(v, &total);
    int carry = carry_s;
    if (idx < nb) sums[idx] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + total;
    __syncthreads();
  }
}

__global__ void k_scan_final(const int32_t* __restrict__ cnt0, const int32_t* __restrict__ cnt1, int64_t N,
                             const int32_t* __restrict__ sums0, const int32_t* __restrict__ sums1,
                             int32_t* __restrict__ rowptr0, int32_t* __restrict__ rowptr1) {
  const int32_t* __restrict__ cnt = blockIdx.y ? cnt1 : cnt0;
  const int32_t* __restrict__ sums = blockIdx.y ? sums1 : sums0;
  int32_t* __restrict__ rowptr = blockIdx.y ? rowptr1 : rowptr0;
  int64_t base = (int64_t)blockIdx.x * kScanChunk;
  int v[kScanItems];
  int local = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t idx = base + (int64_t)threadIdx.x * kScanItems + i;
    v[i] = (idx < N) ? cnt[idx] : 0;
    local += v[i];
  }
  int total;
  int ex = block_excl_scan(local, &total) + sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t idx = base + (int64_t)threadIdx.x * kScanItems + i;
    if (idx < N) rowptr[idx] = ex;
    ex += v[i];
    if (idx == N - 1) rowptr[N] = ex;
  }
}

__global__ void k_fill(const int64_t* __restrict__ ei, int64_t E, int64_t N,
                       const int32_t* __restrict__ rowptr, const int32_t* __restrict__ t_rowptr,
                       int32_t* __restrict__ cur_in, int32_t* __restrict__ cur_out,
                       int32_t* __restrict__ col_u, int32_t* __restrict__ t_col_u) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; e < E; e += stride) {
    int64_t s = ei[e], d = ei[E + e];
    if (s < 0 || s >= N || d < 0 || d >= N) continue;
    int p = atomicAdd(&cur_in[d], 1);
    col_u[rowptr[d] + p] = (int32_t)s;
    int q = atomicAdd(&cur_out[s], 1);
    t_col_u[t_rowptr[s] + q] = (int32_t)d;
  }
}

constexpr int kShortRow = 16;

// rank sort, one thread per short row (deg <= kShortRow); long rows are listed for k_sort_long (a warp per row of a grid
// over ALL rows cost 31 us per CSR on the 389 k-node bench batch, which has no long row at all)
struct SortJob { const int32_t* rowptr; const int32_t* in; int32_t* out; int32_t* long_rows; int32_t* n_long; };
// (blockIdx.y = 0 / 1: the CSR and the transposed CSR are sorted by the same launches)
__global__ void k_sort_short(const SortJob j0, const SortJob j1, int64_t N) {
  const SortJob& jb = blockIdx.y ? j1 : j0;
  const int32_t* __restrict__ rowptr = jb.rowptr;
  const int32_t* __restrict__ in = jb.in;
  int32_t* __restrict__ out = jb.out;
  int32_t* __restrict__ long_rows = jb.long_rows;
  int32_t* __restrict__ n_long = jb.n_long;
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= N) return;
  int b = rowptr[r], e = rowptr[r + 1];
  int d = e - b;
  if (d > kShortRow) { long_rows[atomicAdd(n_long, 1)] = (int32_t)r; return; }
  int v[kShortRow];
#pragma unroll
  for (int i = 0; i < kShortRow; ++i) v[i] = (i < d) ? in[b + i] : 0x7fffffff;
#pragma unroll
  for (int i = 0; i < kShortRow; ++i) {
    if (i < d) {
      int rank = 0;
#pragma unroll
      for (int j = 0; j < kShortRow; ++j) {
        if (j < d) rank += (v[j] < v[i]) || (v[j] == v[i] && j < i);
      }
      out[b + rank] = v[i];
    }
  }
}

// rank sort, one warp per listed long row (the list order is arbitrary; every row is sorted on its own: deterministic)
__global__ void k_sort_long(const SortJob j0, const SortJob j1) {
  const SortJob& jb = blockIdx.y ? j1 : j0;
  const int32_t* __restrict__ rowptr = jb.rowptr;
  const int32_t* __restrict__ in = jb.in;
  int32_t* __restrict__ out = jb.out;
  const int32_t* __restrict__ long_rows = jb.long_rows;
  const int32_t* __restrict__ n_long = jb.n_long;
  constexpr int kStage = 512;                        // ids of a row staged in shared memory per warp (longer rows: from global)
  __shared__ int32_t s_row[8][kStage];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int n = *n_long;
  const int warps = (int)((gridDim.x * (int64_t)blockDim.x) >> 5);
  for (int idx = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5); idx < n; idx += warps) {
    const int r = long_rows[idx];
    const int b = rowptr[r], d = rowptr[r + 1] - b;
    const bool staged = d <= kStage;
    __syncwarp();
    if (staged) for (int i = lane; i < d; i += 32) s_row[wib][i] = in[b + i];
    __syncwarp();
    for (int i = lane; i < d; i += 32) {
      const int vi = staged ? s_row[wib][i] : in[b + i];
      int rank = 0;
      if (staged) {
        for (int j = 0; j < d; ++j) { const int vj = s_row[wib][j]; rank += (vj < vi) || (vj == vi && j < i); }   // broadcast reads
      } else {
        for (int j = 0; j < d; ++j) { const int vj = in[b + j]; rank += (vj < vi) || (vj == vi && j < i); }
      }
      out[b + rank] = vi;
    }
  }
}

// exclusive scans of both degree arrays (in -> rowptr, out -> t_rowptr) in three launches
int scan_counts2(const int32_t* cnt0, const int32_t* cnt1, int64_t N, int32_t* rowptr0, int32_t* rowptr1, int32_t* sums0,
                 int32_t* sums1, cudaStream_t s) {
  int64_t nb = ceil_div64(N, kScanChunk);
  const dim3 grid((unsigned)nb, 2);
  k_chunk_sums<<<grid, kScanThreads, 0, s>>>(cnt0, cnt1, N, sums0, sums1);
  GN_LAUNCHED();
  k_scan_sums<<<dim3(1, 2), kScanThreads, 0, s>>>(sums0, sums1, nb);
  GN_LAUNCHED();
  k_scan_final<<<grid, kScanThreads, 0, s>>>(cnt0, cnt1, N, sums0, sums1, rowptr0, rowptr1);
  GN_LAUNCHED();
  return GNODE_OK;
}

struct CsrWs {
  int32_t *cnt_in, *cnt_out, *sums, *sums_out, *col_u, *t_col_u, *flag, *n_long;
};

CsrWs carve(Arena& a, int64_t N, int64_t E) {
  CsrWs w;
  w.cnt_in = a.take<int32_t>(N + 1);
  w.cnt_out = a.take<int32_t>(N + 1);
  w.sums = a.take<int32_t>(ceil_div64(N > 0 ? N : 1, kScanChunk) + 1);
  w.sums_out = a.take<int32_t>(ceil_div64(N > 0 ? N : 1, kScanChunk) + 1);
  w.col_u = a.take<int32_t>(E > 0 ? E : 1);
  w.t_col_u = a.take<int32_t>(E > 0 ? E : 1);
  w.flag = a.take<int32_t>(1);
  w.n_long = a.take<int32_t>(2);
  return w;
}

}  // namespace
}  // namespace gnode

using namespace gnode;

extern "C" size_t gnode_csr_workspace_bytes(int64_t n_nodes, int64_t n_edges) {
  Arena a(nullptr, 0);
  carve(a, n_nodes, n_edges);
  return a.off;
}

extern "C" int gnode_csr_build_async(const int64_t* edge_index, int64_t E, int64_t N, int32_t* rowptr,
                                     int32_t* col, int32_t* t_rowptr, int32_t* t_col, int32_t* error_flag,
                                     void* workspace, size_t workspace_bytes, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_ARG(error_flag != nullptr, "gnode_csr_build_async: error_flag is null");
  GN_ARG(N > 0 && E >= 0, "gnode_csr_build: n_nodes must be > 0 and n_edges >= 0 (got %lld, %lld)",
         (long long)N, (long long)E);
  GN_ARG(N < (1ll << 31) - 1 && E < (1ll << 31) - 1, "gnode_csr_build: int32 CSR overflow");
  GN_ARG(rowptr && t_rowptr && (E == 0 || (col && t_col && edge_index)), "gnode_csr_build: null pointer");
  Arena a(workspace, workspace_bytes);
  CsrWs w = carve(a, N, E);
  GN_ARENA_OK(a, "gnode_csr_build");

  GN_PROF(s, 0.0, 16.0 * E + 8.0 * E + 8.0 * N, "csr_build");
  // cnt_in and cnt_out are consecutive arena blocks: one fill covers both (and the padding between them)
  const size_t cnt_span = (size_t)(reinterpret_cast<char*>(w.cnt_out + (N + 1)) - reinterpret_cast<char*>(w.cnt_in));
  GN_CUDA(cudaMemsetAsync(w.cnt_in, 0, cnt_span, s));
  GN_CUDA(cudaMemsetAsync(error_flag, 0, sizeof(int32_t), s));
  const int threads = 256;
  unsigned eblocks = (unsigned)(E > 0 ? (ceil_div64(E, threads) < 148 * 16 ? ceil_div64(E, threads) : 148 * 16) : 1);
  if (E > 0) {
    k_count<<<eblocks, threads, 0, s>>>(edge_index, E, N, w.cnt_in, w.cnt_out, error_flag);
    GN_LAUNCHED();
  }
  GN_TRY(scan_counts2(w.cnt_in, w.cnt_out, N, rowptr, t_rowptr, w.sums, w.sums_out, s));
  if (E > 0) {
    GN_CUDA(cudaMemsetAsync(w.cnt_in, 0, cnt_span, s));
    k_fill<<<eblocks, threads, 0, s>>>(edge_index, E, N, rowptr, t_rowptr, w.cnt_in, w.cnt_out,
                                       w.col_u, w.t_col_u);
    GN_LAUNCHED();
    unsigned rb = (unsigned)ceil_div64(N, threads);
    // long rows: listed by the short-row pass (the fill cursors cnt_in / cnt_out are free again and hold the lists), sorted by a
    // fixed grid of warps that walks the list
    int64_t wb64 = ceil_div64(N * 32, threads);
    unsigned wb = (unsigned)(wb64 < kNumSMs * 8 ? wb64 : kNumSMs * 8);
    GN_CUDA(cudaMemsetAsync(w.n_long, 0, 2 * sizeof(int32_t), s));
    const SortJob j0{rowptr, w.col_u, col, w.cnt_in, w.n_long}, j1{t_rowptr, w.t_col_u, t_col, w.cnt_out, w.n_long + 1};
    k_sort_short<<<dim3(rb, 2), threads, 0, s>>>(j0, j1, N);
    GN_LAUNCHED();
    k_sort_long<<<dim3(wb, 2), threads, 0, s>>>(j0, j1);
    GN_LAUNCHED();
  }
  return GNODE_OK;
}

extern "C" int gnode_csr_build(const int64_t* edge_index, int64_t E, int64_t N, int32_t* rowptr,
                               int32_t* col, int32_t* t_rowptr, int32_t* t_col, void* workspace,
                               size_t workspace_bytes, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_ARG(N > 0 && E >= 0, "gnode_csr_build: n_nodes must be > 0 and n_edges >= 0 (got %lld, %lld)",
         (long long)N, (long long)E);
  Arena a(workspace, workspace_bytes);
  CsrWs w = carve(a, N, E);
  GN_ARENA_OK(a, "gnode_csr_build");
  GN_TRY(gnode_csr_build_async(edge_index, E, N, rowptr, col, t_rowptr, t_col, w.flag, workspace, workspace_bytes, stream));
  int32_t flag = 0;
  GN_CUDA(cudaMemcpyAsync(&flag, w.flag, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  GN_CUDA(cudaStreamSynchronize(s));
  if (flag) {
    set_error("gnode_csr_build: edge_index holds node ids outside [0, %lld)", (long long)N);
    return GNODE_ERR_INDEX;
  }
  return GNODE_OK;
}
