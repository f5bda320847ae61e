// Spatial edges of GraphConverter._compute_spatial_edges (scripts/train_gde.py:228-244), batched over
// independent snapshots and bit-exact: for i < j in lexicographic order, an edge pair (i,j),(j,i) is
// emitted iff sqrt((dy*dy) + (dx*dx)) < threshold, evaluated in float32 with the same roundings as
// numpy (no FMA contraction, IEEE sqrt).  One warp per snapshot; emission order is preserved with a
// ballot + popcount prefix inside the warp.
#include "common.cuh"

namespace gnode {
namespace {

__global__ void __launch_bounds__(128) k_spatial_edges(const float* __restrict__ pos, int64_t n_snap, int n,
                                                       float thr, int32_t* __restrict__ counts,
                                                       int32_t* __restrict__ edges) {
  const int lane = threadIdx.x & 31;
  const int64_t snap = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (snap >= n_snap) return;
  const float* p = pos + snap * (int64_t)n * 2;
  int32_t* out = edges + snap * (int64_t)n * (n - 1) * 2;
  int total = 0;  // directed edges so far (warp-uniform)
  for (int i = 0; i < n - 1; ++i) {
    const float yi = __ldg(p + 2 * i), xi = __ldg(p + 2 * i + 1);
    for (int j0 = i + 1; j0 < n; j0 += 32) {
      const int j = j0 + lane;
      bool hit = false;
      if (j < n) {
        const float dy = __fsub_rn(yi, __ldg(p + 2 * j));
        const float dx = __fsub_rn(xi, __ldg(p + 2 * j + 1));
        const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dx, dx)));
        hit = d < thr;
      }
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (hit) {
        const int before = __popc(m & ((1u << lane) - 1u));
        int32_t* o = out + (int64_t)(total + 2 * before) * 2;
        o[0] = i; o[1] = j;   // (src, dst) = (i, j)
        o[2] = j; o[3] = i;   // then (j, i)
      }
      total += 2 * __popc(m);
    }
  }
  if (lane == 0) counts[snap] = total;
}

}  // namespace
}  // namespace gnode

using namespace gnode;

extern "C" int gnode_spatial_edges(const float* pos, int64_t n_snap, int32_t n_agents, float threshold,
                                   int32_t* counts, int32_t* edges, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_ARG(n_snap >= 0 && n_agents >= 1, "gnode_spatial_edges: bad sizes");
  GN_ARG(n_snap == 0 || (pos && counts && (n_agents == 1 || edges)), "gnode_spatial_edges: null pointer");
  if (n_snap == 0) return GNODE_OK;
  const unsigned blocks = (unsigned)ceil_div64(n_snap * 32, 128);
  k_spatial_edges<<<blocks, 128, 0, s>>>(pos, n_snap, n_agents, threshold, counts, edges);
  GN_LAUNCHED();
  return GNODE_OK;
}

// ------------------------------------------------------------------------------------------------
// Whole-episode graph construction on the device: what WarehouseDataset does with one GraphConverter per
// episode (scripts/train_gde.py:308-314 calling :116-184 step by step) and Batch.from_data_list (:367) afterwards, for
// all T steps of an episode at once.  Step t sees the window of snapshots t-k .. t, k = min(t, W-1):
//   x          = the window's observation rows, oldest first                                  [(k+1) n, D]
//   edge_index = spatial(t-k) + 0 n, ..., spatial(t) + k n, then the temporal edges (k-1) n + a -> k n + a (k > 0),
//                each snapshot's spatial edges in the reference's emission order (gnode_spatial_edges)
//   is_current_agent = rows [k n, (k+1) n)
// and the T graphs are laid out as their disjoint union (node ids offset by ptr[t]).  Integer / copy work only.
// ------------------------------------------------------------------------------------------------
namespace gnode {
namespace {

__global__ void k_extract_pos(const float* __restrict__ obs, int64_t T, int n, int D, int num_agvs, float* __restrict__ pos) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= T * n) return;
  const int a = (int)(i % n);
  const float* row = obs + i * D;
  const int c = a < num_agvs ? 3 : 0;          // (y, x) = columns (3, 4) of AGV rows, (0, 1) of picker rows
  pos[2 * i] = row[c];
  pos[2 * i + 1] = row[c + 1];
}

// eoff[t] = first edge of graph t in the union (eoff[T] = total); one thread: T is an episode length
__global__ void k_window_edge_offsets(const int32_t* __restrict__ counts, int64_t T, int n, int W, int64_t* __restrict__ eoff) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int64_t run = 0, acc = 0;                     // run = sum of counts over the current window
  for (int64_t t = 0; t < T; ++t) {
    run += counts[t];
    if (t >= W) run -= counts[t - W];
    eoff[t] = acc;
    acc += run + ((t > 0 && W > 1) ? n : 0);      // temporal edges exist when the window holds a previous snapshot
  }
  eoff[T] = acc;
}

__device__ __forceinline__ int64_t window_ptr(int64_t t, int n, int W) {
  // sum_{t' < t} (min(t', W-1) + 1) n
  const int64_t full = t > W - 1 ? t - (W - 1) : 0;            // steps with a full window
  const int64_t ramp = t < W - 1 ? t : W - 1;                  // steps 0 .. ramp-1 have t'+1 snapshots
  return (ramp * (ramp + 1) / 2 + full * W) * n;
}

// one block per (step t, window slot i): node rows and spatial edges of snapshot t-k+i; slot k also writes the temporal edges
__global__ void __launch_bounds__(256) k_window_assemble(const float* __restrict__ obs, const int32_t* __restrict__ counts,
                                                         const int32_t* __restrict__ edges, const int64_t* __restrict__ eoff,
                                                         int64_t T, int n, int D, int W, int64_t e_cap,
                                                         float* __restrict__ x, int64_t* __restrict__ ei_src,
                                                         int64_t* __restrict__ ei_dst, int64_t* __restrict__ batch,
                                                         bool* __restrict__ cur, int64_t* __restrict__ ptr) {
  const int64_t t = blockIdx.x / W;
  const int i = (int)(blockIdx.x % W);
  const int k = (int)(t < W - 1 ? t : W - 1);
  if (i > k) return;
  const int64_t snap = t - k + i;
  const int64_t p0 = window_ptr(t, n, W);
  if (i == 0 && threadIdx.x == 0) { ptr[t] = p0; if (t == T - 1) ptr[T] = window_ptr(T, n, W); }
  // rows
  const int64_t row0 = p0 + (int64_t)i * n;
  for (int64_t idx = threadIdx.x; idx < (int64_t)n * D; idx += blockDim.x) x[row0 * D + idx] = obs[snap * n * D + idx];
  for (int a = threadIdx.x; a < n; a += blockDim.x) { batch[row0 + a] = t; cur[row0 + a] = (i == k); }
  // spatial edges of this snapshot, after those of the older slots
  int64_t e0 = eoff[t];
  for (int j = 0; j < i; ++j) e0 += counts[t - k + j];
  const int cnt = counts[snap];
  const int32_t* src = edges + snap * (int64_t)n * (n - 1) * 2;
  for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
    ei_src[e0 + e] = p0 + (int64_t)i * n + src[2 * e];
    ei_dst[e0 + e] = p0 + (int64_t)i * n + src[2 * e + 1];
  }
  if (i == k && k > 0) {
    const int64_t te = e0 + cnt;
    for (int a = threadIdx.x; a < n; a += blockDim.x) {
      ei_src[te + a] = p0 + (int64_t)(k - 1) * n + a;
      ei_dst[te + a] = p0 + (int64_t)k * n + a;
    }
  }
  (void)e_cap;
}

}  // namespace
}  // namespace gnode

extern "C" int64_t gnode_window_graphs_nodes(int64_t n_steps, int32_t n_agents, int32_t window) {
  const int64_t W = window, t = n_steps;
  const int64_t full = t > W - 1 ? t - (W - 1) : 0, ramp = t < W - 1 ? t : W - 1;
  return (ramp * (ramp + 1) / 2 + full * W) * n_agents;
}

extern "C" int64_t gnode_window_graphs_edge_capacity(int64_t n_steps, int32_t n_agents, int32_t window) {
  return gnode_window_graphs_nodes(n_steps, n_agents, window) * (int64_t)(n_agents - 1) + n_steps * (int64_t)n_agents;
}

extern "C" size_t gnode_window_graphs_workspace_bytes(int64_t n_steps, int32_t n_agents) {
  Arena a(nullptr, 0);
  a.take<float>((size_t)n_steps * n_agents * 2);
  a.take<int32_t>((size_t)n_steps);
  a.take<int32_t>((size_t)n_steps * (size_t)(n_agents * (n_agents - 1) > 1 ? n_agents * (n_agents - 1) : 1) * 2);
  return a.off;
}

extern "C" int gnode_window_graphs(const float* obs, int64_t n_steps, int32_t n_agents, int32_t node_dim, int32_t num_agvs,
                                   float threshold, int32_t window, float* x, int64_t* edge_index, int64_t edge_capacity,
                                   int64_t* batch, uint8_t* is_current_agent, int64_t* ptr, int64_t* edge_offsets,
                                   void* workspace, size_t workspace_bytes, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_ARG(obs && x && edge_index && batch && is_current_agent && ptr && edge_offsets, "gnode_window_graphs: null pointer");
  GN_ARG(n_steps >= 1 && n_agents >= 1 && node_dim >= 5 && window >= 1 && num_agvs >= 0 && num_agvs <= n_agents,
         "gnode_window_graphs: bad sizes (node_dim must hold the position columns 0..4)");
  GN_ARG(edge_capacity >= gnode_window_graphs_edge_capacity(n_steps, n_agents, window),
         "gnode_window_graphs: edge_index capacity too small (need %lld columns)",
         (long long)gnode_window_graphs_edge_capacity(n_steps, n_agents, window));
  Arena a(workspace, workspace_bytes);
  float* pos = a.take<float>((size_t)n_steps * n_agents * 2);
  int32_t* counts = a.take<int32_t>((size_t)n_steps);
  int32_t* edges = a.take<int32_t>((size_t)n_steps * (size_t)(n_agents * (n_agents - 1) > 1 ? n_agents * (n_agents - 1) : 1) * 2);
  GN_ARENA_OK(a, "gnode_window_graphs");
  k_extract_pos<<<(unsigned)ceil_div64(n_steps * n_agents, 256), 256, 0, s>>>(obs, n_steps, n_agents, node_dim, num_agvs, pos);
  GN_LAUNCHED();
  GN_TRY(gnode_spatial_edges(pos, n_steps, n_agents, threshold, counts, edges, stream));
  k_window_edge_offsets<<<1, 32, 0, s>>>(counts, n_steps, n_agents, window, edge_offsets);
  GN_LAUNCHED();
  k_window_assemble<<<(unsigned)(n_steps * window), 256, 0, s>>>(obs, counts, edges, edge_offsets, n_steps, n_agents, node_dim,
                                                                   window, edge_capacity, x, edge_index,
                                                                   edge_index + edge_capacity, batch,
                                                                   reinterpret_cast<bool*>(is_current_agent), ptr);
  GN_LAUNCHED();
  return GNODE_OK;
}
