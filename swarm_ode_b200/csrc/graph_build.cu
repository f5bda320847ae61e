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
