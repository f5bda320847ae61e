// Engine dispatch for the dense contractions: tcgen05 (3xTF32, gemm_tc.cu) or fp32 FFMA (gemm_simt.cu).
#include "common.cuh"

namespace gnode {
int gemm_nt_simt(const GemmNT& g, cudaStream_t s);
bool gemm_nt_tc_supported(const GemmNT& g);
int gemm_nt_tc(const GemmNT& g, cudaStream_t s);

int gemm_nt(const GemmNT& g, cudaStream_t s) {
  const int engine = current_engine();
  const bool tc = engine != GNODE_ENGINE_SIMT && gemm_nt_tc_supported(g);
  GN_PROF(s, 2.0 * g.M * g.N * g.K, 4.0 * ((double)g.M * g.K + (double)g.N * g.K + (double)g.M * g.N * (g.base ? 2 : 1)),
          "gemm_nt[%s] N=%d K=%d", tc ? "tcgen05" : "ffma", g.N, g.K);
  if (engine == GNODE_ENGINE_SIMT) return gemm_nt_simt(g, s);
  if (gemm_nt_tc_supported(g)) return gemm_nt_tc(g, s);
  if (engine == GNODE_ENGINE_TC) {
    set_error("gemm_nt: shape M=%lld N=%d K=%d not supported by the tcgen05 engine", (long long)g.M, g.N, g.K);
    return GNODE_ERR_ARG;
  }
  return gemm_nt_simt(g, s);
}
}  // namespace gnode
