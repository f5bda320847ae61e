// tcgen05 (UMMA) 3xTF32 dense contraction -- placeholder until the tensor-core engine lands:
// reports "unsupported" for every shape so AUTO falls back to the FFMA anchor.
#include "common.cuh"

namespace gnode {
bool gemm_nt_tc_supported(const GemmNT&) { return false; }
int gemm_nt_tc(const GemmNT&, cudaStream_t) {
  set_error("gemm_nt_tc: tcgen05 engine not built");
  return GNODE_ERR_ARG;
}
}  // namespace gnode
