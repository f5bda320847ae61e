// Optional per-kernel-class timing with CUDA events on the launching stream (used by bench.py for
// the roofline numbers).  Disabled by default: a ProfScope then costs one relaxed atomic load.
#pragma once
#include <cuda_runtime.h>

namespace gnode {

bool prof_enabled();

struct ProfScope {
  // label: kernel class (e.g. "gemm_nt N=128 K=399"); flops / bytes: ALGORITHMIC work of this launch
  ProfScope(const char* label, cudaStream_t s, double flops, double bytes);
  ~ProfScope();
  int slot = -1;
  cudaStream_t stream;
};

}  // namespace gnode
