// Declarations of the elementwise RK helpers (rk.cu).
#pragma once
#include "common.cuh"

namespace gnode {

int norm_blocks(int64_t n);  // number of double partials the two norm kernels need

// *out = sum_i ((a_i - b_i) / (atol + rtol*|y_i|))^2   (b may be null); deterministic
int scaled_sumsq(const float* a, const float* b, const float* y, float atol, float rtol, int64_t n,
                 double* partials, double* out, cudaStream_t s);
// *out = sum_i (err_i / (atol + rtol*max(|y0_i|,|y1_i|)))^2 with err = sum_j lc.in[j]*lc.coef[j]
int error_sumsq(const LinComb& lc, const float* y0, const float* y1, float atol, float rtol,
                double* partials, double* out, cudaStream_t s);
// dense output at x = (t - t0)/(t1 - t0); lc.in = k_0..k_6, lc.coef = dt*c_mid
int dopri_interp(const LinComb& lc, const float* y0, const float* y1, float dt, float x, float* out,
                 cudaStream_t s);

struct PackSegHost {
  float* dst; const float* src;
  int rows, cols;
  int64_t ld_src, ld_dst;
  int transpose, accumulate;
};
int pack_segments(const PackSegHost* segs, int n, cudaStream_t s);

}  // namespace gnode
