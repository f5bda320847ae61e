// ODE integration drivers: the solver loops of torchdiffeq.odeint restated as native host code that
// enqueues the CUDA stages (reference call sites scripts/train_gde.py:78-85, scripts/gnode.py:136-137,
// scripts/run_gnode.py:134-135), plus backprop through the fixed-grid solvers
// (loss.backward(), scripts/train_gde.py:493).
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "field.cuh"

namespace gnode {

// ------------------------------------------------------------------------------------------------
// Tableaus
// ------------------------------------------------------------------------------------------------
namespace {

Tableau make_euler() {
  Tableau t{};
  t.S = 1;
  t.c_sol[0] = 1.0;
  return t;
}
Tableau make_midpoint() {
  Tableau t{};
  t.S = 2;
  t.beta[1][0] = 0.5;
  t.c_sol[1] = 1.0;
  return t;
}
// torchdiffeq 'rk4' is rk4_alt_step_func: the 3/8 rule
Tableau make_rk4_38() {
  Tableau t{};
  t.S = 4;
  t.beta[1][0] = 1.0 / 3.0;
  t.beta[2][0] = -1.0 / 3.0; t.beta[2][1] = 1.0;
  t.beta[3][0] = 1.0; t.beta[3][1] = -1.0; t.beta[3][2] = 1.0;
  t.c_sol[0] = 0.125; t.c_sol[1] = 0.375; t.c_sol[2] = 0.375; t.c_sol[3] = 0.125;
  return t;
}
Tableau make_dopri5() {
  Tableau t{};
  t.S = 7;
  const double b[7][7] = {
      {0},
      {1.0 / 5},
      {3.0 / 40, 9.0 / 40},
      {44.0 / 45, -56.0 / 15, 32.0 / 9},
      {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729},
      {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656},
      {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84},
  };
  std::memcpy(t.beta, b, sizeof(b));
  const double cs[7] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84, 0};
  const double ce[7] = {35.0 / 384 - 1951.0 / 21600,
                        0,
                        500.0 / 1113 - 22642.0 / 50085,
                        125.0 / 192 - 451.0 / 720,
                        -2187.0 / 6784 - -12231.0 / 42400,
                        11.0 / 84 - 649.0 / 6300,
                        -1.0 / 60.0};
  const double cm[7] = {6025192743.0 / 30085553152.0 / 2,
                        0,
                        51252292925.0 / 65400821598.0 / 2,
                        -2691868925.0 / 45128329728.0 / 2,
                        187940372067.0 / 1594534317056.0 / 2,
                        -1776094331.0 / 19743644256.0 / 2,
                        11237099.0 / 235043384.0 / 2};
  std::memcpy(t.c_sol, cs, sizeof(cs));
  std::memcpy(t.c_err, ce, sizeof(ce));
  std::memcpy(t.c_mid, cm, sizeof(cm));
  return t;
}

}  // namespace

const Tableau* tableau_for(int method) {
  static const Tableau euler = make_euler(), mid = make_midpoint(), rk4 = make_rk4_38(), dp = make_dopri5();
  switch (method) {
    case GNODE_EULER: return &euler;
    case GNODE_MIDPOINT: return &mid;
    case GNODE_RK4_38: return &rk4;
    case GNODE_DOPRI5: return &dp;
    default: return nullptr;
  }
}

// x_s = y + dt * sum_{j<s} beta[s][j] * k_j     (coefficients formed in fp32 like the state dtype)
static int stage_input(const Tableau& tb, int s_idx, const float* y, float* const* k, float dt, float* xs,
                       int64_t n, cudaStream_t s) {
  LinComb lc{};
  lc.out = xs; lc.base = y; lc.n = n; lc.n_terms = 0;
  for (int j = 0; j < s_idx; ++j) {
    lc.in[lc.n_terms] = k[j];
    lc.coef[lc.n_terms] = (float)tb.beta[s_idx][j] * dt;
    ++lc.n_terms;
  }
  return lincomb(lc, s);
}

// ------------------------------------------------------------------------------------------------
// Fixed grid: grid == t, solution[j+1] = y_j + step(y_j)
// ------------------------------------------------------------------------------------------------
int integrate_fixed(Field& f, int method, const float* y0, const float* t, int n_t, float* sol,
                    float* const* kbuf, float* xs_shared, cudaStream_t s, StepSaver* saver, bool sol0_by_caller) {
  const Tableau* tbp = tableau_for(method);
  if (!tbp || method == GNODE_DOPRI5) { set_error("integrate_fixed: bad method %d", method); return GNODE_ERR_ARG; }
  const Tableau& tb = *tbp;
  const int64_t n = f.numel();
  if (sol != y0 && !sol0_by_caller) GN_CUDA(cudaMemcpyAsync(sol, y0, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
  for (int j = 0; j + 1 < n_t; ++j) {
    const float dt = t[j + 1] - t[j];
    const float* y = (j == 0 && sol0_by_caller) ? y0 : sol + (int64_t)j * n;
    float* y1 = sol + (int64_t)(j + 1) * n;
    const int S = tb.S;
    // with a saver every stage keeps its input and its layer intermediates (slot = stage) for the backward pass
    float* const* xs_step = saver ? saver->begin_step(j) : nullptr;
    for (int st = 0; st < S - 1; ++st) {
      const float* x = y;
      float* xs = saver ? xs_step[st] : xs_shared;
      if (st > 0) { GN_TRY(stage_input(tb, st, y, kbuf, dt, xs, n, s)); x = xs; }
      GN_TRY(f.eval(x, kbuf[st], nullptr, 1.f, saver ? st : 0, s));
    }
    // last stage: y1 = (y + dt * sum_{j<S-1} c_j k_j) + dt * c_{S-1} * f(x_{S-1}) fused into the field's epilogue
    const float* x = y;
    float* xs = saver ? xs_step[S - 1] : xs_shared;
    if (S > 1) { GN_TRY(stage_input(tb, S - 1, y, kbuf, dt, xs, n, s)); x = xs; }
    const float* base = y;
    if (S > 1) {
      LinComb lc{};
      lc.out = y1; lc.base = y; lc.n = n; lc.n_terms = 0;
      for (int q = 0; q < S - 1; ++q) { lc.in[lc.n_terms] = kbuf[q]; lc.coef[lc.n_terms] = (float)tb.c_sol[q] * dt; ++lc.n_terms; }
      GN_TRY(lincomb(lc, s));
      base = y1;
    }
    GN_TRY(f.eval(x, y1, base, (float)tb.c_sol[S - 1] * dt, saver ? S - 1 : 0, s));
  }
  return GNODE_OK;
}

// ------------------------------------------------------------------------------------------------
// Adaptive dopri5 (torchdiffeq RKAdaptiveStepsizeODESolver semantics)
// ------------------------------------------------------------------------------------------------
namespace {

// device-side exchange of the norm registered for the calling thread (gnode_set_dopri5_device_allreduce)
thread_local gnode_allreduce_dev_fn t_allreduce_dev = nullptr;
thread_local void* t_allreduce_dev_user = nullptr;

struct NormCtx {
  gnode_allreduce_fn allreduce; void* user; double* dsum; cudaStream_t s; int64_t n;
  gnode_allreduce_dev_fn allreduce_dev = nullptr; void* dev_user = nullptr;
  double n_global = 0.0;       // element count over all ranks (device-side exchange: fetched once through the host hook)
  // rms over all ranks of a local sum of squares already sitting in *dsum (device)
  int finish(float* out) {
    double h[2];
    if (allreduce_dev) {
      // the sum is exchanged where it lies: the hook enqueues an in-place SUM all-reduce of *dsum on the solve's stream
      // (NCCL); the host reads the global sum with the one synchronisation the step-size decision needs anyway
      if (n_global == 0.0) {
        h[0] = 0.0; h[1] = (double)n;
        if (allreduce) allreduce(h, user);
        n_global = h[1];
      }
      if (allreduce_dev(dsum, dev_user) != 0) { set_error("dopri5: the device-side norm exchange failed"); return GNODE_ERR_SOLVER; }
    }
    if (cudaMemcpyAsync(&h[0], dsum, sizeof(double), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) {
      set_error("dopri5: device error while reading the error norm: %s", cudaGetErrorString(cudaGetLastError()));
      return GNODE_ERR_CUDA;
    }
    if (allreduce_dev) { *out = (float)std::sqrt(h[0] / n_global); return GNODE_OK; }
    h[1] = (double)n;
    if (allreduce) allreduce(h, user);
    *out = (float)std::sqrt(h[0] / h[1]);
    return GNODE_OK;
  }
};

}  // namespace

int integrate_dopri5(Field& f, const float* y0, const double* t, int n_t, double rtol, double atol, float* sol,
                     gnode_dopri5_stats* stats, const gnode_dopri5_trace* trace, gnode_allreduce_fn allreduce,
                     void* allreduce_user, int64_t max_num_steps, const Dopri5Bufs& b, cudaStream_t s) {
  const Tableau& tb = *tableau_for(GNODE_DOPRI5);
  const int64_t n = f.numel();
  const float rtolf = (float)rtol, atolf = (float)atol;
  gnode_dopri5_stats st{};
  st.min_margin = INFINITY;
  NormCtx nc{allreduce, allreduce_user, b.dsum, s, n, t_allreduce_dev, t_allreduce_dev_user};

  float* k[7];
  for (int i = 0; i < 7; ++i) k[i] = b.k[i];
  float* ya = b.ya;
  float* yb = b.yb;

  if (sol != y0) GN_CUDA(cudaMemcpyAsync(sol, y0, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
  GN_CUDA(cudaMemcpyAsync(ya, y0, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
  // f0
  GN_TRY(f.eval(ya, k[0], nullptr, 1.f, 0, s));
  st.nfe++;

  // ---- _select_initial_step (order = 4), fp32 scalar arithmetic like the state dtype ----
  double dt;
  {
    float d0, d1, d2;
    GN_TRY(scaled_sumsq(ya, nullptr, ya, atolf, rtolf, n, b.partials, b.dsum, s));
    GN_TRY(nc.finish(&d0));
    GN_TRY(scaled_sumsq(k[0], nullptr, ya, atolf, rtolf, n, b.partials, b.dsum, s));
    GN_TRY(nc.finish(&d1));
    float h0;
    if (d0 < 1e-5f || d1 < 1e-5f) h0 = 1e-6f; else h0 = 0.01f * d0 / d1;
    h0 = fabsf(h0);
    LinComb lc{};
    lc.out = b.xs; lc.base = ya; lc.in[0] = k[0]; lc.coef[0] = h0; lc.n_terms = 1; lc.n = n;
    GN_TRY(lincomb(lc, s));
    GN_TRY(f.eval(b.xs, k[1], nullptr, 1.f, 0, s));
    st.nfe++;
    GN_TRY(scaled_sumsq(k[1], k[0], ya, atolf, rtolf, n, b.partials, b.dsum, s));
    GN_TRY(nc.finish(&d2));
    d2 = fabsf(d2 / h0);
    float h1;
    if (d1 <= 1e-15f && d2 <= 1e-15f) h1 = fmaxf(1e-6f, h0 * 1e-3f);
    else h1 = powf(0.01f / fmaxf(d1, d2), 1.0f / 5.0f);
    h1 = fabsf(h1);
    dt = (double)fminf(100.f * h0, h1);
  }
  st.first_step = dt;

  double t_cur = t[0];          // rk_state.t1
  int next_out = 1;
  int64_t n_steps = 0;
  while (next_out < n_t) {
    if (!(t[next_out] > t_cur)) {
      // cannot happen for strictly increasing t after a step; kept for n_t points at/below t0
      GN_CUDA(cudaMemcpyAsync(sol + (int64_t)next_out * n, ya, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
      ++next_out;
      continue;
    }
    if (n_steps >= max_num_steps) { set_error("dopri5: max_num_steps exceeded (%lld)", (long long)n_steps); return GNODE_ERR_SOLVER; }
    const double t0 = t_cur, t1 = t0 + dt;
    if (!(t0 + dt > t0)) { set_error("dopri5: underflow in dt %g", dt); return GNODE_ERR_SOLVER; }
    const float dtf = (float)dt;
    // stages 1..6 ; stage 6 input is the 5th-order solution y1 (FSAL)
    for (int si = 1; si < 7; ++si) {
      float* xs = (si == 6) ? yb : b.xs;
      GN_TRY(stage_input(tb, si, ya, k, dtf, xs, n, s));
      GN_TRY(f.eval(xs, k[si], nullptr, 1.f, 0, s));
      st.nfe++;
    }
    // error ratio
    LinComb le{};
    le.n = n; le.n_terms = 0;
    for (int q = 0; q < 7; ++q) { le.in[le.n_terms] = k[q]; le.coef[le.n_terms] = dtf * (float)tb.c_err[q]; ++le.n_terms; }
    // zero coefficient (c_err[1] == 0) still multiplies in the reference; it contributes exactly +0
    GN_TRY(error_sumsq(le, ya, yb, atolf, rtolf, b.partials, b.dsum, s));
    float ratio;
    GN_TRY(nc.finish(&ratio));
    if (!std::isfinite(ratio)) { set_error("dopri5: non-finite values in state `y` (error ratio %g at t=%g, dt=%g)", (double)ratio, t0, dt); return GNODE_ERR_SOLVER; }
    const bool accept = ratio <= 1.0f;
    if (trace && trace->trace_cap > st.n_attempted) {
      if (trace->error_ratio) trace->error_ratio[st.n_attempted] = ratio;
      if (trace->dt) trace->dt[st.n_attempted] = dt;
      if (trace->accepted) trace->accepted[st.n_attempted] = accept ? 1 : 0;
    }
    st.n_attempted++;
    const double margin = std::fabs((double)ratio - 1.0);
    if (margin < st.min_margin) st.min_margin = margin;
    if (accept) {
      st.n_accepted++;
      // dense output for every requested time inside (t0, t1]
      while (next_out < n_t && t[next_out] <= t1) {
        const float x = (float)((t[next_out] - t0) / (t1 - t0));
        LinComb lm{};
        lm.n = n; lm.n_terms = 7;
        for (int q = 0; q < 7; ++q) { lm.in[q] = k[q]; lm.coef[q] = dtf * (float)tb.c_mid[q]; }
        GN_TRY(dopri_interp(lm, ya, yb, dtf, x, sol + (int64_t)next_out * n, s));
        ++next_out;
      }
      float* tmp = ya; ya = yb; yb = tmp;
      tmp = k[0]; k[0] = k[6]; k[6] = tmp;   // FSAL: f(t1, y1) is the next step's f0
      t_cur = t1;
    }
    // _optimal_step_size
    double factor;
    if (ratio == 0.f) {
      factor = 10.0;
    } else {
      const double dfactor = (ratio < 1.f) ? 1.0 : 0.2;
      const double er = (double)ratio;
      factor = std::fmin(10.0, std::fmax(0.9 / std::pow(er, 0.2), dfactor));
    }
    dt = dt * factor;
    ++n_steps;
  }
  st.last_dt = dt;
  if (stats) *stats = st;
  return GNODE_OK;
}


// ------------------------------------------------------------------------------------------------
// Adaptive dopri5 with folded stages (fold.cu): per attempted step the D-wide state is touched by three dense
// contractions (Z_0 = y @ w1cat^T on the way in; y_1 and the error estimate on the way out) instead of two per stage;
// the seven stages run 2H-wide (graph-resident when the batch carries whole-graph tiles).  The controller, the norms
// and the dense output keep torchdiffeq's arithmetic; only the stage evaluations are re-associated.
// ------------------------------------------------------------------------------------------------
struct Dopri5FoldBufs {
  float *ya, *yb, *k0, *k1, *xs, *err;   // [N, D] each
  float* Cerr;                           // [N, 2H]  dt sum_s c_err[s] cat2_s
  float* znext;                          // [N, 2H]  Z of stage 6 = Z_0 of the next step when this one is accepted
  double* partials; double* dsum;
};

// gnode_set_dopri5_fsal(0) / GNODE_DOPRI5_FSAL=0 switch the two step-level re-associations of the folded dopri5 off (A/B
// and bisecting):
//  * FSAL in the folded space: stage 6 of an attempt is evaluated at the step's solution, so its Z is y_1 @ w1cat^T; an
//    accepted step hands it to the next attempt as Z_0 and a rejected step keeps its Z_0 (same y, new dt): the D-wide
//    contraction Z_0 = y @ w1cat^T runs once per solve instead of once per attempt;
//  * dense output as ONE projection: torchdiffeq's quartic through (y0, y1, y_mid, f0, f1) is linear in the seven stage
//    derivatives and the y0 terms of its coefficients cancel, sol(x) = y0 + dt sum_s w_s(x) k_s (fold.cu:
//    dopri5_dense_weights, the weights the backward pass already uses), so an output inside a step is a 2H-wide stage
//    combination and one D-wide projection instead of three projections and a five-operand D-wide polynomial.
// Both change fp32 rounding order only (parity gates: rel-L2 <= 1e-4, identical accept / reject lists).
static bool dopri5_fsal_on() { return current_dopri5_fsal() != 0; }

int integrate_dopri5_folded(Sage3Ctx& c, FoldWs& f, const float* y0, const double* t, int n_t, double rtol, double atol,
                            float* sol, gnode_dopri5_stats* stats, const gnode_dopri5_trace* trace,
                            gnode_allreduce_fn allreduce, void* allreduce_user, int64_t max_num_steps,
                            const Dopri5FoldBufs& b, cudaStream_t s) {
  const Tableau& tb = *tableau_for(GNODE_DOPRI5);
  const int64_t n = c.numel();
  const int H2 = 2 * c.H;
  const int64_t nh = c.N * H2;
  const float rtolf = (float)rtol, atolf = (float)atol;
  gnode_dopri5_stats st{};
  st.min_margin = INFINITY;
  NormCtx nc{allreduce, allreduce_user, b.dsum, s, n, t_allreduce_dev, t_allreduce_dev_user};
  float* ya = b.ya;
  float* yb = b.yb;

  if (sol != y0) GN_CUDA(cudaMemcpyAsync(sol, y0, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
  GN_CUDA(cudaMemcpyAsync(ya, y0, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
  GN_TRY(f.prepare(c, s));
  f.bind_slots(c, nullptr, 0);
  f.forward_only = true;                // the dopri5 backward replays its accepted steps: nothing of this solve is read back
  const bool fsal = dopri5_fsal_on();
  float* const z0_own = f.z0;           // restored on every way out (the buffers are swapped on accepted steps)
  struct Restore { FoldWs& f; float* z0; ~Restore() { f.z0 = z0; f.z_next = nullptr; f.z0_ready = false; f.z_next_valid = false; } } restore{f, z0_own};
  f.z_next = fsal ? b.znext : nullptr;
  f.z0_ready = false;

  // k = scale_w * (W @ w3cat^T) + scale_b * b3   for a 2H-wide W; optional base
  auto project = [&](const float* W, float* out, const float* base, float bias_scale) -> int {
    GemmNT q{};
    q.A = W; q.lda = H2; q.B = c.w3cat; q.ldb = H2; q.C = out; q.ldc = c.D; q.M = c.N; q.N = c.D; q.K = H2;
    q.bias = c.b3; q.bias_scale = bias_scale; q.base = base; q.ldbase = c.D;
    c.use_w3(q);
    q.rows_engine = 1;      // dense D-wide rows in and out: the row-major K = 128 engine when D <= 440 (gemm_k128.cu)
    return gemm_nt(q, s);
  };
  auto combine = [&](const double* coef, float dtf, double* sum_out) -> int {   // Cbuf = dt * sum_j coef_j cat2_j
    LinComb lc{};
    lc.out = f.Cbuf; lc.base = nullptr; lc.n = nh; lc.n_terms = 0;
    double sum = 0.0;
    for (int j = 0; j < 7; ++j) {
      sum += coef[j];
      if (coef[j] != 0.0) { lc.in[lc.n_terms] = f.cat2[j]; lc.coef[lc.n_terms] = (float)coef[j] * dtf; ++lc.n_terms; }
    }
    *sum_out = sum;
    return lincomb(lc, s);
  };

  // ---- _select_initial_step: f(y0) and f(y0 + h0 f(y0)), D-wide because the norms need the derivatives.  Folded form:
  // stage 0 of a one-stage tableau, then stage 1 of the two-stage tableau beta_10 = 1 with dt = h0 (= the field at
  // y0 + h0 k_0), each followed by one projection k = cat2 @ w3cat^T + b3; Z_0 of y0 stays for the first attempt.
  // Without the re-associations (gnode_set_dopri5_fsal(0)): two direct evaluations of the field. ----
  const int S_own = f.S;
  struct RestoreS { FoldWs& f; int S; ~RestoreS() { f.S = S; } } restore_s{f, S_own};
  Tableau t_init{};
  t_init.S = 1;
  if (fsal) {
    f.S = 1;
    GN_TRY(f.forward_stages(c, t_init, ya, 0.f, s));
    f.z0_ready = true;
    GN_TRY(project(f.cat2[0], b.k0, nullptr, 1.f));
  } else {
    GN_TRY(c.eval(ya, b.k0, nullptr, 1.f, 0, s));
  }
  st.nfe++;
  double dt;
  {
    float d0, d1, d2;
    GN_TRY(scaled_sumsq(ya, nullptr, ya, atolf, rtolf, n, b.partials, b.dsum, s));
    GN_TRY(nc.finish(&d0));
    GN_TRY(scaled_sumsq(b.k0, nullptr, ya, atolf, rtolf, n, b.partials, b.dsum, s));
    GN_TRY(nc.finish(&d1));
    float h0;
    if (d0 < 1e-5f || d1 < 1e-5f) h0 = 1e-6f; else h0 = 0.01f * d0 / d1;
    h0 = fabsf(h0);
    if (fsal) {
      t_init.S = 2;
      t_init.beta[1][0] = 1.0;
      f.S = 2;
      GN_TRY(f.forward_stages(c, t_init, ya, h0, s));
      GN_TRY(project(f.cat2[1], b.k1, nullptr, 1.f));
    } else {
      LinComb lc{};
      lc.out = b.xs; lc.base = ya; lc.in[0] = b.k0; lc.coef[0] = h0; lc.n_terms = 1; lc.n = n;
      GN_TRY(lincomb(lc, s));
      GN_TRY(c.eval(b.xs, b.k1, nullptr, 1.f, 0, s));
    }
    f.S = S_own;
    st.nfe++;
    GN_TRY(scaled_sumsq(b.k1, b.k0, ya, atolf, rtolf, n, b.partials, b.dsum, s));
    GN_TRY(nc.finish(&d2));
    d2 = fabsf(d2 / h0);
    float h1;
    if (d1 <= 1e-15f && d2 <= 1e-15f) h1 = fmaxf(1e-6f, h0 * 1e-3f);
    else h1 = powf(0.01f / fmaxf(d1, d2), 1.0f / 5.0f);
    h1 = fabsf(h1);
    dt = (double)fminf(100.f * h0, h1);
  }
  st.first_step = dt;

  double t_cur = t[0];
  int next_out = 1;
  int64_t n_steps = 0;
  while (next_out < n_t) {
    if (!(t[next_out] > t_cur)) {
      GN_CUDA(cudaMemcpyAsync(sol + (int64_t)next_out * n, ya, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
      ++next_out;
      continue;
    }
    if (n_steps >= max_num_steps) { set_error("dopri5: max_num_steps exceeded (%lld)", (long long)n_steps); return GNODE_ERR_SOLVER; }
    const double t0 = t_cur, t1 = t0 + dt;
    if (!(t0 + dt > t0)) { set_error("dopri5: underflow in dt %g", dt); return GNODE_ERR_SOLVER; }
    const float dtf = (float)dt;
    // seven stages, 2H-wide (stage 6's input is the 5th-order solution: beta[6] == c_sol, FSAL)
    // both 2H-wide stage combinations (solution and error-estimate weights) come out of the stage kernel
    GN_TRY(f.forward_stages(c, tb, ya, dtf, s, f.Cbuf, b.Cerr, tb.c_err));
    st.nfe += 6;
    const bool z_next_valid = f.z_next_valid;
    if (fsal) f.z0_ready = true;        // Z_0 of ya is in f.z0 now: a rejected step reuses it as it is
    double csum = 0.0, cesum = 0.0;
    for (int j = 0; j < 7; ++j) { csum += tb.c_sol[j]; cesum += tb.c_err[j]; }
    GN_TRY(project(f.Cbuf, yb, ya, (float)csum * dtf));                 // y1 = y0 + dt sum c_sol k
    GN_TRY(project(b.Cerr, b.err, nullptr, (float)cesum * dtf));        // err = dt sum c_err k
    // (a norm accumulated in the epilogue of this projection -- no err write, no norm pass -- was built and measured in
    // round 2: 7.5 ms against 2.9 + 1.8 ms for the two kernels at N = 2.3 M, D = 435; the four epilogue warps of the
    // general engine cannot stream two D-wide operands, so the norm stays a separate copy-speed pass)
    LinComb le{};
    le.n = n; le.n_terms = 1; le.in[0] = b.err; le.coef[0] = 1.f;
    GN_TRY(error_sumsq(le, ya, yb, atolf, rtolf, b.partials, b.dsum, s));
    float ratio;
    GN_TRY(nc.finish(&ratio));
    if (!std::isfinite(ratio)) { set_error("dopri5: non-finite values in state `y` (error ratio %g at t=%g, dt=%g)", (double)ratio, t0, dt); return GNODE_ERR_SOLVER; }
    const bool accept = ratio <= 1.0f;
    if (trace && trace->trace_cap > st.n_attempted) {
      if (trace->error_ratio) trace->error_ratio[st.n_attempted] = ratio;
      if (trace->dt) trace->dt[st.n_attempted] = dt;
      if (trace->accepted) trace->accepted[st.n_attempted] = accept ? 1 : 0;
    }
    st.n_attempted++;
    const double margin = std::fabs((double)ratio - 1.0);
    if (margin < st.min_margin) st.min_margin = margin;
    if (accept) {
      st.n_accepted++;
      if (fsal) {
        while (next_out < n_t && t[next_out] <= t1) {
          float* out = sol + (int64_t)next_out * n;
          if (t[next_out] == t1) {          // w(1) = c_sol: the interpolant at the end of the step is y1
            GN_CUDA(cudaMemcpyAsync(out, yb, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
          } else {
            double w[7], wsum;
            dopri5_dense_weights(tb, (t[next_out] - t0) / (t1 - t0), w);
            GN_TRY(combine(w, dtf, &wsum));
            GN_TRY(project(f.Cbuf, out, ya, (float)wsum * dtf));      // sol(x) = y0 + dt sum_s w_s(x) k_s
          }
          ++next_out;
        }
      } else if (next_out < n_t && t[next_out] <= t1) {
        // dense output needs D-wide f0 = k_0, f1 = k_6 and dt * sum c_mid k: three projections, then the
        // reference's quartic in the reference's operation order
        GN_TRY(project(f.cat2[0], b.k0, nullptr, 1.f));
        GN_TRY(project(f.cat2[6], b.k1, nullptr, 1.f));
        double cmsum;
        GN_TRY(combine(tb.c_mid, dtf, &cmsum));
        GN_TRY(project(f.Cbuf, b.xs, nullptr, (float)cmsum * dtf));
        while (next_out < n_t && t[next_out] <= t1) {
          const float x = (float)((t[next_out] - t0) / (t1 - t0));
          LinComb lm{};
          lm.n = n; lm.n_terms = 7;
          for (int q = 0; q < 7; ++q) { lm.in[q] = b.k0; lm.coef[q] = 0.f; }
          lm.in[1] = b.xs; lm.coef[1] = 1.f;      // comb_terms = dt * sum c_mid k  (the other terms add exact zeros)
          lm.in[6] = b.k1;
          GN_TRY(dopri_interp(lm, ya, yb, dtf, x, sol + (int64_t)next_out * n, s));
          ++next_out;
        }
      }
      float* tmp = ya; ya = yb; yb = tmp;
      t_cur = t1;
      if (fsal) {
        if (z_next_valid) { float* z = f.z0; f.z0 = f.z_next; f.z_next = z; }   // Z of stage 6 = Z_0 of the new ya
        else f.z0_ready = false;                                                // kernel-per-op stages: contract again
      }
    }
    double factor;
    if (ratio == 0.f) {
      factor = 10.0;
    } else {
      const double dfactor = (ratio < 1.f) ? 1.0 : 0.2;
      factor = std::fmin(10.0, std::fmax(0.9 / std::pow((double)ratio, 0.2), dfactor));
    }
    dt = dt * factor;
    ++n_steps;
  }
  st.last_dt = dt;
  if (stats) *stats = st;
  return GNODE_OK;
}

}  // namespace gnode

using namespace gnode;

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
namespace {

struct FixedWs {
  float* kbuf[kMaxStages];
  float* xs[kMaxStages];
  float* gk;
  float* gcur;
};

void carve_fixed(Arena& a, Sage3Ctx& c, const Tableau& tb, bool backward, FixedWs& w) {
  const size_t n = (size_t)c.N * c.D;
  std::memset(&w, 0, sizeof(w));
  if (!backward) {
    c.carve(a, 1, false);
    for (int i = 0; i < tb.S - 1; ++i) w.kbuf[i] = a.take<float>(n);
    w.xs[0] = a.take<float>(n);
  } else {
    c.carve(a, tb.S, true);
    for (int i = 0; i < tb.S; ++i) w.kbuf[i] = a.take<float>(n);
    for (int i = 1; i < tb.S; ++i) w.xs[i] = a.take<float>(n);
    w.gk = a.take<float>(n);
    w.gcur = a.take<float>(n);
  }
}

// Per-step save area of a forward pass that will be differentiated: stage inputs xs[1..S-1] ([N, D] each) and
// the per-stage layer intermediates cat1 / cat2 ([N, 2H] each).  Lets the backward skip the stage recompute.
struct Sage3Saver : StepSaver {
  Sage3Ctx* c; int S; size_t n, nc; float* base; float* xs[kMaxStages];
  // every sub-buffer starts on a 256-byte boundary (TMA / 128-bit paths need aligned bases)
  static size_t pad(size_t floats) { return (floats + 63) & ~(size_t)63; }
  size_t step_floats() const { return (size_t)(S - 1) * pad(n) + (size_t)S * 2 * pad(nc); }
  float* const* begin_step(int j) override {
    float* p = base + (size_t)j * step_floats();
    xs[0] = nullptr;
    for (int st = 1; st < S; ++st) { xs[st] = p; p += pad(n); }
    for (int st = 0; st < S; ++st) { c->cat1[st] = p; p += pad(nc); c->cat2[st] = p; p += pad(nc); }
    return xs;
  }
};

}  // namespace

extern "C" size_t gnode_integrate_fixed_save_bytes(int64_t n_nodes, int32_t node_dim, int32_t hidden_dim,
                                                   int32_t method, int32_t n_t) {
  const Tableau* tb = tableau_for(method);
  if (!tb || method == GNODE_DOPRI5 || n_t < 2) return 0;
  if (current_fold()) {
    Sage3Ctx c;
    c.N = n_nodes; c.D = node_dim; c.H = hidden_dim;
    return FoldWs::save_floats_per_step(c, tb->S) * sizeof(float) * (size_t)(n_t - 1);
  }
  Sage3Saver sv{};
  sv.S = tb->S; sv.n = (size_t)n_nodes * node_dim; sv.nc = (size_t)n_nodes * 2 * hidden_dim;
  return sv.step_floats() * sizeof(float) * (size_t)(n_t - 1);
}

extern "C" size_t gnode_integrate_fixed_workspace_bytes(int64_t n_nodes, int32_t node_dim, int32_t hidden_dim,
                                                        int32_t method, int32_t backward) {
  const Tableau* tb = tableau_for(method);
  if (!tb || method == GNODE_DOPRI5) return 0;
  Sage3Ctx c;
  c.N = n_nodes; c.D = node_dim; c.H = hidden_dim;
  Arena a(nullptr, 0);
  if (current_fold()) {
    FoldWs f;
    c.carve(a, tb->S, backward != 0);
    f.carve(a, c, tb->S, backward != 0);
    return a.off;
  }
  FixedWs w;
  carve_fixed(a, c, *tb, backward != 0, w);
  return a.off;
}

extern "C" int gnode_integrate_fixed(const gnode_graph* g, const gnode_sage3_params* p, int32_t method,
                                     const float* y0, const float* t, int32_t n_t, float* sol, void* save,
                                     size_t save_bytes, void* workspace, size_t workspace_bytes,
                                     gnode_stream_t stream) {
  return gnode_integrate_fixed_flags(g, p, method, y0, t, n_t, sol, save, save_bytes, workspace, workspace_bytes, 0, stream);
}

// GraphODE.forward on a fixed grid (scripts/train_gde.py:67-100): the solve AND position_decoder over every time point.
// Folded integrator, n_out <= kDecodeLRMaxOut, 2H = 128: the first time point is decoded from y0 (which also delivers
// sol[0] = y0), every later one from the previous one and the step's 2H-wide combination C -- the D-wide solution is
// written once and never read back.  Otherwise: the solve, then the decoder over sol[1:].
extern "C" size_t gnode_integrate_fixed_decoded_workspace_bytes(int64_t n_nodes, int32_t node_dim, int32_t hidden_dim,
                                                                int32_t method, int32_t n_out) {
  return gnode_integrate_fixed_workspace_bytes(n_nodes, node_dim, hidden_dim, method, 0) +
         align_up(sizeof(float) * (size_t)(n_out > 0 ? n_out : 1) * (2 * (size_t)hidden_dim + 1));
}

extern "C" int gnode_integrate_fixed_decoded(const gnode_graph* g, const gnode_sage3_params* p, int32_t method,
                                             const float* y0, const float* t, int32_t n_t, float* sol, void* save,
                                             size_t save_bytes, const float* dec_w, const float* dec_b, int32_t n_out,
                                             float* traj, void* workspace, size_t workspace_bytes, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_TRY(check_graph(g, "gnode_integrate_fixed_decoded"));
  GN_TRY(check_params(p, "gnode_integrate_fixed_decoded"));
  const Tableau* tb = tableau_for(method);
  GN_ARG(tb && method != GNODE_DOPRI5, "gnode_integrate_fixed_decoded: method %d is not a fixed-grid solver", method);
  GN_ARG(y0 && t && sol && traj && dec_w && dec_b && n_t >= 1 && n_out >= 1, "gnode_integrate_fixed_decoded: null pointer or empty time grid");
  GN_ARG(sol != y0, "gnode_integrate_fixed_decoded: sol aliases y0");
  for (int j = 0; j + 1 < n_t; ++j)
    GN_ARG(t[j + 1] > t[j], "gnode_integrate_fixed_decoded: t must be strictly increasing");
  const int64_t N = g->n_nodes;
  const int D = p->node_dim, H = p->hidden_dim;
  // first time point: decoded from y0, which is copied to sol[0] as it streams through
  GN_TRY(gnode_decoder_fwd_copy(y0, N, D, n_out, dec_w, dec_b, traj, sol, stream));
  if (n_t == 1) return GNODE_OK;
  const bool lowrank = current_fold() && n_out <= kDecodeLRMaxOut && 2 * H == 128;
  if (!lowrank) {
    GN_TRY(gnode_integrate_fixed_flags(g, p, method, y0, t, n_t, sol, save, save_bytes, workspace, workspace_bytes,
                                       GNODE_FIXED_SOL0_BY_CALLER, stream));
    return gnode_decoder_fwd(sol + N * D, (int64_t)(n_t - 1) * N, D, n_out, dec_w, dec_b, traj + N * n_out, stream);
  }
  Sage3Ctx c;
  c.g = *g; c.g_tiles = g->tiles; c.g_tile_err = g->tile_err; c.N = N; c.D = D; c.H = H;
  Arena a(workspace, workspace_bytes);
  FoldWs f;
  c.carve(a, tb->S, false);
  f.carve(a, c, tb->S, false);
  DecodeLR dec{};
  dec.Wd = dec_w; dec.n_out = n_out; dec.traj = traj;
  dec.P = a.take<float>((size_t)n_out * (2 * (size_t)H + 1));
  GN_ARENA_OK(a, "gnode_integrate_fixed_decoded");
  GN_TRY(c.pack(*p, false, s));
  if (save) GN_ARG(save_bytes >= gnode_integrate_fixed_save_bytes(c.N, c.D, c.H, method, n_t),
                   "gnode_integrate_fixed_decoded: save buffer too small (%zu bytes)", save_bytes);
  return integrate_fixed_folded(c, f, *tb, y0, t, n_t, sol, static_cast<float*>(save), s, true, &dec);
}

extern "C" int gnode_integrate_fixed_flags(const gnode_graph* g, const gnode_sage3_params* p, int32_t method,
                                           const float* y0, const float* t, int32_t n_t, float* sol, void* save,
                                           size_t save_bytes, void* workspace, size_t workspace_bytes, int32_t flags,
                                           gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool sol0_by_caller = (flags & GNODE_FIXED_SOL0_BY_CALLER) != 0 && sol != y0;
  GN_TRY(check_graph(g, "gnode_integrate_fixed"));
  GN_TRY(check_params(p, "gnode_integrate_fixed"));
  const Tableau* tb = tableau_for(method);
  GN_ARG(tb && method != GNODE_DOPRI5, "gnode_integrate_fixed: method %d is not a fixed-grid solver", method);
  GN_ARG(y0 && t && sol && n_t >= 1, "gnode_integrate_fixed: null pointer or empty time grid");
  for (int j = 0; j + 1 < n_t; ++j)
    GN_ARG(t[j + 1] > t[j], "gnode_integrate_fixed: t must be strictly increasing");
  Sage3Ctx c;
  c.g = *g; c.g_tiles = g->tiles; c.g_tile_err = g->tile_err; c.N = g->n_nodes; c.D = p->node_dim; c.H = p->hidden_dim;
  Arena a(workspace, workspace_bytes);
  if (current_fold()) {
    FoldWs f;
    c.carve(a, tb->S, false);
    f.carve(a, c, tb->S, false);
    GN_ARENA_OK(a, "gnode_integrate_fixed");
    GN_TRY(c.pack(*p, false, s));
    if (save) GN_ARG(save_bytes >= gnode_integrate_fixed_save_bytes(c.N, c.D, c.H, method, n_t),
                     "gnode_integrate_fixed: save buffer too small (%zu bytes)", save_bytes);
    return integrate_fixed_folded(c, f, *tb, y0, t, n_t, sol, static_cast<float*>(save), s, sol0_by_caller);
  }
  FixedWs w;
  carve_fixed(a, c, *tb, false, w);
  GN_ARENA_OK(a, "gnode_integrate_fixed");
  GN_TRY(c.pack(*p, false, s));
  if (save == nullptr) return integrate_fixed(c, method, y0, t, n_t, sol, w.kbuf, w.xs[0], s, nullptr, sol0_by_caller);
  GN_ARG(save_bytes >= gnode_integrate_fixed_save_bytes(c.N, c.D, c.H, method, n_t),
         "gnode_integrate_fixed: save buffer too small (%zu bytes)", save_bytes);
  Sage3Saver sv{};
  sv.c = &c; sv.S = tb->S; sv.n = (size_t)c.N * c.D; sv.nc = (size_t)c.N * 2 * c.H; sv.base = static_cast<float*>(save);
  return integrate_fixed(c, method, y0, t, n_t, sol, w.kbuf, w.xs[0], s, &sv, sol0_by_caller);
}

extern "C" int gnode_integrate_fixed_bwd(const gnode_graph* g, const gnode_sage3_params* p, int32_t method,
                                         const float* sol, const float* t, int32_t n_t, const float* grad_sol,
                                         float* grad_y0, const gnode_sage3_grads* grads, const void* save,
                                         size_t save_bytes, void* workspace, size_t workspace_bytes,
                                         gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_TRY(check_graph(g, "gnode_integrate_fixed_bwd"));
  GN_TRY(check_params(p, "gnode_integrate_fixed_bwd"));
  const Tableau* tbp = tableau_for(method);
  GN_ARG(tbp && method != GNODE_DOPRI5, "gnode_integrate_fixed_bwd: method %d is not a fixed-grid solver", method);
  GN_ARG(sol && t && grad_sol && n_t >= 1, "gnode_integrate_fixed_bwd: null pointer or empty time grid");
  const Tableau& tb = *tbp;
  const int S = tb.S;
  Sage3Ctx c;
  c.g = *g; c.g_tiles = g->tiles; c.g_tile_err = g->tile_err; c.N = g->n_nodes; c.D = p->node_dim; c.H = p->hidden_dim;
  Arena a(workspace, workspace_bytes);
  if (current_fold()) {
    FoldWs f;
    c.carve(a, S, true);
    f.carve(a, c, S, true);
    GN_ARENA_OK(a, "gnode_integrate_fixed_bwd");
    GN_TRY(c.pack(*p, true, s));
    GN_TRY(c.zero_param_grads(s));
    if (save) GN_ARG(save_bytes >= gnode_integrate_fixed_save_bytes(c.N, c.D, c.H, method, n_t),
                     "gnode_integrate_fixed_bwd: save buffer too small (%zu bytes)", save_bytes);
    GN_TRY(integrate_fixed_folded_bwd(c, f, tb, sol, t, n_t, grad_sol, grad_y0, static_cast<const float*>(save), s));
    if (grads) GN_TRY(c.unpack_grads(*grads, s));
    return GNODE_OK;
  }
  FixedWs w;
  carve_fixed(a, c, tb, true, w);
  GN_ARENA_OK(a, "gnode_integrate_fixed_bwd");
  GN_TRY(c.pack(*p, true, s));
  GN_TRY(c.zero_param_grads(s));
  const int64_t n = c.numel();
  Sage3Saver sv{};
  sv.c = &c; sv.S = S; sv.n = (size_t)n; sv.nc = (size_t)c.N * 2 * c.H;
  sv.base = static_cast<float*>(const_cast<void*>(save));
  if (save) GN_ARG(save_bytes >= gnode_integrate_fixed_save_bytes(c.N, c.D, c.H, method, n_t),
                   "gnode_integrate_fixed_bwd: save buffer too small (%zu bytes)", save_bytes);

  // gcur = cotangent of y_{j+1} (explicit part from grad_sol plus what flowed back from later steps)
  GN_CUDA(cudaMemcpyAsync(w.gcur, grad_sol + (int64_t)(n_t - 1) * n, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
  for (int j = n_t - 2; j >= 0; --j) {
    const float dt = t[j + 1] - t[j];
    const float* y = sol + (int64_t)j * n;
    // ---- recompute the stages of step j, keeping x_s and the layer intermediates per stage ----
    const float* xst[kMaxStages];
    xst[0] = y;
    if (save) {
      // the forward pass kept every stage input and the layer intermediates of this step
      float* const* xs_saved = sv.begin_step(j);
      for (int st = 1; st < S; ++st) xst[st] = xs_saved[st];
    } else {
      for (int st = 0; st < S; ++st) {
        if (st > 0) {
          GN_TRY(stage_input(tb, st, y, w.kbuf, dt, w.xs[st], n, s));
          xst[st] = w.xs[st];
        }
        // k of the last stage is never needed again (y_{j+1} is already known), but its layer
        // intermediates are: evaluate it too, into kbuf[S-1].
        GN_TRY(c.eval(xst[st], w.kbuf[st], nullptr, 1.f, st, s));
      }
    }
    // ---- reverse sweep over the stages; kbuf[s] is reused to hold g_x[s] ----
    for (int st = S - 1; st >= 0; --st) {
      // g_k[st] = dt * c_st * gcur + dt * sum_{i > st} beta[i][st] * g_x[i]
      LinComb lc{};
      lc.out = w.gk; lc.base = nullptr; lc.n = n; lc.n_terms = 0;
      lc.in[lc.n_terms] = w.gcur; lc.coef[lc.n_terms] = (float)tb.c_sol[st] * dt; ++lc.n_terms;
      for (int i = st + 1; i < S; ++i) {
        lc.in[lc.n_terms] = w.kbuf[i]; lc.coef[lc.n_terms] = (float)tb.beta[i][st] * dt; ++lc.n_terms;
      }
      bool any = false;
      for (int q = 0; q < lc.n_terms; ++q) any = any || (lc.coef[q] != 0.f);
      if (!any) {  // stage does not influence the output (cannot happen for the shipped tableaus)
        GN_CUDA(cudaMemsetAsync(w.kbuf[st], 0, sizeof(float) * n, s));
        continue;
      }
      GN_TRY(lincomb(lc, s));
      GN_TRY(c.vjp(xst[st], st, w.gk, w.kbuf[st], s));
    }
    // gcur <- gcur + sum_s g_x[s] + grad_sol[j]
    LinComb lc{};
    lc.out = w.gcur; lc.base = w.gcur; lc.n = n; lc.n_terms = 0;
    for (int st = 0; st < S; ++st) { lc.in[lc.n_terms] = w.kbuf[st]; lc.coef[lc.n_terms] = 1.f; ++lc.n_terms; }
    lc.in[lc.n_terms] = grad_sol + (int64_t)j * n; lc.coef[lc.n_terms] = 1.f; ++lc.n_terms;
    GN_TRY(lincomb(lc, s));
  }
  if (grad_y0) GN_CUDA(cudaMemcpyAsync(grad_y0, w.gcur, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
  if (grads) GN_TRY(c.unpack_grads(*grads, s));
  return GNODE_OK;
}

// ------------------------------------------------------------------------------------------------
// Adjoint backward of the fixed-grid solvers (torchdiffeq odeint_adjoint semantics; the reference trains with plain
// odeint, scripts/train_gde.py:78-85 -- this is the opt-in alternative SURVEY 8-f4 lists).  For every interval
// [t_{i-1}, t_i], last to first, the augmented system
//     d/dt (y, a, g_theta) = ( f(y), -a^T df/dy, -a^T df/dtheta )
// is integrated BACKWARDS over one step of the same Runge-Kutta scheme (h = t_{i-1} - t_i < 0) from (sol[i], a, g_theta);
// then a += grad_sol[i-1] and y is reset to the stored sol[i-1], as torchdiffeq does.  Nothing of the forward pass is
// kept but the solution at the output times: memory O(1) in the number of steps, gradients equal to those of
// backprop-through-the-solver up to the discretisation error of the backward solve.
//
// One vjp per stage yields both parts: with the cotangent (-h c_s) A_s it returns h c_s K^a_s and adds h c_s K^theta_s to
// the parameter gradients (vjp is linear in its cotangent); stages with c_s = 0 (midpoint's first) run a data-only vjp.
// ------------------------------------------------------------------------------------------------
extern "C" size_t gnode_integrate_fixed_adjoint_workspace_bytes(int64_t n_nodes, int32_t node_dim, int32_t hidden_dim,
                                                                int32_t method) {
  const Tableau* tb = tableau_for(method);
  if (!tb || method == GNODE_DOPRI5) return 0;
  Sage3Ctx c;
  c.N = n_nodes; c.D = node_dim; c.H = hidden_dim;
  Arena a(nullptr, 0);
  c.carve(a, 1, true);
  const size_t n = (size_t)n_nodes * node_dim;
  for (int i = 0; i < 2 * tb->S + 4; ++i) a.take<float>(n);
  return a.off;
}

extern "C" int gnode_integrate_fixed_adjoint(const gnode_graph* g, const gnode_sage3_params* p, int32_t method,
                                             const float* sol, const float* t, int32_t n_t, const float* grad_sol,
                                             float* grad_y0, const gnode_sage3_grads* grads, void* workspace,
                                             size_t workspace_bytes, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_TRY(check_graph(g, "gnode_integrate_fixed_adjoint"));
  GN_TRY(check_params(p, "gnode_integrate_fixed_adjoint"));
  const Tableau* tbp = tableau_for(method);
  GN_ARG(tbp && method != GNODE_DOPRI5, "gnode_integrate_fixed_adjoint: method %d is not a fixed-grid solver", method);
  GN_ARG(sol && t && grad_sol && n_t >= 1, "gnode_integrate_fixed_adjoint: null pointer or empty time grid");
  const Tableau& tb = *tbp;
  const int S = tb.S;
  Sage3Ctx c;
  c.g = *g; c.g_tiles = nullptr; c.g_tile_err = nullptr; c.N = g->n_nodes; c.D = p->node_dim; c.H = p->hidden_dim;
  Arena a(workspace, workspace_bytes);
  c.carve(a, 1, true);
  const size_t n = (size_t)c.N * c.D;
  float *Ky[kMaxStages], *Q[kMaxStages];
  for (int i = 0; i < S; ++i) { Ky[i] = a.take<float>(n); Q[i] = a.take<float>(n); }
  float* Ys = a.take<float>(n);
  float* As = a.take<float>(n);
  float* gk = a.take<float>(n);
  float* adj = a.take<float>(n);
  GN_ARENA_OK(a, "gnode_integrate_fixed_adjoint");
  GN_TRY(c.pack(*p, true, s));
  GN_TRY(c.zero_param_grads(s));
  const int64_t nn = (int64_t)n;

  GN_CUDA(cudaMemcpyAsync(adj, grad_sol + (int64_t)(n_t - 1) * nn, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
  for (int i = n_t - 1; i >= 1; --i) {
    const float h = t[i - 1] - t[i];                 // negative: the augmented system runs backwards in time
    const float* y = sol + (int64_t)i * nn;
    double sigma[kMaxStages];                        // K^a_j = sigma_j Q_j
    for (int st = 0; st < S; ++st) {
      const float* Y = y;
      const float* A = adj;
      if (st > 0) {
        GN_TRY(stage_input(tb, st, y, Ky, h, Ys, nn, s));
        Y = Ys;
        LinComb lc{};
        lc.out = As; lc.base = adj; lc.n = nn; lc.n_terms = 0;
        for (int j = 0; j < st; ++j) {
          const float cf = (float)(tb.beta[st][j] * (double)h * sigma[j]);
          if (cf == 0.f) continue;
          lc.in[lc.n_terms] = Q[j]; lc.coef[lc.n_terms] = cf; ++lc.n_terms;
        }
        GN_TRY(lincomb(lc, s));
        A = As;
      }
      GN_TRY(c.eval(Y, Ky[st], nullptr, 1.f, 0, s));           // K^y_s = f(Y_s); intermediates stay in slot 0
      const double cs = tb.c_sol[st];
      LinComb lg{};
      lg.out = gk; lg.base = nullptr; lg.n = nn; lg.n_terms = 1; lg.in[0] = A;
      if (cs != 0.0) { lg.coef[0] = (float)(-(double)h * cs); sigma[st] = 1.0 / ((double)h * cs); c.skip_wgrad = false; }
      else { lg.coef[0] = -1.f; sigma[st] = 1.0; c.skip_wgrad = true; }
      GN_TRY(lincomb(lg, s));
      GN_TRY(c.vjp(Y, 0, gk, Q[st], s));
      c.skip_wgrad = false;
    }
    // a <- a + h sum_s c_s K^a_s (+ the explicit cotangent of the earlier output point)
    LinComb lc{};
    lc.out = adj; lc.base = adj; lc.n = nn; lc.n_terms = 0;
    for (int st = 0; st < S; ++st) {
      const float cf = (float)((double)h * tb.c_sol[st] * sigma[st]);
      if (cf == 0.f) continue;
      lc.in[lc.n_terms] = Q[st]; lc.coef[lc.n_terms] = cf; ++lc.n_terms;
    }
    lc.in[lc.n_terms] = grad_sol + (int64_t)(i - 1) * nn; lc.coef[lc.n_terms] = 1.f; ++lc.n_terms;
    GN_TRY(lincomb(lc, s));
  }
  if (grad_y0) GN_CUDA(cudaMemcpyAsync(grad_y0, adj, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
  if (grads) GN_TRY(c.unpack_grads(*grads, s));
  return GNODE_OK;
}

namespace gnode {

int rk_step_bwd(Field& f, const Tableau& tb, const float* y, float dt, const RkBwdSrc* src, int n_src,
                const RkBwdBufs& b, const float* const* xs_saved, const float* extra, float* gout, cudaStream_t s) {
  const int S = tb.S;
  const int64_t n = f.numel();
  // ---- the stages of the step: stage inputs x_s and the field's intermediates per slot ----
  const float* xst[kMaxStages];
  xst[0] = y;
  if (xs_saved) {
    for (int st = 1; st < S; ++st) xst[st] = xs_saved[st];
  } else {
    for (int st = 0; st < S; ++st) {
      if (st > 0) {
        GN_TRY(stage_input(tb, st, y, b.kbuf, dt, b.xs[st], n, s));
        xst[st] = b.xs[st];
      }
      GN_TRY(f.eval(xst[st], b.kbuf[st], nullptr, 1.f, st, s));
    }
  }
  // ---- reverse sweep; kbuf[s] is reused to hold g_x[s] ----
  for (int st = S - 1; st >= 0; --st) {
    // g_k[st] = dt * sum_q w_q[st] G_q + dt * sum_{i > st} beta[i][st] g_x[i]
    bool any = false;
    LinComb lc{};
    lc.out = b.gk; lc.base = nullptr; lc.n = n; lc.n_terms = 0;
    auto flush = [&]() -> int {
      if (lc.n_terms == 0) return GNODE_OK;
      GN_TRY(lincomb(lc, s));
      lc.base = b.gk; lc.n_terms = 0;
      return GNODE_OK;
    };
    for (int q = 0; q < n_src; ++q) {
      const float cf = (float)src[q].w[st] * dt;
      if (cf == 0.f) continue;
      if (lc.n_terms == kMaxTerms) GN_TRY(flush());
      lc.in[lc.n_terms] = src[q].G; lc.coef[lc.n_terms] = cf; ++lc.n_terms; any = true;
    }
    for (int i = st + 1; i < S; ++i) {
      const float cf = (float)tb.beta[i][st] * dt;
      if (cf == 0.f) continue;
      if (lc.n_terms == kMaxTerms) GN_TRY(flush());
      lc.in[lc.n_terms] = b.kbuf[i]; lc.coef[lc.n_terms] = cf; ++lc.n_terms; any = true;
    }
    if (!any) {   // the stage does not influence any output
      GN_CUDA(cudaMemsetAsync(b.kbuf[st], 0, sizeof(float) * n, s));
      continue;
    }
    GN_TRY(flush());
    GN_TRY(f.vjp(xst[st], st, b.gk, b.kbuf[st], s));
  }
  // gout = sum_q G_q + sum_s g_x[s] (+ extra)
  LinComb lc{};
  lc.out = gout; lc.base = nullptr; lc.n = n; lc.n_terms = 0;
  auto add = [&](const float* p) -> int {
    if (lc.n_terms == kMaxTerms) { GN_TRY(lincomb(lc, s)); lc.base = gout; lc.n_terms = 0; }
    lc.in[lc.n_terms] = p; lc.coef[lc.n_terms] = 1.f; ++lc.n_terms;
    return GNODE_OK;
  };
  for (int q = 0; q < n_src; ++q) GN_TRY(add(src[q].G));        // sources first: gout may alias one of them
  for (int st = 0; st < S; ++st) GN_TRY(add(b.kbuf[st]));
  if (extra) GN_TRY(add(extra));
  GN_TRY(lincomb(lc, s));
  return GNODE_OK;
}

int integrate_fixed_bwd_generic(Field& f, const Tableau& tb, const float* sol, const float* t, int n_t, const float* grad_sol,
                                float* grad_y0, const RkBwdBufs& b, float* gcur, cudaStream_t s) {
  const int64_t n = f.numel();
  GN_CUDA(cudaMemcpyAsync(gcur, grad_sol + (int64_t)(n_t - 1) * n, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
  for (int j = n_t - 2; j >= 0; --j) {
    RkBwdSrc src{};
    src.G = gcur;
    for (int st = 0; st < tb.S; ++st) src.w[st] = tb.c_sol[st];
    GN_TRY(rk_step_bwd(f, tb, sol + (int64_t)j * n, t[j + 1] - t[j], &src, 1, b, nullptr, grad_sol + (int64_t)j * n, gcur, s));
  }
  if (grad_y0) GN_CUDA(cudaMemcpyAsync(grad_y0, gcur, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
  return GNODE_OK;
}

int integrate_dopri5_bwd_generic(Field& f, const float* y0, const double* tau, int n_acc, const double* t, int n_t,
                                 const float* grad_sol, float* grad_y0, const RkBwdBufs& b, float* ys, float* gping,
                                 float* gpong, cudaStream_t s) {
  const Tableau& tb = *tableau_for(GNODE_DOPRI5);
  const int64_t n = f.numel();
  // ---- replay the accepted steps (see fold.cu:integrate_dopri5_folded_bwd for the scheme) ----
  GN_CUDA(cudaMemcpyAsync(ys, y0, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
  for (int k = 0; k + 1 < n_acc; ++k) {
    const float dt = (float)(tau[k + 1] - tau[k]);
    const float* y = ys + (int64_t)k * n;
    for (int st = 0; st < tb.S - 1; ++st) {          // stage 6 is f(y_{k+1}): not needed for y_{k+1} itself
      const float* x = y;
      if (st > 0) { GN_TRY(stage_input(tb, st, y, b.kbuf, dt, b.xs[st], n, s)); x = b.xs[st]; }
      GN_TRY(f.eval(x, b.kbuf[st], nullptr, 1.f, st, s));
    }
    GN_TRY(stage_input(tb, tb.S - 1, y, b.kbuf, dt, ys + (int64_t)(k + 1) * n, n, s));   // beta[6] == c_sol
  }
  const float* Gnext = nullptr;
  float* gout = gping;
  int out_hi = n_t - 1;
  for (int k = n_acc - 1; k >= 0; --k) {
    const double t0 = tau[k], t1 = tau[k + 1];
    int out_lo = out_hi;
    while (out_lo >= 1 && t[out_lo] > t0) --out_lo;
    RkBwdSrc src[kMaxTerms];
    int n_src = 0;
    for (int i = out_lo + 1; i <= out_hi; ++i) {
      if (n_src == kMaxTerms - 1) { set_error("dopri5 backward: more than %d outputs inside one accepted step", kMaxTerms - 1); return GNODE_ERR_SOLVER; }
      src[n_src].G = grad_sol + (int64_t)i * n;
      dopri5_dense_weights(tb, (t[i] - t0) / (t1 - t0), src[n_src].w);
      ++n_src;
    }
    if (Gnext) {
      src[n_src].G = Gnext;
      for (int st = 0; st < tb.S; ++st) src[n_src].w[st] = tb.c_sol[st];
      ++n_src;
    }
    out_hi = out_lo;
    if (n_src == 0) continue;
    GN_TRY(rk_step_bwd(f, tb, ys + (int64_t)k * n, (float)(t1 - t0), src, n_src, b, nullptr, nullptr, gout, s));
    Gnext = gout;
    gout = (gout == gping) ? gpong : gping;
  }
  if (grad_y0) {
    LinComb lc{};
    lc.out = grad_y0; lc.base = grad_sol; lc.n = n; lc.n_terms = 0;
    if (Gnext) { lc.in[0] = Gnext; lc.coef[0] = 1.f; lc.n_terms = 1; }
    GN_TRY(lincomb(lc, s));
  }
  return GNODE_OK;
}

}  // namespace gnode

namespace {
void carve_dopri5_fold(Arena& a, Sage3Ctx& c, FoldWs& f, Dopri5FoldBufs& b) {
  c.carve(a, 7, false);
  f.carve(a, c, 7, false);
  const size_t n = (size_t)c.N * c.D;
  b.ya = a.take<float>(n); b.yb = a.take<float>(n); b.k0 = a.take<float>(n);
  b.k1 = a.take<float>(n); b.xs = a.take<float>(n); b.err = a.take<float>(n);
  b.Cerr = a.take<float>((size_t)c.N * 2 * c.H);
  b.znext = a.take<float>((size_t)c.N * 2 * c.H);
  b.partials = a.take<double>((size_t)norm_blocks((int64_t)n));
  b.dsum = a.take<double>(2);
}
}  // namespace

extern "C" int gnode_set_dopri5_device_allreduce(gnode_allreduce_dev_fn fn, void* user) {
  t_allreduce_dev = fn;
  t_allreduce_dev_user = user;
  return GNODE_OK;
}

extern "C" size_t gnode_integrate_dopri5_workspace_bytes(int64_t n_nodes, int32_t node_dim, int32_t hidden_dim) {
  Sage3Ctx c;
  c.N = n_nodes; c.D = node_dim; c.H = hidden_dim;
  Arena a(nullptr, 0);
  if (current_fold()) {
    FoldWs f;
    Dopri5FoldBufs b{};
    carve_dopri5_fold(a, c, f, b);
    return a.off;
  }
  c.carve(a, 1, false);
  const size_t n = (size_t)n_nodes * node_dim;
  for (int i = 0; i < 10; ++i) a.take<float>(n);
  a.take<double>((size_t)norm_blocks((int64_t)n));
  a.take<double>(2);
  return a.off;
}

extern "C" int gnode_integrate_dopri5(const gnode_graph* g, const gnode_sage3_params* p, const float* y0,
                                      const double* t, int32_t n_t, double rtol, double atol, float* sol,
                                      gnode_dopri5_stats* stats, const gnode_dopri5_trace* trace,
                                      gnode_allreduce_fn allreduce, void* allreduce_user, int64_t max_num_steps,
                                      void* workspace, size_t workspace_bytes, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_TRY(check_graph(g, "gnode_integrate_dopri5"));
  GN_TRY(check_params(p, "gnode_integrate_dopri5"));
  GN_ARG(y0 && t && sol && n_t >= 1, "gnode_integrate_dopri5: null pointer or empty time grid");
  GN_ARG(rtol > 0 || atol > 0, "gnode_integrate_dopri5: rtol and atol are both zero");
  for (int j = 0; j + 1 < n_t; ++j)
    GN_ARG(t[j + 1] > t[j], "gnode_integrate_dopri5: t must be strictly increasing");
  Sage3Ctx c;
  c.g = *g; c.g_tiles = g->tiles; c.g_tile_err = g->tile_err; c.N = g->n_nodes; c.D = p->node_dim; c.H = p->hidden_dim;
  Arena a(workspace, workspace_bytes);
  if (max_num_steps <= 0) max_num_steps = (1ll << 31) - 1;
  if (current_fold()) {
    FoldWs f;
    Dopri5FoldBufs fb{};
    carve_dopri5_fold(a, c, f, fb);
    GN_ARENA_OK(a, "gnode_integrate_dopri5");
    GN_TRY(c.pack(*p, false, s));
    return integrate_dopri5_folded(c, f, y0, t, n_t, rtol, atol, sol, stats, trace, allreduce, allreduce_user,
                                   max_num_steps, fb, s);
  }
  c.carve(a, 1, false);
  const size_t n = (size_t)c.N * c.D;
  Dopri5Bufs b{};
  for (int i = 0; i < 7; ++i) b.k[i] = a.take<float>(n);
  b.ya = a.take<float>(n);
  b.yb = a.take<float>(n);
  b.xs = a.take<float>(n);
  b.partials = a.take<double>((size_t)norm_blocks((int64_t)n));
  b.dsum = a.take<double>(2);
  GN_ARENA_OK(a, "gnode_integrate_dopri5");
  GN_TRY(c.pack(*p, false, s));
  if (max_num_steps <= 0) max_num_steps = (1ll << 31) - 1;
  return integrate_dopri5(c, y0, t, n_t, rtol, atol, sol, stats, trace, allreduce, allreduce_user, max_num_steps, b, s);
}

// Backward through dopri5 over the accepted steps of the forward pass (fold.cu:integrate_dopri5_folded_bwd).
extern "C" size_t gnode_integrate_dopri5_bwd_workspace_bytes(int64_t n_nodes, int32_t node_dim, int32_t hidden_dim,
                                                             int32_t n_accepted) {
  Sage3Ctx c;
  c.N = n_nodes; c.D = node_dim; c.H = hidden_dim;
  Arena a(nullptr, 0);
  FoldWs f;
  c.carve(a, 7, true);
  f.carve(a, c, 7, true);
  a.take<float>((size_t)(n_accepted > 0 ? n_accepted : 1) * (size_t)n_nodes * node_dim);
  return a.off;
}

extern "C" int gnode_integrate_dopri5_bwd(const gnode_graph* g, const gnode_sage3_params* p, const float* y0,
                                          const double* tau, int32_t n_accepted, const double* t, int32_t n_t,
                                          const float* grad_sol, float* grad_y0, const gnode_sage3_grads* grads,
                                          void* workspace, size_t workspace_bytes, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_TRY(check_graph(g, "gnode_integrate_dopri5_bwd"));
  GN_TRY(check_params(p, "gnode_integrate_dopri5_bwd"));
  GN_ARG(y0 && tau && t && grad_sol && n_t >= 1 && n_accepted >= 0, "gnode_integrate_dopri5_bwd: bad argument");
  GN_ARG(current_fold(), "gnode_integrate_dopri5_bwd: needs the folded integrator (gnode_set_fold(1))");
  for (int k = 0; k < n_accepted; ++k)
    GN_ARG(tau[k + 1] > tau[k], "gnode_integrate_dopri5_bwd: accepted step times must be strictly increasing");
  GN_ARG(n_t == 1 || (n_accepted >= 1 && tau[0] == t[0] && tau[n_accepted] >= t[n_t - 1]),
         "gnode_integrate_dopri5_bwd: the accepted steps do not cover the time grid");
  Sage3Ctx c;
  c.g = *g; c.g_tiles = g->tiles; c.g_tile_err = g->tile_err; c.N = g->n_nodes; c.D = p->node_dim; c.H = p->hidden_dim;
  Arena a(workspace, workspace_bytes);
  FoldWs f;
  c.carve(a, 7, true);
  f.carve(a, c, 7, true);
  float* ys = a.take<float>((size_t)(n_accepted > 0 ? n_accepted : 1) * (size_t)c.N * c.D);
  GN_ARENA_OK(a, "gnode_integrate_dopri5_bwd");
  GN_TRY(c.pack(*p, true, s));
  GN_TRY(c.zero_param_grads(s));
  GN_TRY(integrate_dopri5_folded_bwd(c, f, y0, tau, n_accepted, t, n_t, grad_sol, grad_y0, ys, s));
  if (grads) GN_TRY(c.unpack_grads(*grads, s));
  return GNODE_OK;
}

// Backward of a one-step fixed-grid solve whose solution reaches the loss only through position_decoder at the last
// time point (the training step of scripts/train_gde.py:486-493): dL/dy_1 = grad_traj_last @ dec_w has rank n_out, so
// neither it nor its two D-wide contractions are ever formed.
extern "C" size_t gnode_integrate_fixed_bwd_decoded_workspace_bytes(int64_t n_nodes, int32_t node_dim, int32_t hidden_dim,
                                                                    int32_t method, int32_t n_out) {
  const size_t base = gnode_integrate_fixed_workspace_bytes(n_nodes, node_dim, hidden_dim, method, 1);
  Arena a(nullptr, 0);
  a.take<float>((size_t)n_out * 2 * hidden_dim);
  a.take<float>((size_t)n_out * 2 * hidden_dim + n_out);
  a.take<float>(decoder_wgrad_partial_floats(n_nodes, 2 * hidden_dim, n_out));
  return base + a.off + 256;
}

extern "C" int gnode_integrate_fixed_bwd_decoded(const gnode_graph* g, const gnode_sage3_params* p, int32_t method,
                                                 const float* sol, const float* t, int32_t n_t,
                                                 const float* grad_traj_last, const float* dec_w, int32_t n_out,
                                                 const gnode_sage3_grads* grads, const void* save, size_t save_bytes,
                                                 void* workspace, size_t workspace_bytes, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_TRY(check_graph(g, "gnode_integrate_fixed_bwd_decoded"));
  GN_TRY(check_params(p, "gnode_integrate_fixed_bwd_decoded"));
  const Tableau* tbp = tableau_for(method);
  GN_ARG(tbp && method != GNODE_DOPRI5, "gnode_integrate_fixed_bwd_decoded: method %d is not a fixed-grid solver", method);
  GN_ARG(current_fold(), "gnode_integrate_fixed_bwd_decoded: needs the folded integrator (gnode_set_fold(1))");
  GN_ARG(n_t == 2, "gnode_integrate_fixed_bwd_decoded: exactly one solver step (n_t == 2) is supported, got n_t = %d", n_t);
  GN_ARG(sol && t && grad_traj_last && dec_w && n_out >= 1 && n_out <= 8, "gnode_integrate_fixed_bwd_decoded: bad argument");
  const Tableau& tb = *tbp;
  Sage3Ctx c;
  c.g = *g; c.g_tiles = g->tiles; c.g_tile_err = g->tile_err; c.N = g->n_nodes; c.D = p->node_dim; c.H = p->hidden_dim;
  Arena a(workspace, workspace_bytes);
  FoldWs f;
  c.carve(a, tb.S, true);
  f.carve(a, c, tb.S, true);
  LowRankG lr{};
  lr.g1 = grad_traj_last; lr.Wd = dec_w; lr.n_out = n_out;
  lr.WdW3 = a.take<float>((size_t)n_out * 2 * c.H);
  lr.X = a.take<float>((size_t)n_out * 2 * c.H + n_out);
  lr.partials = a.take<float>(decoder_wgrad_partial_floats(c.N, 2 * c.H, n_out));
  GN_ARENA_OK(a, "gnode_integrate_fixed_bwd_decoded");
  GN_TRY(c.pack(*p, true, s));
  GN_TRY(c.zero_param_grads(s));
  if (save) GN_ARG(save_bytes >= gnode_integrate_fixed_save_bytes(c.N, c.D, c.H, method, n_t),
                   "gnode_integrate_fixed_bwd_decoded: save buffer too small (%zu bytes)", save_bytes);
  GN_TRY(integrate_fixed_folded_bwd(c, f, tb, sol, t, n_t, nullptr, nullptr, static_cast<const float*>(save), s, &lr));
  if (grads) GN_TRY(c.unpack_grads(*grads, s));
  return GNODE_OK;
}
