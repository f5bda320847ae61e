// GraphODEFunc on the GPU: dx/dt = conv3(relu(conv2(relu(conv1(x)))))  (scripts/train_gde.py:33-45),
// each conv a mean-aggregating SAGEConv [upstream PyG].
//
// Uses linearity of the mean aggregation A(.) so that the sparse gather only ever touches
// hidden-width (H) rows:
//   conv1:  Z = x @ [w1l; w1r]^T  (N x 2H GEMM over the D-wide state) ;  h1 = relu(A(Z_l) + Z_r + b1)
//   conv2:  h2 = relu([A(h1) | h1] @ [w2l | w2r]^T + b2)
//   conv3:  k  = [A(h2) | h2] @ [w3l | w3r]^T + b3      (N x D GEMM, stage-combination epilogue)
// This is a re-association of the reference arithmetic (fp32 rounding order only).
#include "field.cuh"

namespace gnode {

void Sage3Ctx::carve(Arena& a, int slots, bool backward) {
  const size_t n = (size_t)N;
  n_slots = slots;
  w1cat = a.take<float>((size_t)2 * H * D);
  w2cat = a.take<float>((size_t)H * 2 * H);
  w3cat = a.take<float>((size_t)D * 2 * H);
  s1 = a.take<float>(presplit_floats(2 * H, D));
  s2 = a.take<float>(presplit_floats(H, 2 * H));
  s3 = a.take<float>(presplit_floats(D, 2 * H));
  if (chain_shape_ok(H)) ci2 = a.take<float>(chain_image_floats(H, 2 * H));
  if (chain_shape_ok(H)) ck3 = a.take<float>(gemm_k128_image_floats(D));
  z = a.take<float>(n * 2 * H);
  for (int i = 0; i < slots; ++i) cat1[i] = a.take<float>(n * 2 * H);   // two stacks (stage-major): the backward pass
  for (int i = 0; i < slots; ++i) cat2[i] = a.take<float>(n * 2 * H);   // contracts over all stages at once
  if (backward) {
    w1catT = a.take<float>((size_t)2 * H * D);
    w2catT = a.take<float>((size_t)H * 2 * H);
    w3catT = a.take<float>((size_t)D * 2 * H);
    s1T = a.take<float>(presplit_floats(D, 2 * H));
    s2T = a.take<float>(presplit_floats(2 * H, H));
    s3T = a.take<float>(presplit_floats(2 * H, D));
    if (chain_shape_ok(H)) ci2T = a.take<float>(chain_image_floats(2 * H, H));
    if (chain_shape_ok(H)) ck1T = a.take<float>(gemm_k128_image_floats(D));
    gcat = a.take<float>(n * 2 * H);
    gz = a.take<float>(n * 2 * H);
    gv2 = a.take<float>(n * H);
    size_t pf = gemm_tn_workspace_floats(D, 2 * H, N);
    size_t p2 = gemm_tn_workspace_floats(H, 2 * H, N);
    size_t p3 = gemm_tn_workspace_floats(2 * H, D, N);
    if (p2 > pf) pf = p2;
    if (p3 > pf) pf = p3;
    partials = a.take<float>(pf);
    colpart = a.take<float>(colsum_workspace_floats(D > 2 * H ? D : 2 * H, N));
    dW1cat = a.take<float>((size_t)2 * H * D);
    dW2cat = a.take<float>((size_t)H * 2 * H);
    dW3cat = a.take<float>((size_t)D * 2 * H);
    db1 = a.take<float>(H);
    db2 = a.take<float>(H);
    db3 = a.take<float>(D);
  }
}

int Sage3Ctx::pack(const gnode_sage3_params& p, bool backward, cudaStream_t s) {
  b1 = p.b1; b2 = p.b2; b3 = p.b3;
  PackSegHost sg[12];
  int n = 0;
  auto add = [&](float* dst, const float* src, int rows, int cols, int64_t ld_dst, int transpose) {
    sg[n++] = PackSegHost{dst, src, rows, cols, (int64_t)cols, ld_dst, transpose, 0};
  };
  add(w1cat, p.w1l, H, D, D, 0);
  add(w1cat + (size_t)H * D, p.w1r, H, D, D, 0);
  add(w2cat, p.w2l, H, H, 2 * H, 0);
  add(w2cat + H, p.w2r, H, H, 2 * H, 0);
  add(w3cat, p.w3l, D, H, 2 * H, 0);
  add(w3cat + H, p.w3r, D, H, 2 * H, 0);
  if (backward) {
    add(w3catT, p.w3l, D, H, D, 1);                    // w3catT[c, r] = w3l[r, c]
    add(w3catT + (size_t)H * D, p.w3r, D, H, D, 1);
    add(w2catT, p.w2l, H, H, H, 1);
    add(w2catT + (size_t)H * H, p.w2r, H, H, H, 1);
    add(w1catT, p.w1l, H, D, 2 * H, 1);                // w1catT[c, r] = w1l[r, c]
    add(w1catT + H, p.w1r, H, D, 2 * H, 1);
  }
  GN_TRY(pack_segments(sg, n, s));
  use_tc = current_engine() != GNODE_ENGINE_SIMT;
  // the tcgen05 images of the packed matrices are made by whoever reads them first (see field.cuh)
  pend_s1 = pend_s2 = pend_s3 = use_tc ? 1 : 0;
  pend_ci2 = pend_ck3 = (use_tc && chain_shape_ok(H)) ? 1 : 0;
  pend_s1T = pend_s2T = pend_s3T = (use_tc && backward) ? 1 : 0;
  pend_ci2T = pend_ck1T = (use_tc && backward && chain_shape_ok(H)) ? 1 : 0;
  if (use_tc && lazy_images_poisoned()) {   // GNODE_POISON_LAZY=1 (tests): a reader that skips the flag sees NaNs, not stale images
    GN_CUDA(cudaMemsetAsync(s1, 0xFF, sizeof(float) * presplit_floats(2 * H, D), s));
    GN_CUDA(cudaMemsetAsync(s2, 0xFF, sizeof(float) * presplit_floats(H, 2 * H), s));
    GN_CUDA(cudaMemsetAsync(s3, 0xFF, sizeof(float) * presplit_floats(D, 2 * H), s));
    if (chain_shape_ok(H)) {
      GN_CUDA(cudaMemsetAsync(ci2, 0xFF, sizeof(float) * chain_image_floats(H, 2 * H), s));
      GN_CUDA(cudaMemsetAsync(ck3, 0xFF, sizeof(float) * gemm_k128_image_floats(D), s));
    }
    if (backward) {
      GN_CUDA(cudaMemsetAsync(s1T, 0xFF, sizeof(float) * presplit_floats(D, 2 * H), s));
      GN_CUDA(cudaMemsetAsync(s2T, 0xFF, sizeof(float) * presplit_floats(2 * H, H), s));
      GN_CUDA(cudaMemsetAsync(s3T, 0xFF, sizeof(float) * presplit_floats(2 * H, D), s));
      if (chain_shape_ok(H)) {
        GN_CUDA(cudaMemsetAsync(ci2T, 0xFF, sizeof(float) * chain_image_floats(2 * H, H), s));
        GN_CUDA(cudaMemsetAsync(ck1T, 0xFF, sizeof(float) * gemm_k128_image_floats(D), s));
      }
    }
  }
  return GNODE_OK;
}

int Sage3Ctx::zero_param_grads(cudaStream_t s) {
  // dW1cat .. db3 are consecutive arena blocks (carve): one fill covers all six and the padding between them
  const size_t span = (size_t)(reinterpret_cast<char*>(db3 + D) - reinterpret_cast<char*>(dW1cat));
  GN_CUDA(cudaMemsetAsync(dW1cat, 0, span, s));
  return GNODE_OK;
}

int Sage3Ctx::unpack_grads(const gnode_sage3_grads& gr, cudaStream_t s) {
  PackSegHost sg[9];
  int n = 0;
  auto add = [&](float* dst, const float* src, int rows, int cols, int64_t ld_src) {
    if (dst) sg[n++] = PackSegHost{dst, src, rows, cols, ld_src, (int64_t)cols, 0, 1};
  };
  add(gr.w1l, dW1cat, H, D, D);
  add(gr.w1r, dW1cat + (size_t)H * D, H, D, D);
  add(gr.b1, db1, 1, H, H);
  add(gr.w2l, dW2cat, H, H, 2 * H);
  add(gr.w2r, dW2cat + H, H, H, 2 * H);
  add(gr.b2, db2, 1, H, H);
  add(gr.w3l, dW3cat, D, H, 2 * H);
  add(gr.w3r, dW3cat + H, D, H, 2 * H);
  add(gr.b3, db3, 1, D, D);
  return pack_segments(sg, n, s);
}

int Sage3Ctx::eval(const float* x, float* out, const float* base, float scale, int slot, cudaStream_t s) {
  float* c1 = cat1[slot];
  float* c2 = cat2[slot];
  const int H2 = 2 * H;
  {  // Z = x @ w1cat^T
    GemmNT q{};
    q.A = x; q.lda = D; q.B = w1cat; q.ldb = D; q.C = z; q.ldc = H2; q.M = N; q.N = H2; q.K = D;
    use_w1(q);
    GN_TRY(gemm_nt(q, s));
  }
  // h1 = relu(A(Z_l) + Z_r + b1) -> cat1[:, H:]
  GN_TRY(agg_mean_fwd(g, z, H2, c1 + H, H2, H, z + H, H2, b1, 1, s));
  // A(h1) -> cat1[:, :H]
  GN_TRY(agg_mean_fwd(g, c1 + H, H2, c1, H2, H, nullptr, 0, nullptr, 0, s));
  {  // h2 = relu(cat1 @ w2cat^T + b2) -> cat2[:, H:]
    GemmNT q{};
    q.A = c1; q.lda = H2; q.B = w2cat; q.ldb = H2; q.C = c2 + H; q.ldc = H2; q.M = N; q.N = H; q.K = H2;
    q.bias = b2; q.relu = 1;
    use_w2(q);
    GN_TRY(gemm_nt(q, s));
  }
  // A(h2) -> cat2[:, :H]
  GN_TRY(agg_mean_fwd(g, c2 + H, H2, c2, H2, H, nullptr, 0, nullptr, 0, s));
  {  // k = cat2 @ w3cat^T + b3 ; out = base + scale * k
    GemmNT q{};
    q.A = c2; q.lda = H2; q.B = w3cat; q.ldb = H2; q.C = out; q.ldc = D; q.M = N; q.N = D; q.K = H2;
    q.bias = b3; q.base = base; q.ldbase = D; q.scale = scale;
    use_w3(q);
    GN_TRY(gemm_nt(q, s));
  }
  return GNODE_OK;
}

int Sage3Ctx::vjp(const float* x, int slot, const float* gk, float* gx, cudaStream_t s) {
  const float* c1 = cat1[slot];
  const float* c2 = cat2[slot];
  const int H2 = 2 * H;
  // ---- conv3 ----
  {  // gcat = gk @ w3cat          [N, 2H]
    GemmNT q{};
    q.A = gk; q.lda = D; q.B = w3catT; q.ldb = D; q.C = gcat; q.ldc = H2; q.M = N; q.N = H2; q.K = D;
    use_w3T(q);
    GN_TRY(gemm_nt(q, s));
  }
  if (!skip_wgrad) {  // dW3cat += gk^T @ cat2      [D, 2H]
    GemmTN q{};
    q.A = gk; q.lda = D; q.P = D; q.B = c2; q.ldb = H2; q.Q = H2; q.Nrows = N; q.C = dW3cat; q.ldc = H2;
    q.colsumA = db3;                                   // db3 += colsum(gk), fused into the same pass
    GN_TRY(gemm_tn(q, partials, s));
  }
  // g_v2 = (A^T(gcat_l) + gcat_r) * [h2 > 0]
  GN_TRY(agg_mean_bwd(g, gcat, H2, gv2, H, H, gcat + H, H2, c2 + H, H2, s));
  // ---- conv2 ----
  {  // gcat = g_v2 @ w2cat        [N, 2H]
    GemmNT q{};
    q.A = gv2; q.lda = H; q.B = w2catT; q.ldb = H; q.C = gcat; q.ldc = H2; q.M = N; q.N = H2; q.K = H;
    use_w2T(q);
    GN_TRY(gemm_nt(q, s));
  }
  if (!skip_wgrad) {  // dW2cat += g_v2^T @ cat1    [H, 2H]
    GemmTN q{};
    q.A = gv2; q.lda = H; q.P = H; q.B = c1; q.ldb = H2; q.Q = H2; q.Nrows = N; q.C = dW2cat; q.ldc = H2;
    q.colsumA = db2;
    GN_TRY(gemm_tn(q, partials, s));
  }
  // g_u1 = (A^T(gcat_l) + gcat_r) * [h1 > 0]  -> gz[:, H:] ;  A^T(g_u1) -> gz[:, :H]
  GN_TRY(agg_mean_bwd(g, gcat, H2, gz + H, H2, H, gcat + H, H2, c1 + H, H2, s));
  GN_TRY(agg_mean_bwd(g, gz + H, H2, gz, H2, H, nullptr, 0, nullptr, 0, s));
  // ---- conv1 ----
  {  // gx = gz @ w1cat            [N, D]
    GemmNT q{};
    q.A = gz; q.lda = H2; q.B = w1catT; q.ldb = H2; q.C = gx; q.ldc = D; q.M = N; q.N = D; q.K = H2;
    use_w1T(q);
    GN_TRY(gemm_nt(q, s));
  }
  if (!skip_wgrad) {  // dW1cat += gz^T @ x         [2H, D]
    GemmTN q{};
    q.A = gz; q.lda = H2; q.P = H2; q.B = x; q.ldb = D; q.Q = D; q.Nrows = N; q.C = dW1cat; q.ldc = D;
    GN_TRY(gemm_tn(q, partials, s));
    GN_TRY(colsum_accum(gz + H, H2, N, H, db1, 1.f, colpart, s));
  }
  return GNODE_OK;
}

int check_graph(const gnode_graph* g, const char* who) {
  GN_ARG(g != nullptr, "%s: graph is null", who);
  GN_ARG(g->n_nodes > 0, "%s: graph has no nodes", who);
  GN_ARG(g->rowptr && g->t_rowptr, "%s: graph rowptr is null", who);
  GN_ARG(g->n_edges == 0 || (g->col && g->t_col), "%s: graph col is null", who);
  return GNODE_OK;
}

int check_params(const gnode_sage3_params* p, const char* who) {
  GN_ARG(p != nullptr, "%s: params is null", who);
  GN_ARG(p->node_dim > 0 && p->hidden_dim > 0, "%s: node_dim / hidden_dim must be positive", who);
  GN_ARG(p->hidden_dim % 4 == 0, "%s: hidden_dim must be a multiple of 4 (got %d)", who, p->hidden_dim);
  GN_ARG(p->w1l && p->b1 && p->w1r && p->w2l && p->b2 && p->w2r && p->w3l && p->b3 && p->w3r,
         "%s: null parameter pointer", who);
  return GNODE_OK;
}

}  // namespace gnode

using namespace gnode;

// ------------------------------------------------------------------------------------------------
// C ABI: single RHS evaluation and its vjp
// ------------------------------------------------------------------------------------------------
extern "C" size_t gnode_rhs_workspace_bytes(int64_t n_nodes, int32_t node_dim, int32_t hidden_dim) {
  Sage3Ctx c;
  c.N = n_nodes; c.D = node_dim; c.H = hidden_dim;
  Arena a(nullptr, 0);
  c.carve(a, 1, true);
  return a.off;
}

extern "C" int gnode_rhs_fwd(const gnode_graph* g, const gnode_sage3_params* p, const float* x, float* dxdt,
                             void* workspace, size_t workspace_bytes, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_TRY(check_graph(g, "gnode_rhs_fwd"));
  GN_TRY(check_params(p, "gnode_rhs_fwd"));
  GN_ARG(x && dxdt, "gnode_rhs_fwd: null state pointer");
  Sage3Ctx c;
  c.g = *g; c.g_tiles = g->tiles; c.g_tile_err = g->tile_err; c.N = g->n_nodes; c.D = p->node_dim; c.H = p->hidden_dim;
  Arena a(workspace, workspace_bytes);
  c.carve(a, 1, false);
  GN_ARENA_OK(a, "gnode_rhs_fwd");
  GN_TRY(c.pack(*p, false, s));
  return c.eval(x, dxdt, nullptr, 1.f, 0, s);
}

extern "C" int gnode_rhs_bwd(const gnode_graph* g, const gnode_sage3_params* p, const float* x,
                             const float* grad_out, float* grad_x, const gnode_sage3_grads* grads,
                             void* workspace, size_t workspace_bytes, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_TRY(check_graph(g, "gnode_rhs_bwd"));
  GN_TRY(check_params(p, "gnode_rhs_bwd"));
  GN_ARG(x && grad_out && grad_x, "gnode_rhs_bwd: null pointer");
  Sage3Ctx c;
  c.g = *g; c.g_tiles = g->tiles; c.g_tile_err = g->tile_err; c.N = g->n_nodes; c.D = p->node_dim; c.H = p->hidden_dim;
  Arena a(workspace, workspace_bytes);
  c.carve(a, 1, true);
  GN_ARENA_OK(a, "gnode_rhs_bwd");
  GN_TRY(c.pack(*p, true, s));
  GN_TRY(c.zero_param_grads(s));
  // recompute the forward intermediates; the field value itself lands in grad_x and is overwritten
  GN_TRY(c.eval(x, grad_x, nullptr, 1.f, 0, s));
  GN_TRY(c.vjp(x, 0, grad_out, grad_x, s));
  if (grads) GN_TRY(c.unpack_grads(*grads, s));
  return GNODE_OK;
}
