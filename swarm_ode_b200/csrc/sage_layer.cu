// One SAGEConv layer as a stand-alone operator (reference: SAGEConv(in, out)(x, edge_index),
// scripts/train_gde.py:27-29; hetero sites scripts/gnode.py:92-97).
//   out = act( mean_{j in N(i)} x_j @ wl^T + bl + x_i @ wr^T )
// Forward picks the cheaper association: project-then-aggregate when c_out <= c_in, aggregate-then-
// project otherwise, so the sparse gather always runs at min(c_in, c_out) width.
#include "field.cuh"

namespace gnode {
namespace {

struct SageWs {
  float *wcat, *wcatT, *buf, *gz, *gm, *dwcat, *partials, *colpart;
};

// forward: wcat + buf ; backward: wcatT + gz + gm + dwcat + partials + colpart
void carve_sage(Arena& a, int64_t N, int ci, int co, SageWs& w) {
  const size_t n = (size_t)N;
  const int wide = 2 * (co <= ci ? co : ci);
  w.wcat = a.take<float>((size_t)2 * ci * co);
  w.buf = a.take<float>(n * wide);
  w.wcatT = a.take<float>((size_t)2 * ci * co);
  w.gz = a.take<float>(n * 2 * co);
  w.gm = a.take<float>(n * co);
  w.dwcat = a.take<float>((size_t)2 * ci * co);
  w.partials = a.take<float>(gemm_tn_workspace_floats(2 * co, ci, N));
  w.colpart = a.take<float>(colsum_workspace_floats(co, N));
}

}  // namespace
}  // namespace gnode

using namespace gnode;

extern "C" size_t gnode_sage_workspace_bytes(int64_t n_nodes, int32_t c_in, int32_t c_out) {
  Arena a(nullptr, 0);
  SageWs w;
  carve_sage(a, n_nodes, c_in, c_out, w);
  return a.off;
}

extern "C" int gnode_sage_fwd(const gnode_graph* g, const float* x, int32_t ci, int32_t co, const float* wl,
                              const float* bl, const float* wr, int32_t relu, float* out, void* workspace,
                              size_t workspace_bytes, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_TRY(check_graph(g, "gnode_sage_fwd"));
  GN_ARG(ci > 0 && co > 0 && x && wl && wr && out, "gnode_sage_fwd: bad argument");
  const int64_t N = g->n_nodes;
  Arena a(workspace, workspace_bytes);
  SageWs w;
  carve_sage(a, N, ci, co, w);
  GN_ARENA_OK(a, "gnode_sage_fwd");
  if (co <= ci) {
    // Z = x @ [wl; wr]^T  -> out = act(A(Z_l) + Z_r + bl)
    PackSegHost sg[2] = {{w.wcat, wl, co, ci, ci, ci, 0, 0}, {w.wcat + (size_t)co * ci, wr, co, ci, ci, ci, 0, 0}};
    GN_TRY(pack_segments(sg, 2, s));
    GemmNT q{};
    q.A = x; q.lda = ci; q.B = w.wcat; q.ldb = ci; q.C = w.buf; q.ldc = 2 * co; q.M = N; q.N = 2 * co; q.K = ci;
    GN_TRY(gemm_nt(q, s));
    GN_TRY(agg_mean_fwd(*g, w.buf, 2 * co, out, co, co, w.buf + co, 2 * co, bl, relu ? 1 : 0, s));
  } else {
    // cat = [A(x) | x] -> out = act(cat @ [wl | wr]^T + bl)
    PackSegHost sg[3] = {{w.wcat, wl, co, ci, ci, 2 * ci, 0, 0},
                         {w.wcat + ci, wr, co, ci, ci, 2 * ci, 0, 0},
                         {w.buf + ci, x, (int)N, ci, ci, 2 * ci, 0, 0}};
    GN_TRY(pack_segments(sg, 3, s));
    GN_TRY(agg_mean_fwd(*g, x, ci, w.buf, 2 * ci, ci, nullptr, 0, nullptr, 0, s));
    GemmNT q{};
    q.A = w.buf; q.lda = 2 * ci; q.B = w.wcat; q.ldb = 2 * ci; q.C = out; q.ldc = co; q.M = N; q.N = co; q.K = 2 * ci;
    q.bias = bl; q.relu = relu ? 1 : 0;
    GN_TRY(gemm_nt(q, s));
  }
  return GNODE_OK;
}

extern "C" int gnode_sage_bwd(const gnode_graph* g, const float* x, const float* out, const float* grad_out,
                              int32_t ci, int32_t co, const float* wl, const float* wr, int32_t relu, float* grad_x,
                              float* grad_wl, float* grad_bl, float* grad_wr, void* workspace, size_t workspace_bytes,
                              gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_TRY(check_graph(g, "gnode_sage_bwd"));
  GN_ARG(ci > 0 && co > 0 && x && grad_out && wl && wr, "gnode_sage_bwd: bad argument");
  GN_ARG(!relu || out, "gnode_sage_bwd: the layer output is required for the ReLU mask");
  const int64_t N = g->n_nodes;
  Arena a(workspace, workspace_bytes);
  SageWs w;
  carve_sage(a, N, ci, co, w);
  GN_ARENA_OK(a, "gnode_sage_bwd");
  // gz = [A^T(gm) | gm] with gm = grad_out * act'(out)
  const float* gm = grad_out;
  if (relu) {
    GN_TRY(relu_mask(grad_out, out, w.gm, N * (int64_t)co, s));
    gm = w.gm;
  }
  {
    PackSegHost sg[1] = {{w.gz + co, gm, (int)N, co, co, 2 * co, 0, 0}};
    GN_TRY(pack_segments(sg, 1, s));
  }
  GN_TRY(agg_mean_bwd(*g, gm, co, w.gz, 2 * co, co, nullptr, 0, nullptr, 0, s));
  if (grad_x) {
    // grad_x = gz @ [wl; wr]   -> NT with wcatT [ci, 2co]
    PackSegHost sg[2] = {{w.wcatT, wl, co, ci, ci, 2 * co, 1, 0}, {w.wcatT + co, wr, co, ci, ci, 2 * co, 1, 0}};
    GN_TRY(pack_segments(sg, 2, s));
    GemmNT q{};
    q.A = w.gz; q.lda = 2 * co; q.B = w.wcatT; q.ldb = 2 * co; q.C = grad_x; q.ldc = ci; q.M = N; q.N = ci; q.K = 2 * co;
    GN_TRY(gemm_nt(q, s));
  }
  if (grad_wl || grad_wr) {
    GN_CUDA(cudaMemsetAsync(w.dwcat, 0, sizeof(float) * 2 * co * ci, s));
    GemmTN q{};
    q.A = w.gz; q.lda = 2 * co; q.P = 2 * co; q.B = x; q.ldb = ci; q.Q = ci; q.Nrows = N; q.C = w.dwcat; q.ldc = ci;
    GN_TRY(gemm_tn(q, w.partials, s));
    PackSegHost sg[2];
    int n = 0;
    if (grad_wl) sg[n++] = PackSegHost{grad_wl, w.dwcat, co, ci, ci, ci, 0, 1};
    if (grad_wr) sg[n++] = PackSegHost{grad_wr, w.dwcat + (size_t)co * ci, co, ci, ci, ci, 0, 1};
    GN_TRY(pack_segments(sg, n, s));
  }
  if (grad_bl) GN_TRY(colsum_accum(gm, co, N, co, grad_bl, 1.f, w.colpart, s));
  return GNODE_OK;
}

// ------------------------------------------------------------------------------------------------
// Bipartite SAGEConv forward, the form HeteroConv runs per relation (scripts/gnode.py:92-99, 126-128):
//   out = post( accum + scale * ( [mean_{j in N(i)} x_src[j] | x_dst[i]] @ [wl | wr]^T + bl ) )
// `accum` / `scale` let the caller fold HeteroConv's mean over the relations of one destination type (and the ReLU
// that follows it) into the last relation's epilogue.
// ------------------------------------------------------------------------------------------------
namespace gnode {
namespace {
struct BipWs { float *wcat, *cat, *planes; };
void carve_bip(Arena& a, int64_t n_dst, int ci, int co, BipWs& w) {
  w.wcat = a.take<float>((size_t)co * 2 * ci);
  w.cat = a.take<float>((size_t)n_dst * 2 * ci);
  w.planes = a.take<float>(presplit_floats(co, 2 * ci));
}
}  // namespace
}  // namespace gnode

extern "C" size_t gnode_sage_bipartite_workspace_bytes(int64_t n_dst, int32_t c_in, int32_t c_out) {
  Arena a(nullptr, 0);
  BipWs w;
  carve_bip(a, n_dst, c_in, c_out, w);
  return a.off;
}

extern "C" int gnode_sage_bipartite_fwd(const gnode_graph* g, int64_t n_dst, const float* x_src, const float* x_dst,
                                        int32_t ci, int32_t co, const float* wl, const float* bl, const float* wr,
                                        float scale, const float* accum, int32_t post_relu, float* out,
                                        void* workspace, size_t workspace_bytes, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_ARG(g != nullptr && g->rowptr != nullptr && (g->n_edges == 0 || g->col != nullptr), "gnode_sage_bipartite_fwd: bad graph");
  GN_ARG(n_dst > 0 && n_dst <= g->n_nodes, "gnode_sage_bipartite_fwd: n_dst must be in (0, graph rows]");
  GN_ARG(ci > 0 && co > 0 && x_src && x_dst && wl && wr && out, "gnode_sage_bipartite_fwd: bad argument");
  Arena a(workspace, workspace_bytes);
  BipWs w;
  carve_bip(a, n_dst, ci, co, w);
  GN_ARENA_OK(a, "gnode_sage_bipartite_fwd");
  gnode_graph gd = *g;
  gd.n_nodes = n_dst;   // rows = destination nodes; column ids index x_src
  PackSegHost sg[3] = {{w.wcat, wl, co, ci, ci, 2 * ci, 0, 0},
                       {w.wcat + ci, wr, co, ci, ci, 2 * ci, 0, 0},
                       {w.cat + ci, x_dst, (int)n_dst, ci, ci, 2 * ci, 0, 0}};
  GN_TRY(pack_segments(sg, 3, s));
  GN_TRY(agg_mean_fwd(gd, x_src, ci, w.cat, 2 * ci, ci, nullptr, 0, nullptr, 0, s));
  GemmNT q{};
  q.A = w.cat; q.lda = 2 * ci; q.B = w.wcat; q.ldb = 2 * ci; q.C = out; q.ldc = co; q.M = n_dst; q.N = co; q.K = 2 * ci;
  q.bias = bl; q.scale = scale; q.base = accum; q.ldbase = co; q.post_relu = post_relu ? 1 : 0;
  if (current_engine() != GNODE_ENGINE_SIMT) {
    GN_TRY(presplit_weights(w.wcat, co, 2 * ci, 2 * ci, w.planes, s));
    q.Bsplit = w.planes;
  }
  return gemm_nt(q, s);
}

// ------------------------------------------------------------------------------------------------
// Backward of the bipartite layer: with gs = scale * grad_out * [out > 0] (mask only when `out` is given),
//   grad_wl += gs^T A(x_src)    grad_bl += colsum(gs)    grad_wr += gs^T x_dst
//   grad_x_dst = gs @ wr        grad_x_src = A^T(gs @ wl)                       (both overwritten, may be NULL)
// The Q-network of scripts/gnode.py / scripts/run_gnode.py is trained through these layers (loss.backward()).
// ------------------------------------------------------------------------------------------------
namespace gnode {
namespace {
struct BipBwdWs { float *wcat, *wcatT, *cat, *gm, *gcat, *dwcat, *partials; };
void carve_bip_bwd(Arena& a, int64_t n_dst, int ci, int co, BipBwdWs& w) {
  w.wcat = a.take<float>((size_t)co * 2 * ci);
  w.wcatT = a.take<float>((size_t)co * 2 * ci);
  w.cat = a.take<float>((size_t)n_dst * 2 * ci);
  w.gm = a.take<float>((size_t)n_dst * co);
  w.gcat = a.take<float>((size_t)n_dst * 2 * ci);
  w.dwcat = a.take<float>((size_t)co * 2 * ci);
  w.partials = a.take<float>(gemm_tn_workspace_floats(co, 2 * ci, n_dst));
}
}  // namespace
}  // namespace gnode

extern "C" size_t gnode_sage_bipartite_bwd_workspace_bytes(int64_t n_dst, int32_t c_in, int32_t c_out) {
  Arena a(nullptr, 0);
  BipBwdWs w;
  carve_bip_bwd(a, n_dst, c_in, c_out, w);
  return a.off;
}

extern "C" int gnode_sage_bipartite_bwd(const gnode_graph* g, int64_t n_src, int64_t n_dst, const float* x_src,
                                        const float* x_dst, int32_t ci, int32_t co, const float* wl, const float* wr,
                                        const float* grad_out, const float* out, float scale, float* grad_x_src,
                                        float* grad_x_dst, float* grad_wl, float* grad_bl, float* grad_wr,
                                        void* workspace, size_t workspace_bytes, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_ARG(g != nullptr && g->rowptr != nullptr && g->t_rowptr != nullptr && (g->n_edges == 0 || (g->col && g->t_col)),
         "gnode_sage_bipartite_bwd: bad graph");
  GN_ARG(n_dst > 0 && n_dst <= g->n_nodes && n_src > 0 && n_src <= g->n_nodes, "gnode_sage_bipartite_bwd: n_src / n_dst out of range");
  GN_ARG(ci > 0 && co > 0 && x_src && x_dst && wl && wr && grad_out, "gnode_sage_bipartite_bwd: bad argument");
  Arena a(workspace, workspace_bytes);
  BipBwdWs w;
  carve_bip_bwd(a, n_dst, ci, co, w);
  GN_ARENA_OK(a, "gnode_sage_bipartite_bwd");
  gnode_graph gd = *g;
  gd.n_nodes = n_dst;
  const float* gm = grad_out;
  if (out) {
    GN_TRY(relu_mask(grad_out, out, w.gm, n_dst * (int64_t)co, s));
    gm = w.gm;
  }
  // cat = [A(x_src) | x_dst], wcat = [wl | wr], wcatT = wcat^T
  PackSegHost sg[5] = {{w.wcat, wl, co, ci, ci, 2 * ci, 0, 0},
                       {w.wcat + ci, wr, co, ci, ci, 2 * ci, 0, 0},
                       {w.cat + ci, x_dst, (int)n_dst, ci, ci, 2 * ci, 0, 0},
                       {w.wcatT, wl, co, ci, ci, co, 1, 0},
                       {w.wcatT + (size_t)ci * co, wr, co, ci, ci, co, 1, 0}};
  GN_TRY(pack_segments(sg, 5, s));
  GN_TRY(agg_mean_fwd(gd, x_src, ci, w.cat, 2 * ci, ci, nullptr, 0, nullptr, 0, s));
  if (grad_wl || grad_wr || grad_bl) {
    GN_CUDA(cudaMemsetAsync(w.dwcat, 0, sizeof(float) * co * 2 * ci, s));
    GemmTN q{};
    q.A = gm; q.lda = co; q.P = co; q.B = w.cat; q.ldb = 2 * ci; q.Q = 2 * ci; q.Nrows = n_dst; q.C = w.dwcat; q.ldc = 2 * ci;
    q.scale = scale;
    if (grad_bl) { q.colsumA = grad_bl; q.colsumA_scale = scale; }
    GN_TRY(gemm_tn(q, w.partials, s));
    PackSegHost so[2];
    int n = 0;
    if (grad_wl) so[n++] = PackSegHost{grad_wl, w.dwcat, co, ci, 2 * ci, ci, 0, 1};
    if (grad_wr) so[n++] = PackSegHost{grad_wr, w.dwcat + ci, co, ci, 2 * ci, ci, 0, 1};
    if (n) GN_TRY(pack_segments(so, n, s));
  }
  if (grad_x_src || grad_x_dst) {
    GemmNT q{};   // gcat = scale * gm @ wcat      [n_dst, 2 ci]
    q.A = gm; q.lda = co; q.B = w.wcatT; q.ldb = co; q.C = w.gcat; q.ldc = 2 * ci; q.M = n_dst; q.N = 2 * ci; q.K = co;
    q.scale = scale;
    GN_TRY(gemm_nt(q, s));
    if (grad_x_dst) {
      PackSegHost so[1] = {{grad_x_dst, w.gcat + ci, (int)n_dst, ci, 2 * ci, ci, 0, 0}};
      GN_TRY(pack_segments(so, 1, s));
    }
    if (grad_x_src) {
      gnode_graph gs = *g;
      gs.n_nodes = n_src;      // rows of the transposed CSR = source nodes
      GN_TRY(agg_mean_bwd(gs, w.gcat, 2 * ci, grad_x_src, ci, ci, nullptr, 0, nullptr, 0, s));
    }
  }
  return GNODE_OK;
}

// ------------------------------------------------------------------------------------------------
// Backward of a Linear layer run through the NT contraction:  out = act(x @ w^T + b), act in {identity, relu}.
//   grad_x = gm @ w (overwritten, may be NULL), grad_w += gm^T x, grad_b += colsum(gm), gm = grad_out * [out > 0] if out.
// (embeddings and action heads of HeteroGraphODENetwork, scripts/gnode.py:84-90,104-115)
// ------------------------------------------------------------------------------------------------
extern "C" size_t gnode_linear_bwd_workspace_bytes(int64_t m, int32_t c_in, int32_t c_out) {
  Arena a(nullptr, 0);
  a.take<float>((size_t)m * c_out);
  a.take<float>((size_t)c_in * c_out);
  a.take<float>(gemm_tn_workspace_floats(c_out, c_in, m));
  return a.off;
}

extern "C" int gnode_linear_bwd(const float* x, const float* w, const float* out, const float* grad_out, int64_t m,
                                int32_t ci, int32_t co, float* grad_x, float* grad_w, float* grad_b, void* workspace,
                                size_t workspace_bytes, gnode_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GN_ARG(x && w && grad_out && m > 0 && ci > 0 && co > 0, "gnode_linear_bwd: bad argument");
  Arena a(workspace, workspace_bytes);
  float* gmb = a.take<float>((size_t)m * co);
  float* wT = a.take<float>((size_t)ci * co);
  float* partials = a.take<float>(gemm_tn_workspace_floats(co, ci, m));
  GN_ARENA_OK(a, "gnode_linear_bwd");
  const float* gm = grad_out;
  if (out) {
    GN_TRY(relu_mask(grad_out, out, gmb, m * (int64_t)co, s));
    gm = gmb;
  }
  if (grad_w || grad_b) {
    GemmTN q{};
    q.A = gm; q.lda = co; q.P = co; q.B = x; q.ldb = ci; q.Q = ci; q.Nrows = m; q.C = grad_w; q.ldc = ci;
    if (grad_b) q.colsumA = grad_b;
    if (grad_w) {
      GN_TRY(gemm_tn(q, partials, s));
    } else {
      // bias only: column sums through the same engine into a scratch product
      float* scratch = wT;
      GN_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float) * ci * co, s));
      q.C = scratch;
      GN_TRY(gemm_tn(q, partials, s));
    }
  }
  if (grad_x) {
    PackSegHost sg[1] = {{wT, w, co, ci, ci, co, 1, 0}};      // wT[c, r] = w[r, c]   [ci, co]
    GN_TRY(pack_segments(sg, 1, s));
    GemmNT q{};
    q.A = gm; q.lda = co; q.B = wT; q.ldb = co; q.C = grad_x; q.ldc = ci; q.M = m; q.N = ci; q.K = co;
    GN_TRY(gemm_nt(q, s));
  }
  return GNODE_OK;
}
