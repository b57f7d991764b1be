// Training path of the Graphormer stack: a forward pass that keeps every activation the adjoint needs, and the
// backward pass itself (reference: autograd through ghn3/graphormer.py:208-248 under ghn3/trainer.py:269-345).
// Host-side orchestration only; every kernel lives in dense_kernels.cu / gemm_tcgen05.cu / backward_kernels.cu.
#include "common.cuh"

namespace ghn3 {

int gemm_impl(const ghn3_gemm_args* a, cudaStream_t stream);
int attention_impl(const ghn3_attention_args* a, cudaStream_t stream);
int attention_bwd_impl(const ghn3_attention_bwd_args* a, cudaStream_t stream);
int layernorm_impl(const ghn3_layernorm_args* a, cudaStream_t stream);

namespace {

struct Ctx {
  int C, M, dt, x3, act_dt;
  size_t act_bytes;
  cudaStream_t stream;
};

inline void* at(void* base, int64_t layer, int64_t width, const Ctx& c) {
  return (char*)base + (size_t)layer * c.M * width * c.act_bytes;
}
inline const void* at(const void* base, int64_t layer, int64_t width, const Ctx& c) {
  return (const char*)base + (size_t)layer * c.M * width * c.act_bytes;
}

// D[m][n] (+)= A[m][:k] . B[n][:k]
ghn3_gemm_args mm(const Ctx& c, const void* a, int64_t m, int64_t lda, const void* b, int64_t n, int64_t ldb, int k,
                  const float* bias, void* d, int out_dtype, int accumulate, int b_dynamic) {
  ghn3_gemm_args g = {};
  g.a = a; g.a_rows = m; g.lda = lda;
  g.b = b; g.b_rows = n; g.ldb = ldb;
  g.k = k;
  g.in_dtype = c.dt;
  g.d = d; g.out_dtype = out_dtype;
  g.bias = bias;
  g.act = GHN3_ACT_NONE;
  g.accumulate = accumulate;
  g.tf32_x3 = c.x3;
  g.b_dynamic = b_dynamic;
  g.single.m = (int32_t)m; g.single.n = (int32_t)n;
  g.single.ldd = (int32_t)n;
  g.single.bias_off = bias ? 0 : -1;
  return g;
}

// dst = v^T, optionally with the activation-dtype copy of v, its column sums, and v = src * gelu'(u)
int transpose_to(const Ctx& c, const void* src, int src_dtype, int rows, int cols, void* dst, int64_t ld_dst,
                 void* copy_out = nullptr, float* colsum_out = nullptr, const void* gelu_u = nullptr) {
  ghn3_transpose_args t = {};
  t.src = src; t.src_dtype = src_dtype; t.ld_src = cols;
  t.rows = rows; t.cols = cols;
  t.dst = dst; t.dst_dtype = c.act_dt; t.ld_dst = ld_dst;
  t.mul_gelu_grad = gelu_u; t.mul_dtype = c.act_dt;
  t.copy_out = copy_out; t.copy_dtype = c.act_dt;
  t.colsum_out = colsum_out;
  return ghn3_transpose(&t, (ghn3_stream_t)c.stream);
}

int ew(const Ctx& c, int op, int64_t n, const void* a, int a_dt, const void* b, int b_dt, void* out, int out_dt) {
  ghn3_elementwise_args e = {};
  e.op = op; e.n = n; e.a = a; e.a_dtype = a_dt; e.b = b; e.b_dtype = b_dt; e.out = out; e.out_dtype = out_dt;
  return ghn3_elementwise(&e, (ghn3_stream_t)c.stream);
}

int colsum_to(const Ctx& c, const void* src, int src_dtype, int rows, int cols, float* dst) {
  ghn3_colsum_args s = {};
  s.src = src; s.src_dtype = src_dtype; s.ld = cols; s.rows = rows; s.cols = cols; s.dst = dst;
  return ghn3_colsum(&s, (ghn3_stream_t)c.stream);
}

#define GHN3_TRY(expr)                 \
  do {                                 \
    int rc__ = (expr);                 \
    if (rc__ != GHN3_OK) return rc__;  \
  } while (0)

int make_ctx(const ghn3_graphormer_args& f, cudaStream_t stream, Ctx* c, const char* who) {
  GHN3_REQUIRE(f.layers_host != nullptr, "%s: null layer table", who);
  GHN3_REQUIRE(f.dtype == GHN3_BF16 || f.dtype == GHN3_TF32, "%s: dtype must be BF16 or TF32", who);
  GHN3_REQUIRE(f.hid % 16 == 0, "%s: hid must be a multiple of 16", who);
  c->C = f.hid; c->M = f.total_nodes; c->dt = f.dtype;
  c->x3 = (f.tf32_x3 != 0 && f.dtype == GHN3_TF32) ? 1 : 0;
  c->act_dt = c->x3 ? GHN3_F32 : f.dtype;
  c->act_bytes = f.dtype == GHN3_BF16 ? 2 : 4;
  c->stream = stream;
  return GHN3_OK;
}

}  // namespace

int graphormer_train_fwd_impl(const ghn3_graphormer_train_args* t, cudaStream_t stream) {
  GHN3_REQUIRE(t != nullptr, "ghn3_graphormer_train_fwd: null args");
  const ghn3_graphormer_args& f = t->fwd;
  Ctx c;
  GHN3_TRY(make_ctx(f, stream, &c, "ghn3_graphormer_train_fwd"));
  GHN3_REQUIRE(t->xs && t->xm && t->h1 && t->qkv && t->ao && t->h2 && t->u && t->g,
               "ghn3_graphormer_train_fwd: every saved-activation buffer is required");
  if (c.M <= 0) return GHN3_OK;
  const int C = c.C, M = c.M;
  const size_t xbytes = sizeof(float) * (size_t)M * C;
  for (int l = 0; l < f.layers; ++l) {
    const ghn3_layer_weights& w = f.layers_host[l];
    float* x = t->xs + (size_t)l * M * C;
    float* xm = t->xm + (size_t)l * M * C;
    float* xn = t->xs + (size_t)(l + 1) * M * C;
    void* h1 = at(t->h1, l, C, c);
    void* qkv = at(t->qkv, l, 3 * C, c);
    void* ao = at(t->ao, l, C, c);
    void* h2 = at(t->h2, l, C, c);
    void* u = at(t->u, l, 4 * C, c);
    void* g = at(t->g, l, 4 * C, c);

    ghn3_layernorm_args ln = {};
    ln.rows = M; ln.hid = C; ln.x = x; ln.gamma = w.ln1_w; ln.beta = w.ln1_b; ln.out = h1; ln.out_dtype = c.act_dt;
    GHN3_TRY(layernorm_impl(&ln, stream));
    ghn3_gemm_args g1 = mm(c, h1, M, C, w.w_qkv, 3 * C, C, C, nullptr, qkv, c.act_dt, 0, 0);
    GHN3_TRY(gemm_impl(&g1, stream));

    ghn3_attention_args at_ = {};
    at_.n_graphs = f.n_graphs; at_.hid = C; at_.heads = f.heads; at_.max_nodes = f.max_nodes;
    at_.lut_size = f.lut_size; at_.node_off = f.node_off; at_.mat_off = f.mat_off;
    at_.qkv = qkv; at_.dtype = c.act_dt; at_.pair = f.pair; at_.lut = f.lut; at_.out = ao;
    at_.total_nodes = M;
    at_.lse2 = (t->lse2 != nullptr && c.act_dt == GHN3_BF16) ? t->lse2 + (size_t)l * f.heads * M : nullptr;
    GHN3_TRY(attention_impl(&at_, stream));

    GHN3_CUDA(cudaMemcpyAsync(xm, x, xbytes, cudaMemcpyDeviceToDevice, stream));
    ghn3_gemm_args g2 = mm(c, ao, M, C, w.w_out, C, C, C, w.b_out, xm, GHN3_F32, 1, 0);
    GHN3_TRY(gemm_impl(&g2, stream));

    ln.x = xm; ln.gamma = w.ln2_w; ln.beta = w.ln2_b; ln.out = h2;
    GHN3_TRY(layernorm_impl(&ln, stream));
    ghn3_gemm_args g3 = mm(c, h2, M, C, w.w_ff1, 4 * C, C, C, w.b_ff1, u, c.act_dt, 0, 0);
    GHN3_TRY(gemm_impl(&g3, stream));
    GHN3_TRY(ew(c, GHN3_EW_GELU, (int64_t)M * 4 * C, u, c.act_dt, nullptr, 0, g, c.act_dt));

    GHN3_CUDA(cudaMemcpyAsync(xn, xm, xbytes, cudaMemcpyDeviceToDevice, stream));
    ghn3_gemm_args g4 = mm(c, g, M, 4 * C, w.w_ff2, C, 4 * C, 4 * C, w.b_ff2, xn, GHN3_F32, 1, 0);
    GHN3_TRY(gemm_impl(&g4, stream));
  }
  ghn3_layernorm_args ln = {};
  ln.rows = M; ln.hid = C; ln.x = t->xs + (size_t)f.layers * M * C; ln.gamma = f.ln_w;
  ln.beta = f.ln_w ? f.ln_b : nullptr;                     // ln_w == NULL: identity form (layernorm=False)
  ln.out = f.dec_in; ln.out_dtype = f.dec_dtype; ln.dst_row = f.dst_row; ln.out_f32 = f.emb_f32;
  GHN3_TRY(layernorm_impl(&ln, stream));
  return GHN3_OK;
}

int graphormer_bwd_impl(const ghn3_graphormer_bwd_args* b, cudaStream_t stream) {
  GHN3_REQUIRE(b != nullptr && b->saved != nullptr, "ghn3_graphormer_bwd: null args");
  const ghn3_graphormer_train_args* t = b->saved;
  const ghn3_graphormer_args& f = t->fwd;
  Ctx c;
  GHN3_TRY(make_ctx(f, stream, &c, "ghn3_graphormer_bwd"));
  GHN3_REQUIRE(b->layers_t_host && b->grads_host && (f.ln_w == nullptr || (b->d_ln_w && b->d_ln_b)) && b->d_dec_in && b->dx && b->dxa &&
                   b->dh && b->dhf && b->dqkv && b->dff && b->ta && b->tb && b->lse && b->delta,
               "ghn3_graphormer_bwd: null pointer");
  if (c.M <= 0) return GHN3_OK;
  const int C = c.C, M = c.M, adt = c.act_dt;
  const int64_t mp = b->m_pad;
  GHN3_REQUIRE(mp >= M && mp % 8 == 0, "ghn3_graphormer_bwd: m_pad must be a multiple of 8 and >= total_nodes");
  const ghn3_stream_t s_ = (ghn3_stream_t)stream;

  // tensor-core attention backward: needs the forward's softmax statistics and the dS accumulation scratch
  const bool mma_attn = adt == GHN3_BF16 && t->lse2 != nullptr && b->ds_total != nullptr;
  if (mma_attn) GHN3_CUDA(cudaMemsetAsync(b->ds_total, 0, (size_t)b->ds_total_bytes, stream));
  // final LayerNorm: dx = LN'(xs[L]) . gather(d_dec_in)
  GHN3_CUDA(cudaMemsetAsync(b->dx, 0, sizeof(float) * (size_t)M * C, stream));
  {
    ghn3_layernorm_bwd_args lb = {};
    lb.rows = M; lb.hid = C; lb.x = t->xs + (size_t)f.layers * M * C; lb.gamma = f.ln_w;
    lb.dy = b->d_dec_in; lb.dy_dtype = b->d_dec_dtype; lb.dy_row = f.dst_row;
    lb.dx = b->dx; lb.accumulate = 0; lb.dgamma = b->d_ln_w; lb.dbeta = b->d_ln_b;
    GHN3_TRY(ghn3_layernorm_bwd(&lb, s_));
  }
  for (int l = f.layers - 1; l >= 0; --l) {
    const ghn3_layer_weights& w = f.layers_host[l];
    const ghn3_layer_weights_t& wt = b->layers_t_host[l];
    const ghn3_layer_grads& gr = b->grads_host[l];
    const float* x = t->xs + (size_t)l * M * C;
    const float* xm = t->xm + (size_t)l * M * C;
    const void* h1 = at((const void*)t->h1, l, C, c);
    const void* qkv = at((const void*)t->qkv, l, 3 * C, c);
    const void* ao = at((const void*)t->ao, l, C, c);
    const void* h2 = at((const void*)t->h2, l, C, c);
    const void* u = at((const void*)t->u, l, 4 * C, c);
    const void* g = at((const void*)t->g, l, 4 * C, c);

    // ---- FFN2: xn = xm + g W2^T + b2 ----
    GHN3_TRY(transpose_to(c, b->dx, GHN3_F32, M, C, b->ta, mp, b->dxa, gr.b_ff2));     // dx^T, act copy, bias grad
    GHN3_TRY(transpose_to(c, g, adt, M, 4 * C, b->tb, mp));
    {
      ghn3_gemm_args wg = mm(c, b->ta, C, mp, b->tb, 4 * C, mp, M, nullptr, gr.w_ff2, GHN3_F32, 1, 1);
      GHN3_TRY(gemm_impl(&wg, stream));
      ghn3_gemm_args dg = mm(c, b->dxa, M, C, wt.w_ff2_t, 4 * C, C, C, nullptr, b->dff, adt, 0, 0);
      GHN3_TRY(gemm_impl(&dg, stream));
    }
    // ---- GELU, FFN1: u = h2 W1^T + b1 ----
    // du = dg * gelu'(u) written back in place, du^T and the bias gradient in the same pass
    GHN3_TRY(transpose_to(c, b->dff, adt, M, 4 * C, b->ta, mp, b->dff, gr.b_ff1, u));
    GHN3_TRY(transpose_to(c, h2, adt, M, C, b->tb, mp));
    {
      ghn3_gemm_args wg = mm(c, b->ta, 4 * C, mp, b->tb, C, mp, M, nullptr, gr.w_ff1, GHN3_F32, 1, 1);
      GHN3_TRY(gemm_impl(&wg, stream));
      ghn3_gemm_args dg = mm(c, b->dff, M, 4 * C, wt.w_ff1_t, C, 4 * C, 4 * C, nullptr, b->dhf, GHN3_F32, 0, 0);
      GHN3_TRY(gemm_impl(&dg, stream));
    }
    {
      ghn3_layernorm_bwd_args lb = {};
      lb.rows = M; lb.hid = C; lb.x = xm; lb.gamma = w.ln2_w; lb.dy = b->dhf; lb.dy_dtype = GHN3_F32;
      lb.dx = b->dx; lb.accumulate = 1; lb.dgamma = gr.ln2_w; lb.dbeta = gr.ln2_b;
      GHN3_TRY(ghn3_layernorm_bwd(&lb, s_));
    }
    // ---- attention output projection: xm = x + ao Wo^T + bo ----
    GHN3_TRY(transpose_to(c, b->dx, GHN3_F32, M, C, b->ta, mp, b->dxa, gr.b_out));
    GHN3_TRY(transpose_to(c, ao, adt, M, C, b->tb, mp));
    {
      ghn3_gemm_args wg = mm(c, b->ta, C, mp, b->tb, C, mp, M, nullptr, gr.w_out, GHN3_F32, 1, 1);
      GHN3_TRY(gemm_impl(&wg, stream));
      ghn3_gemm_args dg = mm(c, b->dxa, M, C, wt.w_out_t, C, C, C, nullptr, b->dh, adt, 0, 0);
      GHN3_TRY(gemm_impl(&dg, stream));
    }
    // ---- attention ----
    {
      ghn3_attention_bwd_args ab = {};
      ab.n_graphs = f.n_graphs; ab.hid = C; ab.heads = f.heads; ab.max_nodes = f.max_nodes; ab.total_nodes = M;
      ab.lut_size = f.lut_size; ab.node_off = f.node_off; ab.mat_off = f.mat_off;
      ab.qkv = qkv; ab.out = ao; ab.d_out = b->dh; ab.dtype = adt == GHN3_BF16 ? GHN3_BF16 : GHN3_F32;
      ab.pair = f.pair; ab.lut = f.lut; ab.d_qkv = b->dqkv; ab.d_lut = b->d_lut; ab.lse = b->lse; ab.delta = b->delta;
      if (mma_attn) {
        ab.fwd_lse2 = t->lse2 + (size_t)l * f.heads * M;
        ab.ds_total = b->d_lut != nullptr ? b->ds_total : nullptr;
      }
      GHN3_TRY(attention_bwd_impl(&ab, stream));
    }
    // ---- QKV projection: qkv = h1 Wqkv^T ----
    GHN3_TRY(transpose_to(c, b->dqkv, adt, M, 3 * C, b->ta, mp));
    GHN3_TRY(transpose_to(c, h1, adt, M, C, b->tb, mp));
    {
      ghn3_gemm_args wg = mm(c, b->ta, 3 * C, mp, b->tb, C, mp, M, nullptr, gr.w_qkv, GHN3_F32, 1, 1);
      GHN3_TRY(gemm_impl(&wg, stream));
      ghn3_gemm_args dg = mm(c, b->dqkv, M, 3 * C, wt.w_qkv_t, C, 3 * C, 3 * C, nullptr, b->dhf, GHN3_F32, 0, 0);
      GHN3_TRY(gemm_impl(&dg, stream));
    }
    {
      ghn3_layernorm_bwd_args lb = {};
      lb.rows = M; lb.hid = C; lb.x = x; lb.gamma = w.ln1_w; lb.dy = b->dhf; lb.dy_dtype = GHN3_F32;
      lb.dx = b->dx; lb.accumulate = 1; lb.dgamma = gr.ln1_w; lb.dbeta = gr.ln1_b;
      GHN3_TRY(ghn3_layernorm_bwd(&lb, s_));
    }
  }
  if (mma_attn && b->d_lut != nullptr) {       // edge-bias gradient of all layers at once
    ghn3_lut_bin_args lb = {};
    lb.n_graphs = f.n_graphs; lb.heads = f.heads; lb.max_nodes = f.max_nodes; lb.lut_size = f.lut_size;
    lb.node_off = f.node_off; lb.mat_off = f.mat_off; lb.pair = f.pair; lb.ds_total = b->ds_total; lb.d_lut = b->d_lut;
    GHN3_TRY(ghn3_lut_bin(&lb, s_));
  }
  return GHN3_OK;
}

}  // namespace ghn3

extern "C" int ghn3_graphormer_train_fwd(const ghn3_graphormer_train_args* args, ghn3_stream_t stream) {
  return ghn3::graphormer_train_fwd_impl(args, (cudaStream_t)stream);
}

extern "C" int ghn3_graphormer_bwd(const ghn3_graphormer_bwd_args* args, ghn3_stream_t stream) {
  return ghn3::graphormer_bwd_impl(args, (cudaStream_t)stream);
}
