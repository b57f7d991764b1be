// C-ABI plumbing: error state, launch counter, and the composite Graphormer-stack entry point.
#include <atomic>
#include <cstdarg>
#include <cstdlib>

#include "common.cuh"

namespace ghn3 {

static thread_local char g_error[1024] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Programmatic dependent launch is a latency optimisation: a kernel's successor is made resident early and waits in
// griddepcontrol.wait, so every kernel chain holds the SM resources of TWO kernels. For ONE chain that hides launch and
// prologue latency (Graphormer stack 1.08 vs 1.49 ms); when several chains run side by side (GHN3.pipeline_depth >= 3)
// the parked CTAs starve the other chains and throughput drops (1.31 vs 1.11 ms per step). Hence a process-wide switch.
static std::atomic<int> g_pdl{getenv("GHN3_NO_PDL") == nullptr ? 1 : 0};
bool pdl_enabled() { return g_pdl.load(std::memory_order_relaxed) != 0; }

// Throughput mode (several predictions in flight): the persistent weight-streaming GEMMs of one prediction's decoders
// take one CTA with ~190 KB of shared memory on EVERY SM, so the Graphormer kernels of the next predictions cannot get
// an SM until a decoder GEMM ends. Capping their grid at ~2/3 of the SMs costs the decoders ~6-16 % and leaves the rest
// of the machine to the latency-bound chains: 1.05 -> 1.01 ms per step end to end (measured at 100 of 148).
static std::atomic<int> g_persistent_cap{0};
int persistent_cta_cap() { return g_persistent_cap.load(std::memory_order_relaxed); }

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      sms = 148;
  }
  return sms;
}

int gemm_impl(const ghn3_gemm_args* a, cudaStream_t stream);
int attention_impl(const ghn3_attention_args* a, cudaStream_t stream);
int layernorm_impl(const ghn3_layernorm_args* a, cudaStream_t stream);

static ghn3_gemm_args linear(const void* x, int64_t rows, int k, const void* w, int n, const float* bias, void* out,
                             int in_dtype, int out_dtype, int act, int accumulate, int x3) {
  ghn3_gemm_args g = {};
  g.a = x; g.a_rows = rows; g.lda = k;
  g.b = w; g.b_rows = n; g.ldb = k;
  g.k = k;
  g.in_dtype = in_dtype;
  g.d = out; g.out_dtype = out_dtype;
  g.bias = bias;
  g.act = act;
  g.accumulate = accumulate;
  g.tf32_x3 = x3;
  g.single.a_row0 = 0; g.single.b_row0 = 0;
  g.single.m = (int32_t)rows; g.single.n = n;
  g.single.d_off = 0; g.single.ldd = n;
  g.single.bias_off = bias ? 0 : -1;
  return g;
}

int graphormer_impl(const ghn3_graphormer_args* a, cudaStream_t stream) {
  GHN3_REQUIRE(a != nullptr && a->layers_host != nullptr, "ghn3_graphormer_stack: null args");
  GHN3_REQUIRE(a->dtype == GHN3_BF16 || a->dtype == GHN3_TF32, "ghn3_graphormer_stack: dtype must be BF16 or TF32");
  GHN3_REQUIRE(a->hid % 16 == 0, "ghn3_graphormer_stack: hid must be a multiple of 16");
  if (a->total_nodes <= 0) return GHN3_OK;
  const int C = a->hid, M = a->total_nodes, dt = a->dtype;
  const int x3 = (a->tf32_x3 != 0 && dt == GHN3_TF32) ? 1 : 0;
  const int act_dt = x3 ? GHN3_F32 : dt;      // storage dtype tag of the activations produced between GEMMs
  int rc;
  const bool fuse_ln = a->ln_counters != nullptr && C <= 1024;
  if (fuse_ln) GHN3_CUDA(cudaMemsetAsync(a->ln_counters, 0, sizeof(int32_t) * ((M + 127) / 128), stream));
  for (int l = 0; l < a->layers; ++l) {
    const ghn3_layer_weights& w = a->layers_host[l];
    ghn3_layernorm_args ln = {};
    ln.rows = M; ln.hid = C; ln.x = a->x; ln.gamma = w.ln1_w; ln.beta = w.ln1_b; ln.out = a->h; ln.out_dtype = act_dt;
    if (!fuse_ln || l == 0) {                  // otherwise produced by the previous layer's FFN2 epilogue
      if ((rc = layernorm_impl(&ln, stream)) != GHN3_OK) return rc;
    }

    ghn3_gemm_args qkv = linear(a->h, M, C, w.w_qkv, 3 * C, nullptr, a->qkv, dt, act_dt, GHN3_ACT_NONE, 0, x3);
    if ((rc = gemm_impl(&qkv, stream)) != GHN3_OK) return rc;

    ghn3_attention_args at = {};
    at.n_graphs = a->n_graphs; at.hid = C; at.heads = a->heads; at.max_nodes = a->max_nodes;
    at.lut_size = a->lut_size; at.node_off = a->node_off; at.mat_off = a->mat_off;
    at.qkv = a->qkv; at.dtype = act_dt; at.pair = a->pair; at.lut = a->lut; at.out = a->h;
    if ((rc = attention_impl(&at, stream)) != GHN3_OK) return rc;

    ghn3_gemm_args proj = linear(a->h, M, C, w.w_out, C, w.b_out, a->x, dt, GHN3_F32, GHN3_ACT_NONE, 1, x3);
    if (fuse_ln) {                             // LN2 of this layer, written into h2 (h is this GEMM's A operand)
      proj.ln_out = a->h2; proj.ln_gamma = w.ln2_w; proj.ln_beta = w.ln2_b;
      proj.ln_counters = a->ln_counters; proj.ln_out_dtype = act_dt;
    }
    if ((rc = gemm_impl(&proj, stream)) != GHN3_OK) return rc;

    ln.gamma = w.ln2_w; ln.beta = w.ln2_b;
    if (!fuse_ln) {
      if ((rc = layernorm_impl(&ln, stream)) != GHN3_OK) return rc;
    }

    ghn3_gemm_args ff1 = linear(fuse_ln ? a->h2 : a->h, M, C, w.w_ff1, 4 * C, w.b_ff1, a->ff, dt, act_dt, GHN3_ACT_GELU,
                                0, x3);
    if ((rc = gemm_impl(&ff1, stream)) != GHN3_OK) return rc;

    ghn3_gemm_args ff2 = linear(a->ff, M, 4 * C, w.w_ff2, C, w.b_ff2, a->x, dt, GHN3_F32, GHN3_ACT_NONE, 1, x3);
    if (fuse_ln && l + 1 < a->layers) {        // LN1 of the next layer
      const ghn3_layer_weights& wn = a->layers_host[l + 1];
      ff2.ln_out = a->h; ff2.ln_gamma = wn.ln1_w; ff2.ln_beta = wn.ln1_b;
      ff2.ln_counters = a->ln_counters; ff2.ln_out_dtype = act_dt;
    }
    if ((rc = gemm_impl(&ff2, stream)) != GHN3_OK) return rc;
  }
  if (a->skip_final_ln) return GHN3_OK;
  {
    // ln_w == NULL (layernorm=False, ghn3/nn.py:262): the same kernel in its identity form -- conversion + row scatter
    ghn3_layernorm_args ln = {};
    ln.rows = M; ln.hid = C; ln.x = a->x; ln.gamma = a->ln_w; ln.beta = a->ln_w ? a->ln_b : nullptr;
    ln.out = a->dec_in; ln.out_dtype = a->dec_dtype; ln.dst_row = a->dst_row; ln.out_f32 = a->emb_f32;
    if ((rc = layernorm_impl(&ln, stream)) != GHN3_OK) return rc;
  }
  return GHN3_OK;
}

}  // namespace ghn3

extern "C" const char* ghn3_last_error(void) { return ghn3::g_error; }
extern "C" int ghn3_abi_version(void) { return GHN3_ABI_VERSION; }
extern "C" int ghn3_set_programmatic_launch(int enabled) {
  const int old = ghn3::g_pdl.exchange(enabled ? 1 : 0);
  return old;
}
extern "C" int ghn3_set_persistent_ctas(int ctas) {
  return ghn3::g_persistent_cap.exchange(ctas > 0 ? ctas : 0);
}
extern "C" int64_t ghn3_launch_count(void) { return ghn3::g_launches.load(std::memory_order_relaxed); }

extern "C" int ghn3_run_sequence(const ghn3_op* ops, int32_t n, ghn3_stream_t stream) {
  if (ops == nullptr && n > 0) {
    ghn3::set_error("ghn3_run_sequence: null op table");
    return GHN3_ERR_BAD_ARG;
  }
  // auxiliary lane (ops[i].lane == 1): one library-owned stream + two events, shared by all sequences of the process
  // (every use is bracketed by FORK / JOIN on the caller's stream, so sequences cannot interleave on it)
  // (one set per device: streams and events belong to the device that was current when they were created)
  constexpr int kMaxDev = 64;
  static cudaStream_t aux_of[kMaxDev] = {};
  static cudaEvent_t ev_fork_of[kMaxDev] = {}, ev_join_of[kMaxDev] = {};
  cudaStream_t aux = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  const ghn3_stream_t main_stream = stream;
  for (int i = 0; i < n; ++i) {
    int rc;
    if (aux == nullptr && (ops[i].op == GHN3_OP_FORK || ops[i].op == GHN3_OP_JOIN || ops[i].lane == 1)) {
      int dev = 0;
      GHN3_CUDA(cudaGetDevice(&dev));
      if (dev < 0 || dev >= kMaxDev) {
        ghn3::set_error("ghn3_run_sequence: device ordinal %d out of range for the auxiliary lane", dev);
        return GHN3_ERR_BAD_ARG;
      }
      if (aux_of[dev] == nullptr) {
        GHN3_CUDA(cudaStreamCreateWithFlags(&aux_of[dev], cudaStreamNonBlocking));
        GHN3_CUDA(cudaEventCreateWithFlags(&ev_fork_of[dev], cudaEventDisableTiming));
        GHN3_CUDA(cudaEventCreateWithFlags(&ev_join_of[dev], cudaEventDisableTiming));
      }
      aux = aux_of[dev]; ev_fork = ev_fork_of[dev]; ev_join = ev_join_of[dev];
    }
    if (ops[i].op == GHN3_OP_FORK) {
      GHN3_CUDA(cudaEventRecord(ev_fork, (cudaStream_t)main_stream));
      GHN3_CUDA(cudaStreamWaitEvent(aux, ev_fork, 0));
      continue;
    }
    if (ops[i].op == GHN3_OP_JOIN) {
      GHN3_CUDA(cudaEventRecord(ev_join, aux));
      GHN3_CUDA(cudaStreamWaitEvent((cudaStream_t)main_stream, ev_join, 0));
      continue;
    }
    stream = ops[i].lane == 1 ? (ghn3_stream_t)aux : main_stream;
    switch (ops[i].op) {
      case GHN3_OP_NODE_FEATURES: rc = ghn3_node_features((const ghn3_node_features_args*)ops[i].args, stream); break;
      case GHN3_OP_GRAPHORMER: rc = ghn3_graphormer_stack((const ghn3_graphormer_args*)ops[i].args, stream); break;
      case GHN3_OP_GEMM: rc = ghn3_gemm((const ghn3_gemm_args*)ops[i].args, stream); break;
      case GHN3_OP_GEMM_SIMT: rc = ghn3_gemm_simt((const ghn3_gemm_simt_args*)ops[i].args, stream); break;
      case GHN3_OP_SCATTER: rc = ghn3_scatter((const ghn3_scatter_args*)ops[i].args, stream); break;
      case GHN3_OP_RELU_TRANSPOSE: rc = ghn3_relu_transpose((const ghn3_relu_transpose_args*)ops[i].args, stream); break;
      case GHN3_OP_GRAPHORMER_TRAIN_FWD: rc = ghn3_graphormer_train_fwd((const ghn3_graphormer_train_args*)ops[i].args, stream); break;
      case GHN3_OP_GRAPHORMER_BWD: rc = ghn3_graphormer_bwd((const ghn3_graphormer_bwd_args*)ops[i].args, stream); break;
      case GHN3_OP_TRANSPOSE: rc = ghn3_transpose((const ghn3_transpose_args*)ops[i].args, stream); break;
      case GHN3_OP_ELEMENTWISE: rc = ghn3_elementwise((const ghn3_elementwise_args*)ops[i].args, stream); break;
      case GHN3_OP_COLSUM: rc = ghn3_colsum((const ghn3_colsum_args*)ops[i].args, stream); break;
      case GHN3_OP_LAYERNORM_BWD: rc = ghn3_layernorm_bwd((const ghn3_layernorm_bwd_args*)ops[i].args, stream); break;
      case GHN3_OP_ATTENTION_BWD: rc = ghn3_attention_bwd((const ghn3_attention_bwd_args*)ops[i].args, stream); break;
      case GHN3_OP_SCATTER_BWD: rc = ghn3_scatter_bwd((const ghn3_scatter_bwd_args*)ops[i].args, stream); break;
      case GHN3_OP_NODE_FEATURES_BWD: rc = ghn3_node_features_bwd((const ghn3_node_features_bwd_args*)ops[i].args, stream); break;
      case GHN3_OP_EDGE_LUT_BWD: rc = ghn3_edge_lut_bwd((const ghn3_edge_lut_bwd_args*)ops[i].args, stream); break;
      case GHN3_OP_FC_BWD: rc = ghn3_fc_bwd((const ghn3_fc_bwd_args*)ops[i].args, stream); break;
      case GHN3_OP_RELU_TRANSPOSE_BWD: rc = ghn3_relu_transpose_bwd((const ghn3_relu_transpose_bwd_args*)ops[i].args, stream); break;
      case GHN3_OP_GRAPHORMER_FUSED: rc = ghn3_graphormer_fused((const ghn3_graphormer_fused_args*)ops[i].args, stream); break;
      case GHN3_OP_LAYERNORM: rc = ghn3_layernorm((const ghn3_layernorm_args*)ops[i].args, stream); break;
      case GHN3_OP_EXPAND_COLS: rc = ghn3_expand_cols((const ghn3_expand_args*)ops[i].args, stream); break;
      case GHN3_OP_MEMSET: {
        const ghn3_memset_args* m = (const ghn3_memset_args*)ops[i].args;
        rc = GHN3_OK;
        if (m->bytes > 0 && cudaMemsetAsync(m->ptr, 0, (size_t)m->bytes, (cudaStream_t)stream) != cudaSuccess) {
          ghn3::set_error("ghn3_run_sequence: cudaMemsetAsync failed at index %d", i);
          rc = GHN3_ERR_CUDA;
        }
        break;
      }
      case GHN3_OP_MEMCPY: {
        const ghn3_memcpy_args* m = (const ghn3_memcpy_args*)ops[i].args;
        rc = GHN3_OK;
        if (m->bytes > 0 && cudaMemcpyAsync(m->dst, m->src, (size_t)m->bytes, cudaMemcpyDeviceToDevice,
                                            (cudaStream_t)stream) != cudaSuccess) {
          ghn3::set_error("ghn3_run_sequence: cudaMemcpyAsync failed at index %d", i);
          rc = GHN3_ERR_CUDA;
        }
        break;
      }
      default:
        ghn3::set_error("ghn3_run_sequence: unknown op %d at index %d", ops[i].op, i);
        return GHN3_ERR_BAD_ARG;
    }
    if (rc != GHN3_OK) return rc;
  }
  return GHN3_OK;
}

struct ghn3_sequence {
  cudaGraphExec_t exec;
  int64_t launches;
};

extern "C" int ghn3_sequence_capture(const ghn3_op* ops, int32_t n, int32_t high_priority, ghn3_sequence** out) {
  if (out == nullptr || (ops == nullptr && n > 0)) {
    ghn3::set_error("ghn3_sequence_capture: null argument");
    return GHN3_ERR_BAD_ARG;
  }
  *out = nullptr;
  // the caller's stream may be the legacy default stream, which cannot be captured: record on streams of our own
  // (one per priority class; kernel nodes inherit the capturing stream's priority)
  constexpr int kMaxDev = 64;
  static cudaStream_t cap_of[kMaxDev][2] = {};
  const int pr = high_priority ? 1 : 0;
  int dev = 0;
  GHN3_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDev) {
    ghn3::set_error("ghn3_sequence_capture: device ordinal %d out of range", dev);
    return GHN3_ERR_BAD_ARG;
  }
  cudaStream_t* cap = cap_of[dev];
  if (cap[pr] == nullptr) {
    int lo = 0, hi = 0;
    GHN3_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    GHN3_CUDA(cudaStreamCreateWithPriority(&cap[pr], cudaStreamNonBlocking, pr ? hi : lo));
  }
  const int64_t before = ghn3::g_launches.load(std::memory_order_relaxed);
  GHN3_CUDA(cudaStreamBeginCapture(cap[pr], cudaStreamCaptureModeRelaxed));
  const int rc = ghn3_run_sequence(ops, n, (ghn3_stream_t)cap[pr]);
  cudaGraph_t graph = nullptr;
  const cudaError_t end = cudaStreamEndCapture(cap[pr], &graph);
  const int64_t captured = ghn3::g_launches.exchange(before, std::memory_order_relaxed) - before;   // nothing ran
  if (rc != GHN3_OK) {
    if (graph != nullptr) cudaGraphDestroy(graph);
    (void)cudaGetLastError();
    return rc;
  }
  if (end != cudaSuccess || graph == nullptr) {
    (void)cudaGetLastError();
    ghn3::set_error("ghn3_sequence_capture: cudaStreamEndCapture failed: %s", cudaGetErrorString(end));
    return GHN3_ERR_CUDA;
  }
  cudaGraphExec_t exec = nullptr;
  const cudaError_t inst = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (inst != cudaSuccess) {
    (void)cudaGetLastError();
    ghn3::set_error("ghn3_sequence_capture: cudaGraphInstantiate failed: %s", cudaGetErrorString(inst));
    return GHN3_ERR_CUDA;
  }
  ghn3_sequence* s = new ghn3_sequence();
  s->exec = exec;
  s->launches = captured;
  *out = s;
  return GHN3_OK;
}

extern "C" int ghn3_sequence_launch(ghn3_sequence* seq, ghn3_stream_t stream) {
  if (seq == nullptr) {
    ghn3::set_error("ghn3_sequence_launch: null sequence");
    return GHN3_ERR_BAD_ARG;
  }
  GHN3_CUDA(cudaGraphLaunch(seq->exec, (cudaStream_t)stream));
  ghn3::count_launch((int)seq->launches);
  return GHN3_OK;
}

extern "C" int ghn3_sequence_destroy(ghn3_sequence* seq) {
  if (seq == nullptr) return GHN3_OK;
  const cudaError_t err = cudaGraphExecDestroy(seq->exec);
  delete seq;
  if (err != cudaSuccess) {
    (void)cudaGetLastError();
    ghn3::set_error("ghn3_sequence_destroy: %s", cudaGetErrorString(err));
    return GHN3_ERR_CUDA;
  }
  return GHN3_OK;
}

extern "C" int ghn3_graphormer_stack(const ghn3_graphormer_args* args, ghn3_stream_t stream) {
  return ghn3::graphormer_impl(args, (cudaStream_t)stream);
}
