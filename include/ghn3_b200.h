/*
 * ghn3_b200 -- C ABI of the B200-native GHN-3 parameter-prediction hot path.
 *
 * The reference (SamsungSAILMontreal/ghn3) is pure Python/PyTorch and has no FFI layer: its boundary is the Python
 * API (`GHN3.forward`, ghn3/nn.py:186) and the checkpoint layout. This header is the boundary UNDER that API: each
 * entry point replaces one group of ATen calls of the reference (cited per function, paths relative to the
 * reference root). The Python host (ghn3_b200/nn.py) binds these symbols with ctypes; INTEGRATION.md shows the
 * stub a maintainer of the reference would add.
 *
 * Conventions:
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in `_host`;
 *  - the caller (PyTorch) owns all memory; the library never allocates device memory, never frees, never retains
 *    pointers after a call returns; target parameters are written in place;
 *  - all work is enqueued on the caller's stream; no internal synchronisation; re-entrant;
 *  - return 0 on success or a negative ghn3_status; ghn3_last_error() returns a per-thread message;
 *  - there is no CPU fallback: without a CUDA device every compute entry point fails with GHN3_ERR_CUDA.
 */
#ifndef GHN3_B200_H_
#define GHN3_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* ghn3_stream_t; /* a cudaStream_t */

enum ghn3_status {
  GHN3_OK = 0,
  GHN3_ERR_BAD_ARG = -1,
  GHN3_ERR_UNSUPPORTED = -2,
  GHN3_ERR_CUDA = -3
};

enum ghn3_dtype {
  GHN3_BF16 = 0,      /* bfloat16 storage, kind::f16 tensor-core path */
  GHN3_TF32 = 1,      /* fp32 storage with values rounded to tf32 (cvt.rna), kind::tf32 tensor-core path */
  GHN3_F32 = 2        /* plain fp32 (outputs only) */
};

enum ghn3_act { GHN3_ACT_NONE = 0, GHN3_ACT_RELU = 1, GHN3_ACT_GELU = 2 };

const char* ghn3_last_error(void);
#define GHN3_ABI_VERSION 3   /* bumped when an argument struct changes layout */
int ghn3_abi_version(void);  /* == GHN3_ABI_VERSION of the header the library was built from */
/* Process-wide switch of programmatic dependent launch (default on; GHN3_NO_PDL=1 starts with it off): with it every
 * kernel of a chain is made resident while its predecessor still runs -- lowest latency for ONE chain, but the parked
 * CTAs hold SM resources, which costs throughput when several independent chains run side by side. Returns the
 * previous setting. */
int ghn3_set_programmatic_launch(int enabled);
/* Process-wide cap on the grid of the persistent (weight-streaming) GEMM launches: `ctas` > 0 leaves the other SMs to
 * kernels of other streams (throughput mode, several predictions in flight); 0 = one CTA per SM (default; best for one
 * prediction at a time). GHN3_PERSISTENT_CTAS in the environment overrides it. Returns the previous setting. */
int ghn3_set_persistent_ctas(int ctas);
/* Number of kernels this library has launched since load (all streams); backs bench.py's "gpu_launches". */
int64_t ghn3_launch_count(void);

/* ---------------------------------------------------------------------------------------------------------------
 * (2) Shortest-path "virtual edges" and structural indices  -- replaces networkx all_pairs_shortest_path_length in
 * ghn3/graph.py:755-798 and the index arithmetic of ghn3/graphormer.py:229-237. Integer, bit-exact.
 *
 * A batch holds n_graphs graphs packed back to back. Graph g has n_g = node_off[g+1]-node_off[g] nodes, its 1-hop
 * edges are edge_src/dst[edge_off[g] .. edge_off[g+1]) with LOCAL node ids, and its N x N matrices live at byte /
 * element offset mat_off[g] with leading dimension ld_g = round_up(n_g, 16).
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t n_graphs;
  int32_t cutoff;            /* ve_cutoff, 1..254 (reference: 50) */
  const int32_t* node_off;   /* [n_graphs+1] */
  const int32_t* edge_off;   /* [n_graphs+1] */
  const int64_t* mat_off;    /* [n_graphs+1], multiples of 16 */
  const int32_t* edge_src;   /* [total_edges] */
  const int32_t* edge_dst;   /* [total_edges] */
  int32_t max_nodes;         /* max n_g over the batch (host copy, sizes the launch) */
  int32_t total_nodes;
  int32_t total_edges;
  int64_t bits_total;        /* uint32 words in adj_bits = bits_off[n_graphs] */
  int64_t mat_total;         /* bytes in spd = mat_off[n_graphs] */
  uint32_t* adj_bits;        /* workspace: sum_g n_g * words_g uint32, words_g = ceil(n_g/32); offset = bits_off[g] */
  const int64_t* bits_off;   /* [n_graphs+1] in uint32 units */
  uint8_t* spd;              /* out: SPD matrices, uint8 */
} ghn3_spd_args;

/* edges -> uint8 shortest path distances (0 = none / self / beyond cutoff). */
int ghn3_spd_bfs(const ghn3_spd_args* args, ghn3_stream_t stream);

typedef struct {
  int32_t n_graphs;
  int32_t vmax;              /* largest SPD value that may occur (= cutoff) */
  const int32_t* node_off;
  const int64_t* mat_off;
  int32_t max_nodes;
  int32_t total_nodes;
  const uint8_t* spd;        /* in  */
  uint16_t* pair;            /* out: pair[i][j] = spd[i][j] * (vmax+1) + spd[j][i]  (same offsets / ld, in elements) */
  int32_t* deg_in;           /* out [total_nodes]: min(#{i: spd[i][j]==1}, 100)     graphormer.py:229-230 */
  int32_t* deg_out;          /* out [total_nodes]: min(#{j: spd[i][j]==1}, 100)     graphormer.py:231 */
  int32_t* dist0;            /* out [total_nodes]: min(spd[0][j], 1000)             graphormer.py:232 */
} ghn3_derive_args;

int ghn3_graph_derive(const ghn3_derive_args* args, ghn3_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * (1) Node features -- replaces embed + ppuda ShapeEncoder.forward (ghn3/nn.py:248-249) and the three structural
 * embedding adds of ghn3/graphormer.py:230-232. fp32, same order of additions as the reference (bit-exact).
 *   x[n] = (((E_op[op] + cat(E_ch[s0], E_ch[s1], E_sp[s2], E_sp[s3])) + E_in[deg_in]) + E_out[deg_out]) + E_dist[dist0]
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t total_nodes;
  int32_t hid;                 /* C, multiple of 16 */
  const int32_t* op;           /* [total_nodes] primitive id */
  const int32_t* shape_idx;    /* [total_nodes][4] rows of embed_channel (x2) and embed_spatial (x2) */
  const int32_t* deg_in;
  const int32_t* deg_out;
  const int32_t* dist0;
  const float* embed_op;       /* [15][C] */
  const float* embed_ch;       /* [n_ch+1][C/4] */
  const float* embed_sp;       /* [n_sp+1][C/4] */
  const float* cent_in;        /* [101][C] */
  const float* cent_out;       /* [101][C] */
  const float* dist_embed;     /* [1001][C] */
  float* x;                    /* out [total_nodes][C] */
} ghn3_node_features_args;

int ghn3_node_features(const ghn3_node_features_args* args, ghn3_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * (3b) Edge-bias look-up table -- replaces EdgeEmbedding + proj_e over the materialised (B,N,N,2C) tensor
 * (ghn3/graphormer.py:114-117) by evaluating the same MLP on the (vmax+1)^2 grid of (A_ij, A_ji) values.
 *   lut[h][a*(vmax+1)+b] = W2[h] . relu(W1 . [E[a+2]; E[b+2]] + b1) + b2[h]
 * workspace: 2*(vmax+1)*C floats.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t hid, heads, vmax;
  const float* edge_embed;   /* [257][C] */
  const float* w1;           /* [C][2C] */
  const float* b1;           /* [C] */
  const float* w2;           /* [H][C] */
  const float* b2;           /* [H] */
  float* workspace;
  float* lut;                /* out [H][(vmax+1)^2] */
} ghn3_edge_lut_args;

int ghn3_edge_lut(const ghn3_edge_lut_args* args, ghn3_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * LayerNorm over rows (eps 1e-5) -- ghn3/graphormer.py:239,241 and ghn3/nn.py:262-263. fp32 math; the output is
 * written in the GEMM input dtype. Optional row scatter: out row = dst_row[r] (skip if < 0); optional fp32 copy.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t rows, hid;
  const float* x;
  const float* gamma;        /* NULL (with beta NULL): identity -- no normalisation, only the conversion / row scatter */
  const float* beta;         /*   (GHNs built with layernorm=False, ghn3/nn.py:262) */
  void* out;                 /* [*, C] in out_dtype (GHN3_BF16 / GHN3_TF32 / GHN3_F32) */
  int32_t out_dtype;
  const int32_t* dst_row;    /* optional [rows] */
  float* out_f32;            /* optional [rows][C] un-permuted fp32 copy (return_embeddings) */
} ghn3_layernorm_args;

int ghn3_layernorm(const ghn3_layernorm_args* args, ghn3_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * (3a)/(4a) Tensor-core GEMM  D = epilogue(A . B^T + bias)  -- replaces every nn.Linear of the Graphormer stack
 * (ghn3/graphormer.py:38-44,121,141) and of the decoders (ghn3/nn.py:738,748,758,289-294).
 * A [a_rows][K] and B [b_rows][K] are K-major (row-major with leading dims lda / ldb), both in `in_dtype`.
 * tcgen05.mma (kind::f16 or kind::tf32), accumulators in TMEM, operands staged by TMA with 128B swizzle.
 * A launch runs either ONE problem (`problems == NULL`, described by `single`) or a GROUP of problems that share
 * A, B, K and the epilogue: `tiles[t] = {problem, m_tile, n_tile, 0}` with 128 x block_n tiles.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t a_row0;     /* first row of A used by the problem */
  int32_t b_row0;     /* first row of B (output feature) used by the problem */
  int32_t m, n;       /* extent */
  int64_t d_off;      /* element offset of D[0][0] in the output buffer */
  int32_t ldd;        /* leading dimension of D, elements */
  int32_t bias_off;   /* element offset into bias for column 0, or -1 for no bias */
} ghn3_gemm_problem;

typedef struct {
  const void* a; int64_t a_rows; int64_t lda;
  const void* b; int64_t b_rows; int64_t ldb;
  int32_t k;
  int32_t in_dtype;          /* GHN3_BF16 | GHN3_TF32 */
  void* d;
  int32_t out_dtype;         /* GHN3_BF16 | GHN3_TF32 | GHN3_F32 */
  const float* bias;
  int32_t act;               /* ghn3_act, applied after the bias */
  int32_t accumulate;        /* 1: D (fp32) += result   (residual update in place, graphormer.py:240-241) */
  ghn3_gemm_problem single;
  const ghn3_gemm_problem* problems;   /* device, or NULL */
  const int32_t* tiles;                /* device int32[n_tiles][4], or NULL */
  int32_t n_tiles;
  int32_t block_n;           /* 0 = auto; 64/128/256 */
  int32_t k_splits;          /* single-problem launches: 0 = auto, >1 splits K over blockIdx.z (needs accumulate) */
  int32_t tf32_x3;           /* in_dtype TF32 only: 3-term error-compensated tf32 (hi/lo split in shared memory) */
  /* Grouped launches only. b_group = g > 0 views B as [b_rows / b_group_stride][b_group_stride][K] and makes the
   * problem's columns the COMPACT set {outer * g + inner : inner < g}: column c reads B row
   * (b_row0 + c / g) * b_group_stride + c % g (b_row0 counts OUTER rows; bias is indexed the same way).
   * An N tile holds floor(block_n / g) * g columns. Used for the decoder's o' x i' sub-blocks with i' < max_shape[1]
   * (reference ghn3/nn.py:750,760: x[:, :o, :i]). */
  int32_t b_group;
  int32_t b_group_stride;
  int32_t bias_rows;         /* 1: bias is indexed by the output ROW (bias[bias_off + m]) instead of the column */
  int32_t b_dynamic;         /* 1: B is written by an earlier kernel of the same step. By default B is assumed to hold
                                weights and its first tiles are fetched BEFORE the programmatic-dependent-launch wait. */
  const int32_t* rowmap;     /* optional device table (static per plan): row m of a problem is stored to output row
                                rowmap[problem.d_off + m] instead of d_off / ldd addressing. Lets the decoder's fc
                                stage run one problem per grid POSITION over all nodes whose crop window contains it
                                (reference ghn3/nn.py:738-745) while writing the (node, position)-major h0 layout. */
  int32_t swap_ab;           /* grouped launches whose problems have few rows (m <~ 64): the WEIGHT rows fill the 128
                                UMMA-M lanes and the activation rows become the N dimension (64 per tile), so all four
                                epilogue warps work and the output is stored column-coalesced without staging.
                                Tiles are then {problem, m / 64, n / 128 (or the b_group tile), 0}. */
  /* Fused LayerNorm (eps 1e-5) for single-problem residual GEMMs whose output rows are complete rows of the
   * residual stream (accumulate = 1, ldd == n): the CTA that finishes the last tile / K-split of a 128-row block
   * writes LN(D rows) * gamma + beta into ln_out [m][n] (ln_out_dtype). ln_counters: device int32[ceil(m/128)],
   * zero on entry, zero again on exit. Replaces the separate LayerNorm launches of ghn3/graphormer.py:239,241. */
  void* ln_out;
  const float* ln_gamma;
  const float* ln_beta;
  int32_t* ln_counters;
  int32_t ln_out_dtype;
  /* Sparse-K mode for single-problem launches whose operands are known to be zero outside a few K blocks (the
   * column-expanded operands of the decoder conv.2 backward): the M tile mt (128 rows) only visits the K blocks
   * kb_list[kb_off[mt] .. kb_off[mt+1]) (block = 64 bf16 / 32 fp32 elements of K). An M tile with an empty list is
   * skipped and leaves D untouched (the caller pre-zeroes D). Device arrays; NULL = dense. */
  const int32_t* kb_list;
  const int32_t* kb_off;
  /* single-problem launches: 1 = run on the persistent kernel (one CTA per SM walks tiles enumerated in the kernel,
   * epilogue of tile j under the main loop of tile j + 1). Measured equal to two one-tile CTAs per SM on the K = C GEMMs
   * of an 18 666-row batch (the epilogue warps are the limit either way), so 0 (off) is the default. */
  int32_t persistent_single;
} ghn3_gemm_args;

int ghn3_gemm(const ghn3_gemm_args* args, ghn3_stream_t stream);
/* Profiling aid: subsequent ghn3_gemm launches write per-CTA phase timestamps (globaltimer) into a device buffer of
 * int64[4096][8]; pass NULL to switch it off. */
int ghn3_debug_gemm_trace(void* device_buffer);

/* Small strided fp32 GEMM on CUDA cores for shapes that are too small or too oddly laid out for TMA:
 *   D[m][n] = act(bias[n] + sum_k f(A[m*sam + k*sak]) * B[n*sbn + k*sbk]),  f = relu if relu_a else identity.
 * Used for the classification heads (ghn3/nn.py:757-758, 294). */
typedef struct {
  const float* a; int64_t sam, sak;
  const float* b; int64_t sbn, sbk;
  const float* bias;
  float* d; int64_t sdm, sdn;
  int32_t m, n, k;
  int32_t relu_a;
  int32_t act;
  int32_t batch; int64_t a_bs, d_bs;   /* optional batch over A and D (B and bias shared) */
} ghn3_gemm_simt_args;

int ghn3_gemm_simt(const ghn3_gemm_simt_args* args, ghn3_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * (3b) Fused multi-head attention with the SPD/edge bias -- replaces ghn3/graphormer.py:121-140
 * (QK^T * d^-1/2 + bias, softmax, PV) without materialising (B,H,N,N) or (B,N,N,H).
 * Graphs are packed (no padding => no mask); logit(i,j,h) += lut[h][pair[i][j]].
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t n_graphs, hid, heads, max_nodes;
  int32_t lut_size;          /* (vmax+1)^2 */
  const int32_t* node_off;
  const int64_t* mat_off;
  const void* qkv;           /* [total_nodes][3C] in dtype: q | k | v, head-major inside each */
  int32_t dtype;             /* GHN3_BF16 | GHN3_TF32 (fp32 storage, output rounded to tf32) | GHN3_F32 */
  const uint16_t* pair;
  const float* lut;          /* [H][lut_size] */
  void* out;                 /* [total_nodes][C] in dtype */
  /* optional (GHN3_BF16 only): log2-domain softmax statistics m + log2(l) of every (head, node) row,
   * [H][total_nodes] fp32, kept for ghn3_attention_bwd's tensor-core path */
  float* lse2;
  int32_t total_nodes;
} ghn3_attention_args;

int ghn3_attention(const ghn3_attention_args* args, ghn3_stream_t stream);
/* bf16 batches whose largest graph has at least `min_nodes` nodes run on the tcgen05 kernel (S and P.V accumulated in
 * TMEM, csrc/attention_tcgen05.cu), smaller ones on the mma.sync kernel. Default 2048 (GHN3_ATTN_TC_MIN); returns the
 * previous threshold. */
int ghn3_set_attention_tc_min(int min_nodes);

/* ---------------------------------------------------------------------------------------------------------------
 * The whole Graphormer stack on packed node features (ghn3/nn.py:258-263, graphormer.py:208-248): per layer
 * LN1 -> QKV GEMM -> attention -> out-proj GEMM (+residual) -> LN2 -> FFN1 GEMM (+GELU) -> FFN2 GEMM (+residual),
 * then the final LayerNorm with row scatter into the decoder input buffers. One call enqueues 7*L+1 kernels.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct {
  const float* ln1_w; const float* ln1_b;
  const void* w_qkv;                       /* [3C][C] */
  const void* w_out; const float* b_out;   /* [C][C] */
  const float* ln2_w; const float* ln2_b;
  const void* w_ff1; const float* b_ff1;   /* [4C][C] */
  const void* w_ff2; const float* b_ff2;   /* [C][4C] */
} ghn3_layer_weights;

typedef struct {
  int32_t hid, heads, layers, dtype;
  const ghn3_layer_weights* layers_host;   /* HOST array [layers] of device pointers */
  const float* ln_w; const float* ln_b;    /* final LayerNorm, may be NULL (layernorm=False) */
  /* batch */
  int32_t n_graphs, total_nodes, max_nodes, lut_size;
  const int32_t* node_off;
  const int64_t* mat_off;
  const uint16_t* pair;
  const float* lut;
  float* x;                  /* in/out [total_nodes][C] fp32 residual stream */
  /* workspace, all [total_nodes][*] in dtype */
  void* h;                   /* [total_nodes][C]   LN output / attention output */
  void* qkv;                 /* [total_nodes][3C] */
  void* ff;                  /* [total_nodes][4C] */
  /* final LN outputs */
  void* dec_in; int32_t dec_dtype;         /* rows scattered by dst_row */
  const int32_t* dst_row;                  /* [total_nodes] or NULL (identity) */
  float* emb_f32;                          /* optional [total_nodes][C] */
  int32_t* ln_counters;                    /* optional int32[ceil(total_nodes/128)] scratch: enables the LayerNorms
                                              fused into the residual GEMMs (only the first and the final one are
                                              then separate launches) */
  void* h2;                                /* [total_nodes][C] in dtype, needed when ln_counters is given */
  int32_t tf32_x3;                         /* dtype TF32 only: activations stay un-rounded fp32, GEMMs use the
                                              3-term compensated tf32 mode (~fp32 accuracy) */
  int32_t skip_final_ln;                   /* 1: stop after the last layer; the caller runs the final LayerNorm itself
                                              (ghn3_layernorm with dst_row), e.g. on another stream */
} ghn3_graphormer_args;

int ghn3_graphormer_stack(const ghn3_graphormer_args* args, ghn3_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * (3c) The same L Graphormer layers (ghn3/graphormer.py:119-142,208-248; ghn3/nn.py:258-261) as ONE persistent
 * kernel for small batches (total_nodes up to a few thousand), bf16 operands. One CTA per SM walks the layers; per
 * layer five stages of tiles -- LN1+QKV GEMM, attention, out-proj (+residual), LN2+FFN1 (+GELU), FFN2 (+residual,
 * split-K) -- whose only synchronisation is a set of per-16-row arrival counters in global memory (no grid barrier,
 * no kernel boundary). GEMM tiles are 128 weight rows (UMMA M) x rb activation rows (UMMA N) on tcgen05 with the
 * accumulator in TMEM; weight blocks stream through a TMA ring that runs ahead of the dependencies; LayerNorm is the
 * prologue of the GEMM that consumes it; attention is the flash-style mma.sync kernel of (3b).
 * Weights of all layers are stacked along rows so that four TMA descriptors cover the stack. The final LayerNorm
 * is NOT included (run ghn3_layernorm afterwards).
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t hid, heads, layers;
  const void* w_qkv;                       /* bf16 [layers*3C][C] */
  const void* w_out;                       /* bf16 [layers*C][C]  */
  const void* w_ff1;                       /* bf16 [layers*4C][C] */
  const void* w_ff2;                       /* bf16 [layers*C][4C] */
  const ghn3_layer_weights* layers_dev;    /* DEVICE array [layers]: only the fp32 vectors (ln*, b_*) are read */
  int32_t n_graphs, total_nodes, max_nodes, lut_size;
  const int32_t* node_off;
  const int64_t* mat_off;
  const uint16_t* pair;
  const float* lut;
  float* x;                  /* in/out [total_nodes][C] fp32 residual stream */
  void* ao;                  /* bf16 [total_nodes][C]      attention output            */
  void* qkv;                 /* bf16 [2][total_nodes][3C]  (double-buffered by layer)  */
  void* ff;                  /* bf16 [total_nodes][4C]                                 */
  int32_t* sync;             /* int32 [ghn3_graphormer_fused_sync_ints(total_nodes)] arrival counters (zeroed here) */
  int32_t stop_after;        /* bring-up aid: > 0 stops after that many stages (5 per layer); 0 = run everything */
  int32_t max_ctas;          /* 0 = one CTA per SM */
} ghn3_graphormer_fused_args;

int64_t ghn3_graphormer_fused_sync_ints(int32_t total_nodes);
int ghn3_graphormer_fused(const ghn3_graphormer_fused_args* args, ghn3_stream_t stream);
/* bring-up aid: [n_ctas][1024][3] int64 (tag, clock64, globaltimer) records per CTA; NULL switches tracing off */
int ghn3_debug_fused_trace(void* device_buffer);

/* ---------------------------------------------------------------------------------------------------------------
 * (4b) Tile / slice / normalise / scatter -- replaces _tile_params + _normalize + _set_params
 * (ghn3/nn.py:422-506, 554-592, 508-552): every predicted tensor of a model is written straight into the target
 * parameter storage by ONE launch driven by a descriptor table.
 *
 * Target element with (padded-to-4D) index (a, b, y, x), dims (t0, t1, t2, t3):
 *   row = a*ra + (y+cy)*kw_src + (x+cx);   col = (a % so)*ca + (b % si);   v = src[row*ld + col]
 *   mode 0: dst = v * scale      (fan-in normalisation, nn.py:583; scale = 1 for positional encodings)
 *   mode 1: dst = 2*sigmoid(0.5 v)   (1-D weights, nn.py:588)
 *   mode 2: dst = tanh(0.2 v)        (1-D biases,  nn.py:590)
 *   mode 3: like mode 0 but (y, x) are bilinearly resampled from a kh_src x kw_src source window (nn.py:751-753)
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct {
  float* dst;
  const float* src;
  int64_t numel;
  int64_t chunk0;        /* index of this tensor's first chunk (prefix sum of ceil(numel / GHN3_SCATTER_CHUNK)) */
  int32_t t1, t2, t3;
  int32_t so, si;
  int32_t ld, ca, ra;
  int32_t kh_src, kw_src, cy, cx;
  float scale;
  int32_t mode;
  /* multiply-high division constants for d in {t1, t2, t3, so, si}: d == 1 -> m = 0; else with s = ceil(log2 d):
   * m = ceil(2^(31+s) / d), shift = s - 1; n / d == umulhi(n, m) >> shift for every n < 2^31 */
  uint32_t m_t1, s_t1, m_t2, s_t2, m_t3, s_t3, m_so, s_so, m_si, s_si;
  int32_t norm_slot;     /* >= 0: the sum of squares of the written values is added to norm_out[norm_slot] */
  int32_t reserved;
} ghn3_scatter_desc;       /* 136 bytes; numel < 2^31 */

#define GHN3_SCATTER_CHUNK 16384

typedef struct {
  const ghn3_scatter_desc* descs;   /* device [n_descs] */
  int32_t n_descs;
  int64_t n_chunks;
  const int32_t* chunk_desc;        /* optional device [n_chunks]: descriptor index of every chunk (else binary search) */
  double* norm_out;                 /* optional device [n_norm_slots]: zeroed, then receives sum(v^2) per norm_slot
                                       (the norm_check metric of ghn3/nn.py:783-797 without re-reading the parameters) */
  int32_t n_norm_slots;
} ghn3_scatter_args;

int ghn3_scatter(const ghn3_scatter_args* args, ghn3_stream_t stream);

/* Sum of squares of a set of fp32 tensors (norm_check metric, ghn3/nn.py:783-797): out[0] += sum_i |t_i|^2. */
typedef struct {
  const float* const* ptrs;    /* device array of device pointers */
  const int64_t* numels;       /* device */
  int32_t n;
  double* out;                 /* device, one double, zeroed by the call */
} ghn3_sumsq_args;

int ghn3_sumsq(const ghn3_sumsq_args* args, ghn3_stream_t stream);

/* dst[(z*cols + c)][r] = relu(src[z*src_bs + r*ld + c]) converted to dst_dtype, for r < rows, c < cols, z < batch:
 * turns the (ms x i') prediction block of a classification-weight node into the K-major operand of the class-head
 * GEMM (reference ghn3/nn.py:757-758: class_layer_predictor = [ReLU, Linear] applied along the out-channel axis). */
typedef struct {
  const float* src; int64_t ld; int64_t src_bs;
  void* dst; int32_t dst_dtype;
  int32_t rows, cols, batch;
} ghn3_relu_transpose_args;
int ghn3_relu_transpose(const ghn3_relu_transpose_args* args, ghn3_stream_t stream);

/* ===============================================================================================================
 * Training path: adjoints of the kernels above, used by GHN3.forward(keep_grads=True) so that
 * autograd can flow from the target networks' loss into the GHN parameters (reference: ghn3/trainer.py:238-411,
 * ghn3/nn.py:526-545). Gradients of GHN parameters are fp32 and are ACCUMULATED (+=) into caller-zeroed buffers
 * unless stated otherwise.
 * ============================================================================================================= */

/* dst[c][r] = src[row(r)][c], row(r) = (r / group) * group_stride + r % group when group > 0, else r.
 * src/dst dtypes are GHN3_BF16 / GHN3_TF32 / GHN3_F32 storage tags. Produces the K-major operands of the dgrad /
 * wgrad GEMMs (W^T, X^T, dY^T). */
typedef struct {
  const void* src; int32_t src_dtype; int64_t ld_src;
  int32_t rows, cols;
  int32_t group, group_stride;
  void* dst; int32_t dst_dtype; int64_t ld_dst;
  /* optional fused extras; v(r, c) below is the value that gets transposed */
  const void* mul_gelu_grad; int32_t mul_dtype;  /* v = src * gelu'(mul_gelu_grad[r][c]) (dense [rows][cols]); NULL: v = src */
  void* copy_out; int32_t copy_dtype;            /* un-transposed dense [rows][cols] copy of v (may alias src) */
  float* colsum_out;                             /* colsum_out[c] += sum_r v(r, c) */
} ghn3_transpose_args;
int ghn3_transpose(const ghn3_transpose_args* args, ghn3_stream_t stream);

enum ghn3_elementwise_op {
  GHN3_EW_COPY = 0,       /* out = a (dtype conversion) */
  GHN3_EW_GELU = 1,       /* out = gelu(a), exact erf form (ghn3/nn.py:142) */
  GHN3_EW_GELU_BWD = 2,   /* out = a * gelu'(b) */
  GHN3_EW_RELU_BWD = 3,   /* out = a * (b > 0) */
  GHN3_EW_ADD = 4         /* out = a + b */
};
typedef struct {
  int32_t op;
  int64_t n;
  const void* a; int32_t a_dtype;
  const void* b; int32_t b_dtype;
  void* out; int32_t out_dtype;
} ghn3_elementwise_args;
int ghn3_elementwise(const ghn3_elementwise_args* args, ghn3_stream_t stream);

/* dst[col(c)] += sum_r src[r][c] (bias gradients); col(c) uses the same (group, group_stride) mapping as above. */
typedef struct {
  const void* src; int32_t src_dtype; int64_t ld;
  int32_t rows, cols;
  int32_t group, group_stride;
  float* dst;
} ghn3_colsum_args;
int ghn3_colsum(const ghn3_colsum_args* args, ghn3_stream_t stream);

/* LayerNorm backward (eps 1e-5): dx (+)= dLN/dx . dy, dgamma += sum dy * xhat, dbeta += sum dy.
 * dy_row (optional): the gradient of row r is row dy_row[r] of dy (< 0: no gradient) -- the adjoint of the row
 * scatter of ghn3_layernorm (final LayerNorm -> decoder input rows). */
typedef struct {
  int32_t rows, hid;
  const float* x;            /* forward input [rows][C] */
  const float* gamma;        /* NULL: adjoint of the identity form (dx (+)= dy through dy_row; dgamma / dbeta unused) */
  const void* dy; int32_t dy_dtype;
  const int32_t* dy_row;
  float* dx;                 /* [rows][C] fp32 */
  int32_t accumulate;        /* 1: dx += ..., 0: dx = ... (rows without gradient are left untouched) */
  float* dgamma; float* dbeta;
} ghn3_layernorm_bwd_args;
int ghn3_layernorm_bwd(const ghn3_layernorm_bwd_args* args, ghn3_stream_t stream);

/* Attention backward (adjoint of ghn3_attention): given qkv, the forward output and d_out, writes d_qkv [M][3C]
 * (same dtype) and accumulates the edge-bias gradient d_lut[h][pair(i,j)] += dS_ij (the bias is shared by all
 * layers, ghn3/graphormer.py:126-130). Softmax statistics are recomputed; lse / delta are [H][total_nodes] fp32
 * workspaces. */
typedef struct {
  int32_t n_graphs, hid, heads, max_nodes, total_nodes;
  int32_t lut_size;
  const int32_t* node_off;
  const int64_t* mat_off;
  const void* qkv;
  const void* out;           /* forward attention output [M][C] */
  const void* d_out;         /* [M][C] */
  int32_t dtype;             /* storage of qkv / out / d_out / d_qkv: GHN3_BF16, else fp32 */
  const uint16_t* pair;
  const float* lut;
  void* d_qkv;               /* out [M][3C] */
  float* d_lut;              /* += [H][lut_size], may be NULL */
  float* lse; float* delta;  /* workspaces [H][total_nodes] */
  /* Tensor-core path (dtype GHN3_BF16 and fwd_lse2 != NULL): mma.sync kernels that take the softmax statistics the
   * forward kept (ghn3_attention_args.lse2) instead of recomputing them. The edge-bias gradient is then NOT added to
   * d_lut; dS is accumulated into ds_total (fp32, graph g / head h plane at mat_off[g] * heads + h * n_g * ld_g,
   * caller-zeroed, may be NULL when the bias gradient is not wanted) and binned once by ghn3_lut_bin. */
  const float* fwd_lse2;
  float* ds_total;
} ghn3_attention_bwd_args;
int ghn3_attention_bwd(const ghn3_attention_bwd_args* args, ghn3_stream_t stream);

/* d_lut[h][pair[i][j]] += ds_total[g][h][i][j]: the edge-bias gradient of ALL layers in one pass (the bias is shared by
 * the layers, ghn3/graphormer.py:126-130). */
typedef struct {
  int32_t n_graphs, heads, max_nodes, lut_size;
  const int32_t* node_off;
  const int64_t* mat_off;
  const uint16_t* pair;
  const float* ds_total;
  float* d_lut;              /* += [H][lut_size] */
} ghn3_lut_bin_args;
int ghn3_lut_bin(const ghn3_lut_bin_args* args, ghn3_stream_t stream);

/* Adjoint of ghn3_scatter: grads[i] is the gradient of the i-th target tensor (NULL: none), d_src[i] the fp32
 * gradient buffer parallel to descs[i].src (same indexing); contributions are added atomically. */
typedef struct {
  const ghn3_scatter_desc* descs;
  int32_t n_descs;
  int64_t n_chunks;
  const int32_t* chunk_desc;        /* required */
  const float* const* grads;        /* device [n_descs] */
  float* const* d_src;              /* device [n_descs] */
} ghn3_scatter_bwd_args;
int ghn3_scatter_bwd(const ghn3_scatter_bwd_args* args, ghn3_stream_t stream);

/* Adjoint of ghn3_node_features: scatter-add of dx rows into the embedding-table gradients. */
typedef struct {
  int32_t total_nodes, hid;
  const int32_t* op; const int32_t* shape_idx; const int32_t* deg_in; const int32_t* deg_out; const int32_t* dist0;
  const float* dx;
  float* d_embed_op; float* d_embed_ch; float* d_embed_sp; float* d_cent_in; float* d_cent_out; float* d_dist_embed;
} ghn3_node_features_bwd_args;
int ghn3_node_features_bwd(const ghn3_node_features_bwd_args* args, ghn3_stream_t stream);

/* Adjoint of ghn3_edge_lut. workspace: 4*(vmax+1)*C floats. */
typedef struct {
  int32_t hid, heads, vmax;
  const float* edge_embed; const float* w1; const float* b1; const float* w2;
  const float* d_lut;        /* [H][(vmax+1)^2] */
  float* workspace;
  float* d_edge_embed; float* d_w1; float* d_b1; float* d_w2; float* d_b2;
} ghn3_edge_lut_bwd_args;
int ghn3_edge_lut_bwd(const ghn3_edge_lut_bwd_args* args, ghn3_stream_t stream);

/* Decoder fc stage backward (adjoint of the per-position grouped GEMM + ReLU of ghn3/nn.py:738-745; dh0 is the
 * gradient AFTER the ReLU mask). Gradients go straight to the ORIGINAL parameter layout of decoder.fc.0
 * (weight [(j*S*S + pos)][C], bias [j*S*S + pos]) and to the decoder-input rows, all accumulated atomically. */
typedef struct {
  const ghn3_gemm_problem* problems;  /* device: the forward fc problems (a_row0, b_row0 = pos*n_out, m, d_off) */
  int32_t n_problems;
  int32_t max_m;                      /* max problems[].m (sizes the launch) */
  const int32_t* rowmap;              /* forward row map: h0 row of problems[p].d_off + k */
  const void* dh0; int32_t dtype;     /* [R][n_out], GHN3_BF16 or fp32 storage; also the dtype of dec_in */
  const void* dec_in;                 /* [n_dec][C] */
  const float* fc_w;                  /* original fp32 weight */
  int32_t hid, n_out, grid_positions; /* C, 4C, S*S */
  float* d_fc_w; float* d_fc_b; float* d_dec_in;
} ghn3_fc_bwd_args;
int ghn3_fc_bwd(const ghn3_fc_bwd_args* args, ghn3_stream_t stream);

/* Adjoint of ghn3_relu_transpose: d_src[z*src_bs + r*ld + c] += (src > 0) * d_rt[(z*cols + c)*rows + r]. */
typedef struct {
  const float* src; float* d_src; int64_t ld; int64_t src_bs;
  const float* d_rt;
  int32_t rows, cols, batch;
} ghn3_relu_transpose_bwd_args;
int ghn3_relu_transpose_bwd(const ghn3_relu_transpose_bwd_args* args, ghn3_stream_t stream);

/* Decoder conv.2 backward, operand preparation: expands the compact per-column-class gradients of the conv decoder
 * output (class c: [rows_c][o'*i'] at src + src_off) into the full max_shape column space,
 *   x[row0 + r][col(k)] = xt[col(k)][row0 + r] = src_c[r][k],  col(k) = (k / group) * group_stride + k % group
 * (group = i' when i' < max_shape[1], else 0 = identity), and adds the column sums into d_bias[col(k)]
 * (gradient of decoder.conv.2.bias). x / xt must be zeroed by the caller. One 32x32 tile per CTA; tile0 is the
 * prefix sum of ceil(rows/32) * tiles_c, tiles_c = ceil(ld/32). */
typedef struct { int64_t src_off; int32_t row0, rows, ld, group, tile0, tiles_c; } ghn3_expand_seg;
typedef struct {
  const ghn3_expand_seg* segs; int32_t n_segs; int32_t n_tiles;
  const float* src;
  int32_t group_stride;
  void* x; int64_t ld_x;
  void* xt; int64_t ld_xt;
  int32_t dtype;
  float* d_bias;
} ghn3_expand_args;
int ghn3_expand_cols(const ghn3_expand_args* args, ghn3_stream_t stream);

/* Optimizer step of the training path: global-norm clipping (nn.utils.clip_grad_norm_, trainer.py:343-349) and AdamW
 * (decoupled weight decay, bias-corrected) in one pass. grads / exp_avg / exp_avg_sq are FLAT fp32 buffers with the
 * same layout (tensor t at offsets[t], numels[t] elements, offsets multiples of 4); params[t] are the parameter
 * tensors. chunk tables as in ghn3_scatter: chunk0[t] = first chunk of tensor t, chunk_tensor[c] = tensor of chunk c,
 * chunks of GHN3_ADAMW_CHUNK elements. max_norm <= 0 disables clipping; otherwise sumsq (one device double) receives
 * |g|^2 and the coefficient min(1, max_norm / (|g| + 1e-6)) is applied on the fly -- no host synchronisation.
 * bias_correction{1,2} = 1 - beta{1,2}^step are computed by the host. */
#define GHN3_ADAMW_CHUNK 8192
typedef struct {
  float* const* params;        /* device [n_tensors] */
  const float* grads; float* exp_avg; float* exp_avg_sq;
  const int64_t* offsets; const int64_t* numels; const int64_t* chunk0;   /* device [n_tensors] */
  const int32_t* chunk_tensor; /* device [n_chunks] */
  int64_t n_chunks; int64_t total;   /* total = elements of the flat buffers */
  float lr, beta1, beta2, eps, weight_decay, bias_correction1, bias_correction2, max_norm;
  double* sumsq;
  /* Non-finite guard (reference trainer.py:240-257 skips / raises on a NaN loss BEFORE the update): when `skipped` is
   * given, |g|^2 is always computed and the whole step -- parameters AND both moments -- is skipped on the device if
   * |g|^2 or *loss (optional, one device float) is not finite; *skipped is incremented. No host synchronisation.
   * Under data parallelism the gradients are already averaged over ranks, so every rank takes the same decision. */
  const float* loss;
  int32_t* skipped;
  /* Sharded optimizer step (data parallelism with the gradient reduce-scattered over the ranks): only the elements
   * with flat index in [range_lo, range_hi) are updated (both multiples of 4; range_hi = 0 means the whole buffer);
   * chunk_begin = index of the first chunk that intersects the range, n_chunks = how many do. With sumsq_ready != 0
   * *sumsq already holds |g|^2 of the WHOLE averaged gradient (the caller summed the shards' parts over the ranks)
   * and is not recomputed (2 = the same, for the second and later ranges of one step: a skipped step is counted
   * once). sumsq_ready = -1 / -2: no update at all -- |g|^2 of the elements in [range_lo, range_hi) is stored in /
   * added to *sumsq (the per-rank part of the global norm). */
  int64_t range_lo, range_hi;
  int64_t chunk_begin;
  int32_t sumsq_ready;
} ghn3_adamw_args;
int ghn3_adamw(const ghn3_adamw_args* args, ghn3_stream_t stream);

/* Predicted-parameter regularisation of the training loss (reference ghn3/trainer.py:288-294:
 * loss += predparam_wd * sum_p ||p||_F over the predicted tensors) on the flat buffer that holds every predicted
 * parameter of the meta-batch: segment s = elements [seg_off[s], seg_off[s] + seg_numel[s]). Chunk tables as in
 * ghn3_adamw (GHN3_ADAMW_CHUNK elements per chunk).
 *   mode 0 (forward) : sumsq[s] = sum v^2 (double), then *total = coef * sum_s sqrt(sumsq[s])
 *   mode 1 (backward): grad[i] += (*gscale) * coef * src[i] / sqrt(sumsq[s])   (0 where the norm is 0) */
typedef struct {
  const float* src;
  const int64_t* seg_off; const int64_t* seg_numel; const int64_t* chunk0;   /* device [n_segs] */
  const int32_t* chunk_seg;                                                  /* device [n_chunks] */
  int64_t n_chunks; int32_t n_segs; int32_t mode;
  double* sumsq;             /* device [n_segs] */
  float* total;              /* mode 0: one device float */
  float* grad;               /* mode 1: flat, same layout as src */
  const float* gscale;       /* mode 1: one device float (upstream gradient), NULL = 1 */
  float coef;
} ghn3_segnorm_args;
int ghn3_segnorm(const ghn3_segnorm_args* args, ghn3_stream_t stream);

/* Graphormer stack, training flavour. ghn3_graphormer_train_fwd computes the same function as ghn3_graphormer_stack
 * (fwd.x is ignored: the input node features are xs[0]) but keeps every activation of every layer;
 * ghn3_graphormer_bwd consumes them. Buffers are [layers][total_nodes][width] (xs: layers+1), contiguous. */
typedef struct {
  ghn3_graphormer_args fwd;  /* weights, batch description, decoder-input outputs; workspaces h/qkv/ff unused */
  float* xs;                 /* [L+1][M][C] fp32 residual stream: xs[l] = input of layer l, xs[L] = stack output */
  float* xm;                 /* [L][M][C]   fp32 residual stream after the attention block */
  void* h1;                  /* [L][M][C]   LN1 output                (activation dtype) */
  void* qkv;                 /* [L][M][3C] */
  void* ao;                  /* [L][M][C]   attention output */
  void* h2;                  /* [L][M][C]   LN2 output */
  void* u;                   /* [L][M][4C]  FFN pre-activation */
  void* g;                   /* [L][M][4C]  GELU(u) */
  float* lse2;               /* optional [L][H][M] fp32: softmax statistics of the bf16 attention (log2 domain) */
} ghn3_graphormer_train_args;
int ghn3_graphormer_train_fwd(const ghn3_graphormer_train_args* args, ghn3_stream_t stream);

typedef struct {             /* transposed copies of the layer's GEMM weights, GEMM input dtype */
  const void* w_qkv_t;       /* [C][3C] */
  const void* w_out_t;       /* [C][C]  */
  const void* w_ff1_t;       /* [C][4C] */
  const void* w_ff2_t;       /* [4C][C] */
} ghn3_layer_weights_t;

typedef struct {             /* fp32 gradient accumulators (+=), shapes of the corresponding parameters */
  float* ln1_w; float* ln1_b; float* w_qkv; float* w_out; float* b_out;
  float* ln2_w; float* ln2_b; float* w_ff1; float* b_ff1; float* w_ff2; float* b_ff2;
} ghn3_layer_grads;

typedef struct {
  const ghn3_graphormer_train_args* saved;   /* HOST pointer */
  const ghn3_layer_weights_t* layers_t_host; /* HOST array [layers] */
  const ghn3_layer_grads* grads_host;        /* HOST array [layers] */
  float* d_ln_w; float* d_ln_b;              /* final LayerNorm (+=) */
  const void* d_dec_in; int32_t d_dec_dtype; /* gradient wrt the decoder-input rows [n_dec][C] */
  float* d_lut;                              /* [H][lut_size] (+=) */
  float* dx;                                 /* [M][C] fp32: on return the gradient wrt the node features xs[0] */
  /* workspaces */
  void* dxa;                 /* [M][C]  activation dtype */
  void* dh;                  /* [M][C]  activation dtype */
  float* dhf;                /* [M][C]  fp32 */
  void* dqkv;                /* [M][3C] activation dtype */
  void* dff;                 /* [M][4C] activation dtype */
  void* ta; void* tb;        /* [4C][m_pad] activation dtype (transposed operands of the wgrad GEMMs) */
  int32_t m_pad;             /* multiple of 8, >= total_nodes */
  float* lse; float* delta;  /* [H][M] */
  float* ds_total;           /* optional [mat_total * H] fp32 scratch (zeroed by the call): with saved->lse2 it enables
                                the tensor-core attention backward + one ghn3_lut_bin pass */
  int64_t ds_total_bytes;
} ghn3_graphormer_bwd_args;
int ghn3_graphormer_bwd(const ghn3_graphormer_bwd_args* args, ghn3_stream_t stream);

/* Runs a prebuilt sequence of the entry points above with ONE call (the host side of `ghn(model)` is then a single
 * FFI crossing per prediction): ops[i].args points to the argument struct of the entry point named by ops[i].op. */
enum ghn3_opcode {
  GHN3_OP_NODE_FEATURES = 1, GHN3_OP_GRAPHORMER = 2, GHN3_OP_GEMM = 3, GHN3_OP_GEMM_SIMT = 4, GHN3_OP_SCATTER = 5,
  GHN3_OP_RELU_TRANSPOSE = 6,
  /* training path */
  GHN3_OP_GRAPHORMER_TRAIN_FWD = 7, GHN3_OP_GRAPHORMER_BWD = 8, GHN3_OP_TRANSPOSE = 9, GHN3_OP_ELEMENTWISE = 10,
  GHN3_OP_COLSUM = 11, GHN3_OP_LAYERNORM_BWD = 12, GHN3_OP_ATTENTION_BWD = 13, GHN3_OP_SCATTER_BWD = 14,
  GHN3_OP_NODE_FEATURES_BWD = 15, GHN3_OP_EDGE_LUT_BWD = 16, GHN3_OP_FC_BWD = 17, GHN3_OP_RELU_TRANSPOSE_BWD = 18,
  GHN3_OP_EXPAND_COLS = 19, GHN3_OP_MEMSET = 20, GHN3_OP_LAYERNORM = 21, GHN3_OP_GRAPHORMER_FUSED = 22,
  GHN3_OP_MEMCPY = 23, GHN3_OP_FORK = 24, GHN3_OP_JOIN = 25
};
/* GHN3_OP_MEMSET: args points to a ghn3_memset_args; clears `bytes` bytes at `ptr` (cudaMemsetAsync). */
typedef struct { void* ptr; int64_t bytes; } ghn3_memset_args;
/* GHN3_OP_MEMCPY: args points to a ghn3_memcpy_args; device-to-device cudaMemcpyAsync of `bytes` bytes. */
typedef struct { void* dst; const void* src; int64_t bytes; } ghn3_memcpy_args;
/* Two-lane sequences: an op with lane = 1 is issued on an auxiliary stream owned by the library, so that independent
 * small launches (the 1-D decoder, column classes of conv.2 with few tiles) run beside the large ones instead of in
 * front of them. GHN3_OP_FORK (args ignored) makes the auxiliary lane wait for everything issued on the main lane so
 * far; GHN3_OP_JOIN makes the main lane wait for the auxiliary one -- a sequence that used lane 1 must end joined.
 * Captured (ghn3_sequence_capture) the lanes become parallel branches of the graph. */
typedef struct { int32_t op; int32_t lane; const void* args; } ghn3_op;
int ghn3_run_sequence(const ghn3_op* ops, int32_t n, ghn3_stream_t stream);

/* The same sequence recorded once as a CUDA graph and replayed with ONE driver call per prediction (183 kernel launches
 * of `ghn(model)` cost 0.4-0.5 ms of host time when issued one by one; inside a graph dependent kernels also start
 * with less latency). ghn3_sequence_capture records ops[0..n) on an internal capture stream -- nothing executes -- and
 * instantiates the graph; every pointer and size in the argument structs (and the process-wide switches
 * ghn3_set_programmatic_launch / ghn3_set_attention_tc_min) is frozen at that moment, so callers keep the buffers the
 * structs name at fixed addresses and re-capture when anything else changes. high_priority != 0 gives the kernel nodes
 * the priority of the device's highest-priority streams. Run the sequence once with ghn3_run_sequence before capturing
 * it (first launches set function attributes). ghn3_launch_count() advances by the captured launch count per replay. */
typedef struct ghn3_sequence ghn3_sequence;
int ghn3_sequence_capture(const ghn3_op* ops, int32_t n, int32_t high_priority, ghn3_sequence** out);
int ghn3_sequence_launch(ghn3_sequence* seq, ghn3_stream_t stream);
int ghn3_sequence_destroy(ghn3_sequence* seq);

/* dtype conversion helpers used when a checkpoint is prepared for the device (one-time, not on the hot path). */
int ghn3_convert_f32(const float* src, void* dst, int64_t n, int32_t dst_dtype, ghn3_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GHN3_B200_H_ */
