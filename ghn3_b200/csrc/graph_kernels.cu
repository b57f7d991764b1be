// Integer graph kernels and node-feature construction (sm_100a).
//   ghn3_spd_bfs        bitset BFS shortest-path distances   (reference ghn3/graph.py:755-798, networkx on the host)
//   ghn3_graph_derive   (A_ij, A_ji) pair index + degrees + distance from the input node (graphormer.py:229-237)
//   ghn3_node_features  op-type + shape + structural embedding gathers, one pass (nn.py:248-249, graphormer.py:230-232)
//   ghn3_edge_lut       the edge-bias MLP evaluated on the (vmax+1)^2 value grid (graphormer.py:114-117)
// All of these are HBM/latency-bound integer or gather work: coalesced rows, 16-byte vectors, no tensor cores.
#include "common.cuh"

namespace ghn3 {

__device__ __forceinline__ int find_segment(const int32_t* __restrict__ off, int n, int v) {
  int lo = 0, hi = n;   // largest g with off[g] <= v
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= v) lo = mid; else hi = mid;
  }
  return lo;
}

// one thread per 1-hop edge: set bit (src, dst) of the graph's adjacency bit matrix
__global__ void adj_bits_kernel(int n_graphs, int total_edges, const int32_t* __restrict__ node_off,
                                const int32_t* __restrict__ edge_off, const int64_t* __restrict__ bits_off,
                                const int32_t* __restrict__ edge_src, const int32_t* __restrict__ edge_dst,
                                uint32_t* __restrict__ bits) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total_edges) return;
  const int g = find_segment(edge_off, n_graphs, e);
  const int n = node_off[g + 1] - node_off[g];
  const int words = (n + 31) >> 5;
  const int s = edge_src[e], d = edge_dst[e];
  if (s == d || s < 0 || d < 0 || s >= n || d >= n) return;   // no self loops (graph.py:764)
  atomicOr(bits + bits_off[g] + (int64_t)s * words + (d >> 5), 1u << (d & 31));
}

// One warp per BFS source. Lane l owns bit-words l, l+32, ... (WPL words per lane => graphs up to 1024*WPL nodes).
// Level d: every newly reached node gets distance d; the next frontier is the OR of the adjacency rows of the
// current frontier minus everything visited. Stops at `cutoff` (graph.py:792: cutoff=ve_cutoff).
template <int WPL>
__global__ void __launch_bounds__(256) spd_bfs_kernel(int cutoff, const int32_t* __restrict__ node_off,
                                                      const int64_t* __restrict__ mat_off,
                                                      const int64_t* __restrict__ bits_off,
                                                      const uint32_t* __restrict__ bits, uint8_t* __restrict__ spd) {
  extern __shared__ uint8_t dist_smem[];
  const int g = blockIdx.y;
  const int n = node_off[g + 1] - node_off[g];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int src = blockIdx.x * 8 + warp;
  if (src >= n) return;
  const int words = (n + 31) >> 5;
  const int ld = (n + 15) & ~15;
  const uint32_t* __restrict__ adj = bits + bits_off[g];
  uint8_t* dist = dist_smem + (size_t)warp * (WPL * 1024);
  for (int j = lane * 4; j < ld; j += 128) *(uint32_t*)(dist + j) = 0u;
  __syncwarp();

  uint32_t visited[WPL], frontier[WPL];
#pragma unroll
  for (int k = 0; k < WPL; ++k) {
    const int w = lane + 32 * k;
    visited[k] = (w == (src >> 5)) ? (1u << (src & 31)) : 0u;
    frontier[k] = (w < words) ? (__ldg(adj + (int64_t)src * words + w) & ~visited[k]) : 0u;
  }
  for (int d = 1; d <= cutoff; ++d) {
    uint32_t any = 0;
#pragma unroll
    for (int k = 0; k < WPL; ++k) any |= frontier[k];
    if (__ballot_sync(0xffffffffu, any != 0) == 0) break;
#pragma unroll
    for (int k = 0; k < WPL; ++k) {
      uint32_t f = frontier[k];
      while (f) {
        const int b = __ffs(f) - 1;
        f &= f - 1;
        dist[(lane + 32 * k) * 32 + b] = (uint8_t)d;
      }
      visited[k] |= frontier[k];
    }
    if (d == cutoff) break;
    uint32_t next[WPL];
#pragma unroll
    for (int k = 0; k < WPL; ++k) next[k] = 0u;
#pragma unroll
    for (int k = 0; k < WPL; ++k) {
      uint32_t lanes = __ballot_sync(0xffffffffu, frontier[k] != 0);
      while (lanes) {
        const int sl = __ffs(lanes) - 1;
        lanes &= lanes - 1;
        uint32_t f = __shfl_sync(0xffffffffu, frontier[k], sl);
        while (f) {
          const int b = __ffs(f) - 1;
          f &= f - 1;
          const int u = (sl + 32 * k) * 32 + b;
          const uint32_t* row = adj + (int64_t)u * words;
#pragma unroll
          for (int kk = 0; kk < WPL; ++kk) {
            const int w = lane + 32 * kk;
            if (w < words) next[kk] |= __ldg(row + w);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < WPL; ++k) frontier[k] = next[k] & ~visited[k];
  }
  __syncwarp();
  uint8_t* out = spd + mat_off[g] + (int64_t)src * ld;
  for (int j = lane * 4; j < ld; j += 128) *(uint32_t*)(out + j) = *(const uint32_t*)(dist + j);
}

// pair[i][j] = spd[i][j]*(vmax+1) + spd[j][i], 32x32 tiles transposed through shared memory
__global__ void __launch_bounds__(256) pair_kernel(int vmax, const int32_t* __restrict__ node_off,
                                                   const int64_t* __restrict__ mat_off,
                                                   const uint8_t* __restrict__ spd, uint16_t* __restrict__ pair) {
  __shared__ uint8_t tile_t[32][33];
  const int g = blockIdx.z;
  const int n = node_off[g + 1] - node_off[g];
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  if (i0 >= n || j0 >= n) return;
  const int ld = (n + 15) & ~15;
  const uint8_t* A = spd + mat_off[g];
  uint16_t* P = pair + mat_off[g];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int r = ty; r < 32; r += 8) {       // tile of A^T source: rows j0.., cols i0..
    const int jj = j0 + r, ii = i0 + tx;
    tile_t[r][tx] = (jj < n && ii < n) ? A[(int64_t)jj * ld + ii] : 0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int ii = i0 + r, jj = j0 + tx;
    if (ii < n && jj < ld) {
      const int fw = (jj < n) ? A[(int64_t)ii * ld + jj] : 0;
      const int bw = (jj < n) ? tile_t[tx][r] : 0;
      P[(int64_t)ii * ld + jj] = (uint16_t)(fw * (vmax + 1) + bw);
    }
  }
}

// one warp per node: in/out degree over entries equal to 1, distance from node 0
__global__ void __launch_bounds__(256) degree_kernel(int n_graphs, int total_nodes, const int32_t* __restrict__ node_off,
                                                     const int64_t* __restrict__ mat_off,
                                                     const uint8_t* __restrict__ spd, int32_t* __restrict__ deg_in,
                                                     int32_t* __restrict__ deg_out, int32_t* __restrict__ dist0) {
  const int node = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (node >= total_nodes) return;
  const int g = find_segment(node_off, n_graphs, node);
  const int n = node_off[g + 1] - node_off[g];
  const int i = node - node_off[g];
  const int ld = (n + 15) & ~15;
  const uint8_t* A = spd + mat_off[g];
  int cin = 0, cout = 0;
  for (int j = lane; j < n; j += 32) {
    cout += (A[(int64_t)i * ld + j] == 1);
    cin += (A[(int64_t)j * ld + i] == 1);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cin += __shfl_xor_sync(0xffffffffu, cin, o);
    cout += __shfl_xor_sync(0xffffffffu, cout, o);
  }
  if (lane == 0) {
    deg_in[node] = min(cin, 100);
    deg_out[node] = min(cout, 100);
    dist0[node] = min((int)A[i], 1000);
  }
}

// one warp per node, float4 per lane
__global__ void __launch_bounds__(256) node_features_kernel(const ghn3_node_features_args a) {
  const int node = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (node >= a.total_nodes) return;
  const int C = a.hid, Q = C >> 2;
  const int op = a.op[node];
  const int4 si = *(const int4*)(a.shape_idx + 4 * (int64_t)node);
  const int din = a.deg_in[node], dout = a.deg_out[node], d0 = a.dist0[node];
  for (int c = lane * 4; c < C; c += 128) {
    const int q = c / Q, cc = c - q * Q;
    const int sidx = q == 0 ? si.x : (q == 1 ? si.y : (q == 2 ? si.z : si.w));
    const float* stab = q < 2 ? a.embed_ch : a.embed_sp;
    const float4 e = *(const float4*)(a.embed_op + (int64_t)op * C + c);
    const float4 s = *(const float4*)(stab + (int64_t)sidx * Q + cc);
    const float4 ci = *(const float4*)(a.cent_in + (int64_t)din * C + c);
    const float4 co = *(const float4*)(a.cent_out + (int64_t)dout * C + c);
    const float4 di = *(const float4*)(a.dist_embed + (int64_t)d0 * C + c);
    float4 x;
    x.x = (((e.x + s.x) + ci.x) + co.x) + di.x;
    x.y = (((e.y + s.y) + ci.y) + co.y) + di.y;
    x.z = (((e.z + s.z) + ci.z) + co.z) + di.z;
    x.w = (((e.w + s.w) + ci.w) + co.w) + di.w;
    *(float4*)(a.x + (int64_t)node * C + c) = x;
  }
}

// stage 1: P[side][v][c] = sum_k W1[c][side*C + k] * E[v+2][k]; one warp per output
__global__ void __launch_bounds__(256) edge_lut_stage1(int C, int V, const float* __restrict__ E,
                                                       const float* __restrict__ W1, float* __restrict__ P) {
  const int o = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (o >= 2 * V * C) return;
  const int c = o % C, v = (o / C) % V, side = o / (C * V);
  const float* w = W1 + (int64_t)c * 2 * C + side * C;
  const float* e = E + (int64_t)(v + 2) * C;
  float acc = 0.f;
  for (int k = lane; k < C; k += 32) acc = fmaf(w[k], e[k], acc);
  acc = warp_sum(acc);
  if (lane == 0) P[o] = acc;
}

// stage 2: lut[h][a*V+b] = b2[h] + sum_c W2[h][c] * relu(Pa[a][c] + Pb[b][c] + b1[c]); one warp per (a, b)
template <int HMAX>
__global__ void __launch_bounds__(256) edge_lut_stage2(int C, int V, int H, const float* __restrict__ P,
                                                       const float* __restrict__ b1, const float* __restrict__ W2,
                                                       const float* __restrict__ b2, float* __restrict__ lut) {
  const int ab = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (ab >= V * V) return;
  const int a = ab / V, b = ab % V;
  const float* pa = P + (int64_t)a * C;
  const float* pb = P + (int64_t)(V + b) * C;
  float acc[HMAX];
#pragma unroll
  for (int h = 0; h < HMAX; ++h) acc[h] = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float t = fmaxf((pa[c] + pb[c]) + b1[c], 0.f);
#pragma unroll
    for (int h = 0; h < HMAX; ++h)
      if (h < H) acc[h] = fmaf(W2[(int64_t)h * C + c], t, acc[h]);
  }
#pragma unroll
  for (int h = 0; h < HMAX; ++h) {
    if (h < H) {
      const float s = warp_sum(acc[h]);
      if (lane == 0) lut[(int64_t)h * V * V + ab] = s + b2[h];
    }
  }
}

}  // namespace ghn3

using namespace ghn3;

extern "C" int ghn3_spd_bfs(const ghn3_spd_args* a, ghn3_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHN3_REQUIRE(a != nullptr, "ghn3_spd_bfs: null args");
  GHN3_REQUIRE(a->cutoff >= 1 && a->cutoff <= 254, "ghn3_spd_bfs: cutoff must be in [1, 254]");
  GHN3_REQUIRE(a->max_nodes <= 4096, "ghn3_spd_bfs: graphs with more than 4096 nodes are not supported (got %d)",
               a->max_nodes);
  if (a->n_graphs <= 0 || a->total_nodes <= 0) return GHN3_OK;
  GHN3_CUDA(cudaMemsetAsync(a->adj_bits, 0, (size_t)a->bits_total * sizeof(uint32_t), stream));
  if (a->total_edges > 0) {
    adj_bits_kernel<<<(unsigned)ceil_div(a->total_edges, 256), 256, 0, stream>>>(
        a->n_graphs, a->total_edges, a->node_off, a->edge_off, a->bits_off, a->edge_src, a->edge_dst, a->adj_bits);
    GHN3_LAUNCH_CHECK("adj_bits_kernel");
  }
  const dim3 grid((unsigned)ceil_div(a->max_nodes, 8), (unsigned)a->n_graphs);
  if (a->max_nodes <= 1024) {
    spd_bfs_kernel<1><<<grid, 256, 8 * 1024, stream>>>(a->cutoff, a->node_off, a->mat_off, a->bits_off, a->adj_bits, a->spd);
  } else if (a->max_nodes <= 2048) {
    spd_bfs_kernel<2><<<grid, 256, 8 * 2048, stream>>>(a->cutoff, a->node_off, a->mat_off, a->bits_off, a->adj_bits, a->spd);
  } else {
    spd_bfs_kernel<4><<<grid, 256, 8 * 4096, stream>>>(a->cutoff, a->node_off, a->mat_off, a->bits_off, a->adj_bits, a->spd);
  }
  GHN3_LAUNCH_CHECK("spd_bfs_kernel");
  return GHN3_OK;
}

extern "C" int ghn3_graph_derive(const ghn3_derive_args* a, ghn3_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHN3_REQUIRE(a != nullptr, "ghn3_graph_derive: null args");
  GHN3_REQUIRE(a->vmax >= 1 && (a->vmax + 1) * (a->vmax + 1) <= 65536, "ghn3_graph_derive: vmax out of range");
  if (a->n_graphs <= 0 || a->total_nodes <= 0) return GHN3_OK;
  const unsigned t = (unsigned)ceil_div(a->max_nodes, 32);
  pair_kernel<<<dim3(t, t, (unsigned)a->n_graphs), 256, 0, stream>>>(a->vmax, a->node_off, a->mat_off, a->spd, a->pair);
  GHN3_LAUNCH_CHECK("pair_kernel");
  degree_kernel<<<(unsigned)ceil_div(a->total_nodes, 8), 256, 0, stream>>>(a->n_graphs, a->total_nodes, a->node_off,
                                                                           a->mat_off, a->spd, a->deg_in, a->deg_out,
                                                                           a->dist0);
  GHN3_LAUNCH_CHECK("degree_kernel");
  return GHN3_OK;
}

extern "C" int ghn3_node_features(const ghn3_node_features_args* a, ghn3_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHN3_REQUIRE(a != nullptr, "ghn3_node_features: null args");
  GHN3_REQUIRE(a->hid > 0 && a->hid % 16 == 0, "ghn3_node_features: hid must be a positive multiple of 16 (got %d)", a->hid);
  if (a->total_nodes <= 0) return GHN3_OK;
  node_features_kernel<<<(unsigned)ceil_div(a->total_nodes, 8), 256, 0, stream>>>(*a);
  GHN3_LAUNCH_CHECK("node_features_kernel");
  return GHN3_OK;
}

extern "C" int ghn3_edge_lut(const ghn3_edge_lut_args* a, ghn3_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHN3_REQUIRE(a != nullptr, "ghn3_edge_lut: null args");
  GHN3_REQUIRE(a->heads >= 1 && a->heads <= 32, "ghn3_edge_lut: heads must be in [1, 32]");
  GHN3_REQUIRE(a->vmax >= 1 && a->vmax + 2 < 257, "ghn3_edge_lut: vmax out of range");
  const int V = a->vmax + 1, C = a->hid;
  edge_lut_stage1<<<(unsigned)ceil_div(2 * V * C, 8), 256, 0, stream>>>(C, V, a->edge_embed, a->w1, a->workspace);
  GHN3_LAUNCH_CHECK("edge_lut_stage1");
  if (a->lut == nullptr) return GHN3_OK;       // projections only (used by ghn3_edge_lut_bwd)
  if (a->heads <= 8) {
    edge_lut_stage2<8><<<(unsigned)ceil_div(V * V, 8), 256, 0, stream>>>(C, V, a->heads, a->workspace, a->b1, a->w2, a->b2, a->lut);
  } else if (a->heads <= 16) {
    edge_lut_stage2<16><<<(unsigned)ceil_div(V * V, 8), 256, 0, stream>>>(C, V, a->heads, a->workspace, a->b1, a->w2, a->b2, a->lut);
  } else {
    edge_lut_stage2<32><<<(unsigned)ceil_div(V * V, 8), 256, 0, stream>>>(C, V, a->heads, a->workspace, a->b1, a->w2, a->b2, a->lut);
  }
  GHN3_LAUNCH_CHECK("edge_lut_stage2");
  return GHN3_OK;
}
