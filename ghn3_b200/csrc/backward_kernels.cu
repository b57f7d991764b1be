// Backward kernels of the GHN-3 training path (reference: ghn3/trainer.py:238-411 runs autograd through
// GHN3.forward(keep_grads=True), ghn3/nn.py:186-349). Each kernel is the hand-written adjoint of one forward kernel
// of this library; the linear layers reuse the tcgen05 GEMM on explicitly transposed operands (ghn3_transpose).
// All gradients of GHN parameters are accumulated in fp32.
#include "common.cuh"

namespace ghn3 {

__device__ __forceinline__ float ld_f(const void* p, int64_t i, int dt) {
  return dt == GHN3_BF16 ? __bfloat162float(((const __nv_bfloat16*)p)[i]) : ((const float*)p)[i];
}
__device__ __forceinline__ void st_f(void* p, int64_t i, int dt, float v) {
  if (dt == GHN3_BF16) ((__nv_bfloat16*)p)[i] = __float2bfloat16_rn(v);
  else if (dt == GHN3_TF32) ((float*)p)[i] = round_tf32(v);
  else ((float*)p)[i] = v;
}

// ---------------------------------------------------------------------------------------------------------------
// dst[c][r] = v(r, c), v = src[row(r)][c] (* gelu'(mul[r][c])): 32 x 32 tiles through shared memory, dtype conversion
// on the way; optionally also the un-transposed copy of v and its column sums (bias gradients)
__device__ __forceinline__ float gelu_grad_f(float u) {
  return 0.5f * (1.f + erff(u * 0.70710678118654752440f)) + u * 0.39894228040143267794f * __expf(-0.5f * u * u);
}

__global__ void __launch_bounds__(256) transpose_kernel(const ghn3_transpose_args a) {
  __shared__ float tile[32][33];
  pdl_launch_dependents();
  pdl_wait();
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    float v = 0.f;
    if (r < a.rows && c < a.cols) {
      const int64_t sr = a.group > 0 ? (int64_t)(r / a.group) * a.group_stride + r % a.group : r;
      v = ld_f(a.src, sr * a.ld_src + c, a.src_dtype);
      if (a.mul_gelu_grad != nullptr) v *= gelu_grad_f(ld_f(a.mul_gelu_grad, (int64_t)r * a.cols + c, a.mul_dtype));
      if (a.copy_out != nullptr) st_f(a.copy_out, (int64_t)r * a.cols + c, a.copy_dtype, v);
    }
    tile[ty + 8 * i][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, r = r0 + tx;
    if (c < a.cols && r < a.rows) st_f(a.dst, (int64_t)c * a.ld_dst + r, a.dst_dtype, tile[tx][ty + 8 * i]);
  }
  if (a.colsum_out != nullptr && ty == 0 && c0 + tx < a.cols) {
    float t = 0.f;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) t += tile[r][tx];
    if (t != 0.f) atomicAdd(a.colsum_out + c0 + tx, t);
  }
}

// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_f(float u) { return 0.5f * u * (1.f + erff(u * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad(float u) { return gelu_grad_f(u); }

__device__ __forceinline__ float ew_apply(int op, float x, float y) {
  switch (op) {
    case GHN3_EW_GELU: return gelu_f(x);
    case GHN3_EW_GELU_BWD: return x * gelu_grad(y);
    case GHN3_EW_RELU_BWD: return y > 0.f ? x : 0.f;
    case GHN3_EW_ADD: return x + y;
    default: return x;                         // GHN3_EW_COPY
  }
}

// four consecutive elements of a dtype-tagged array (i multiple of 4, base 16-byte aligned)
__device__ __forceinline__ float4 ld4_f(const void* p, int64_t i, int dt) {
  if (dt == GHN3_BF16) {
    const uint2 raw = *(const uint2*)((const __nv_bfloat16*)p + i);
    const __nv_bfloat162 a = *(const __nv_bfloat162*)&raw.x, b = *(const __nv_bfloat162*)&raw.y;
    return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
  }
  return *(const float4*)((const float*)p + i);
}
__device__ __forceinline__ void st4_f(void* p, int64_t i, int dt, float4 v) {
  if (dt == GHN3_BF16) {
    uint2 raw;
    raw.x = pack_bf16(v.x, v.y);
    raw.y = pack_bf16(v.z, v.w);
    *(uint2*)((__nv_bfloat16*)p + i) = raw;
  } else {
    if (dt == GHN3_TF32) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
    *(float4*)((float*)p + i) = v;
  }
}

__global__ void __launch_bounds__(256) elementwise_kernel(const ghn3_elementwise_args a, int vec) {
  pdl_launch_dependents();
  pdl_wait();
  const int op = a.op;
  const bool two = a.b != nullptr;
  if (vec) {                                   // 4 elements per thread and step (8- / 16-byte accesses)
    const int64_t n4 = a.n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
      const float4 x = ld4_f(a.a, i * 4, a.a_dtype);
      const float4 y = two ? ld4_f(a.b, i * 4, a.b_dtype) : make_float4(0.f, 0.f, 0.f, 0.f);
      st4_f(a.out, i * 4, a.out_dtype,
            make_float4(ew_apply(op, x.x, y.x), ew_apply(op, x.y, y.y), ew_apply(op, x.z, y.z), ew_apply(op, x.w, y.w)));
    }
    return;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    const float y = two ? ld_f(a.b, i, a.b_dtype) : 0.f;
    st_f(a.out, i, a.out_dtype, ew_apply(op, ld_f(a.a, i, a.a_dtype), y));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// dst[col(c)] += sum_r src[r][c]
__global__ void __launch_bounds__(256) colsum_kernel(const ghn3_colsum_args a) {
  __shared__ float part[8][33];
  pdl_launch_dependents();
  pdl_wait();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * 256;
  const int r1 = min(r0 + 256, a.rows);
  float acc = 0.f;
  if (c < a.cols)
    for (int r = r0 + ty; r < r1; r += 8) acc += ld_f(a.src, (int64_t)r * a.ld + c, a.src_dtype);
  part[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < a.cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][tx];
    const int64_t dc = a.group > 0 ? (int64_t)(c / a.group) * a.group_stride + c % a.group : c;
    atomicAdd(a.dst + dc, t);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm backward (eps 1e-5): one warp per row, per-block shared accumulators for dgamma / dbeta
// kLnMaxT = ceil(hid / 32) rounded up to an instantiated size: the per-lane arrays stay in registers
template <int kLnMaxT>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const ghn3_layernorm_bwd_args a) {
  extern __shared__ float ln_smem[];
  const int C = a.hid;
  float* sG = ln_smem;
  float* sB = ln_smem + C;
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) ln_smem[i] = 0.f;
  pdl_wait();
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float invC = 1.f / (float)C;
  float ga[kLnMaxT], ba[kLnMaxT];            // this lane's dgamma / dbeta partial sums over the rows of its warp
#pragma unroll
  for (int t = 0; t < kLnMaxT; ++t) { ga[t] = 0.f; ba[t] = 0.f; }
  for (int row = blockIdx.x * 8 + warp; row < a.rows; row += gridDim.x * 8) {
    const int64_t dyr = a.dy_row ? a.dy_row[row] : row;
    if (dyr < 0) continue;                     // no gradient reaches this row through this LayerNorm
    if (a.gamma == nullptr) {                  // identity form (layernorm=False): dx (+)= dy
      float* dxi = a.dx + (int64_t)row * C;
      for (int c = lane; c < C; c += 32) {
        const float dy = ld_f(a.dy, dyr * C + c, a.dy_dtype);
        dxi[c] = a.accumulate ? dxi[c] + dy : dy;
      }
      continue;
    }
    const float* x = a.x + (int64_t)row * C;
    float xv[kLnMaxT], gv[kLnMaxT];
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < kLnMaxT; ++t) {
      const int c = lane + 32 * t;
      xv[t] = c < C ? x[c] : 0.f;
      sum += xv[t];
    }
    const float mean = warp_sum(sum) * invC;
    float sq = 0.f;
#pragma unroll
    for (int t = 0; t < kLnMaxT; ++t) {
      const int c = lane + 32 * t;
      const float d = c < C ? xv[t] - mean : 0.f;
      sq += d * d;
    }
    const float rstd = 1.0f / sqrtf(warp_sum(sq) * invC + 1e-5f);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int t = 0; t < kLnMaxT; ++t) {
      const int c = lane + 32 * t;
      if (c < C) {
        const float dy = ld_f(a.dy, dyr * C + c, a.dy_dtype);
        const float xh = (xv[t] - mean) * rstd;
        xv[t] = xh;
        const float g = dy * a.gamma[c];
        gv[t] = g;
        s1 += g;
        s2 += g * xh;
        ga[t] += dy * xh;
        ba[t] += dy;
      } else {
        gv[t] = 0.f;
      }
    }
    s1 = warp_sum(s1) * invC;
    s2 = warp_sum(s2) * invC;
    float* dx = a.dx + (int64_t)row * C;
#pragma unroll
    for (int t = 0; t < kLnMaxT; ++t) {
      const int c = lane + 32 * t;
      if (c < C) {
        const float v = rstd * (gv[t] - s1 - xv[t] * s2);
        dx[c] = a.accumulate ? dx[c] + v : v;
      }
    }
  }
#pragma unroll
  for (int t = 0; t < kLnMaxT; ++t) {
    const int c = lane + 32 * t;
    if (c < C) {
      if (ga[t] != 0.f) atomicAdd(sG + c, ga[t]);
      if (ba[t] != 0.f) atomicAdd(sB + c, ba[t]);
    }
  }
  __syncthreads();
  if (a.gamma == nullptr) return;              // identity form: nothing to accumulate
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    if (sG[c] != 0.f) atomicAdd(a.dgamma + c, sG[c]);
    if (sB[c] != 0.f) atomicAdd(a.dbeta + c, sB[c]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Attention backward, fp32 math on CUDA cores. Softmax statistics are recomputed (no state saved by the forward).
//   S = Q K^T d^-1/2 + lut[h][pair];  P = softmax(S);  O = P V
//   delta_i = dO_i . O_i;  dP = dO V^T;  dS = P o (dP - delta);  dQ = dS K d^-1/2;  dK = dS^T Q d^-1/2;  dV = P^T dO
//   dlut[h][pair(i,j)] += dS_ij
constexpr int kBwdKT = 256;     // keys (kernel A) / queries (kernel B) staged per shared-memory tile
constexpr int kBwdWarps = 8;
constexpr int kBwdBlock = 64;   // queries (A) / keys (B) owned by a CTA: 8 per warp, state kept in shared memory

// rows x D slice of a [*, ld] matrix of T -> fp32 shared-memory tile with row stride D+1 (bank-conflict free),
// 16-byte global loads when the slice allows it
template <typename T, int D>
__device__ __forceinline__ void bwd_load_tile(float* dst, const T* base, int64_t ld, int rows, float scale) {
  constexpr int DP = D + 1;
  constexpr int EPV = 16 / (int)sizeof(T);                 // elements per 16-byte vector
  if constexpr (D % EPV == 0) {
    const bool aligned = ((((uintptr_t)base) & 15) == 0) && ((ld * sizeof(T)) % 16 == 0);
    if (aligned) {
      constexpr int VPR = D / EPV;
      for (int idx = threadIdx.x; idx < rows * VPR; idx += blockDim.x) {
        const int j = idx / VPR, v = idx - j * VPR;
        const uint4 raw = *(const uint4*)(base + (int64_t)j * ld + v * EPV);
        const T* e = (const T*)&raw;
#pragma unroll
        for (int t = 0; t < EPV; ++t) dst[j * DP + v * EPV + t] = to_float(e[t]) * scale;
      }
      return;
    }
  }
  for (int idx = threadIdx.x; idx < rows * D; idx += blockDim.x) {
    const int j = idx / D, d = idx - j * D;
    dst[j * DP + d] = to_float(base[(int64_t)j * ld + d]) * scale;
  }
}

// Kernel A: one CTA = 64 queries of one (graph, head). Pass 0 recomputes the softmax statistics, pass 1 the gradients;
// the key/value tiles are the OUTER loop (loaded once per pass), every warp walks its 8 queries per tile.
template <typename T, int D>
__global__ void __launch_bounds__(256, (D <= 24) ? 2 : 1) attention_bwd_dq_kernel(const ghn3_attention_bwd_args a) {
  extern __shared__ float bw_smem[];
  constexpr int DP = D + 1;
  constexpr int U = kBwdKT / 32;
  constexpr int QPW = kBwdBlock / kBwdWarps;
  float* sK = bw_smem;                         // [KT][DP]
  float* sV = sK + kBwdKT * DP;                // [KT][DP]
  float* sQ = sV + kBwdKT * DP;                // [64][DP]  (pre-scaled by d^-1/2)
  float* sdO = sQ + kBwdBlock * DP;            // [64][DP]
  float* sdQ = sdO + kBwdBlock * DP;           // [64][DP]  gradient accumulator
  float* sM = sdQ + kBwdBlock * DP;            // [64] running max
  float* sL = sM + kBwdBlock;                  // [64] running sum
  float* sDelta = sL + kBwdBlock;              // [64]
  float* sBias = sDelta + kBwdBlock;           // [warps][KT]
  uint16_t* sPair = (uint16_t*)(sBias + kBwdWarps * kBwdKT);   // [warps][KT]
  float* sLut = (float*)(sPair + kBwdWarps * kBwdKT);          // [lut_size]
  float* sHist = sLut + a.lut_size;                            // [lut_size]

  const int g = blockIdx.z, h = blockIdx.y;
  const int n0 = a.node_off[g];
  const int n = a.node_off[g + 1] - n0;
  const int q0 = blockIdx.x * kBwdBlock;
  if (q0 >= n) return;
  const int nq = min(kBwdBlock, n - q0);
  const int ld = (n + 15) & ~15;
  const int C = a.hid, C3 = 3 * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const T* qkv = (const T*)a.qkv + (int64_t)n0 * C3;
  const T* dout = (const T*)a.d_out + (int64_t)n0 * C;
  const T* out = (const T*)a.out + (int64_t)n0 * C;
  const uint16_t* pair = a.pair + a.mat_off[g];
  const float scale = rsqrtf((float)D);

  pdl_launch_dependents();
  for (int i = threadIdx.x; i < a.lut_size; i += blockDim.x) {
    sLut[i] = a.lut[(int64_t)h * a.lut_size + i];
    sHist[i] = 0.f;
  }
  pdl_wait();
  bwd_load_tile<T, D>(sQ, qkv + (int64_t)q0 * C3 + h * D, C3, nq, scale);
  bwd_load_tile<T, D>(sdO, dout + (int64_t)q0 * C + h * D, C, nq, 1.f);
  for (int i = threadIdx.x; i < kBwdBlock * DP; i += blockDim.x) sdQ[i] = 0.f;
  for (int i = threadIdx.x; i < kBwdBlock; i += blockDim.x) { sM[i] = -INFINITY; sL[i] = 0.f; }
  __syncthreads();
  // delta_i = dO_i . O_i
  for (int lq = warp; lq < nq; lq += kBwdWarps) {
    float t = 0.f;
    for (int d = lane; d < D; d += 32) t += sdO[lq * DP + d] * to_float(out[(int64_t)(q0 + lq) * C + h * D + d]);
    t = warp_sum(t);
    if (lane == 0) sDelta[lq] = t;
  }

  float* myBias = sBias + warp * kBwdKT;
  uint16_t* myPair = sPair + warp * kBwdKT;
  for (int pass = 0; pass < 2; ++pass) {
    for (int k0 = 0; k0 < n; k0 += kBwdKT) {
      const int kt = min(kBwdKT, n - k0);
      __syncthreads();                         // previous tile fully consumed (and sDelta / statistics visible)
      bwd_load_tile<T, D>(sK, qkv + (int64_t)k0 * C3 + C + h * D, C3, kt, 1.f);
      bwd_load_tile<T, D>(sV, qkv + (int64_t)k0 * C3 + 2 * C + h * D, C3, kt, 1.f);
      __syncthreads();
      for (int r = 0; r < QPW; ++r) {
        const int lq = r * kBwdWarps + warp;   // this warp owns local queries warp, warp+8, ...
        if (lq >= nq) break;                   // warp-uniform
        const int qi = q0 + lq;
        for (int j = lane; j < kt; j += 32) {
          const uint16_t p = pair[(int64_t)qi * ld + k0 + j];
          myPair[j] = p;
          myBias[j] = sLut[p];
        }
        float q[D];
#pragma unroll
        for (int d = 0; d < D; ++d) q[d] = sQ[lq * DP + d];
        __syncwarp();
        if (pass == 0) {
          const float m_old = sM[lq];
          float mx = m_old;
          float sv[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int j = lane + 32 * u;
            float s_ = -INFINITY;
            if (j < kt) {
              s_ = myBias[j];
#pragma unroll
              for (int d = 0; d < D; ++d) s_ = fmaf(q[d], sK[j * DP + d], s_);
            }
            sv[u] = s_;
            mx = fmaxf(mx, s_);
          }
          mx = warp_max(mx);
          float sum = 0.f;
#pragma unroll
          for (int u = 0; u < U; ++u) sum += (sv[u] == -INFINITY) ? 0.f : __expf(sv[u] - mx);
          sum = warp_sum(sum);
          if (lane == 0) {
            sL[lq] = sL[lq] * __expf(m_old - mx) + sum;
            sM[lq] = mx;
          }
        } else {
          const float lse = sM[lq] + __logf(sL[lq]);
          const float delta = sDelta[lq];
          float dO[D], dq[D];
#pragma unroll
          for (int d = 0; d < D; ++d) { dO[d] = sdO[lq * DP + d]; dq[d] = 0.f; }
#pragma unroll 2
          for (int u = 0; u < U; ++u) {
            const int j = lane + 32 * u;
            if (j < kt) {
              float s_ = myBias[j], dp = 0.f;
#pragma unroll
              for (int d = 0; d < D; ++d) {
                s_ = fmaf(q[d], sK[j * DP + d], s_);
                dp = fmaf(dO[d], sV[j * DP + d], dp);
              }
              const float p = __expf(s_ - lse);
              const float ds = p * (dp - delta);
#pragma unroll
              for (int d = 0; d < D; ++d) dq[d] = fmaf(ds, sK[j * DP + d], dq[d]);
              if (a.d_lut != nullptr) atomicAdd(sHist + myPair[j], ds);
            }
          }
#pragma unroll
          for (int d = 0; d < D; ++d) {
            const float v = warp_sum(dq[d]);
            if (lane == 0) sdQ[lq * DP + d] += v;
          }
        }
        __syncwarp();                          // bias row is overwritten by the next query
      }
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < nq * D; idx += blockDim.x) {
    const int lq = idx / D, d = idx - lq * D;
    ((T*)a.d_qkv)[(int64_t)(n0 + q0 + lq) * C3 + h * D + d] = from_float<T>(sdQ[lq * DP + d] * scale);
  }
  for (int lq = threadIdx.x; lq < nq; lq += blockDim.x) {
    a.lse[(int64_t)h * a.total_nodes + n0 + q0 + lq] = sM[lq] + __logf(sL[lq]);
    a.delta[(int64_t)h * a.total_nodes + n0 + q0 + lq] = sDelta[lq];
  }
  if (a.d_lut != nullptr)
    for (int i = threadIdx.x; i < a.lut_size; i += blockDim.x)
      if (sHist[i] != 0.f) atomicAdd(a.d_lut + (int64_t)h * a.lut_size + i, sHist[i]);
}

// Kernel B: one CTA = 64 keys of one (graph, head); query tiles are the outer loop.
template <typename T, int D>
__global__ void __launch_bounds__(256, (D <= 24) ? 2 : 1) attention_bwd_dkv_kernel(const ghn3_attention_bwd_args a) {
  extern __shared__ float bw_smem[];
  constexpr int DP = D + 1;
  constexpr int U = kBwdKT / 32;
  constexpr int KPW = kBwdBlock / kBwdWarps;
  float* sQ = bw_smem;                         // [QT][DP]   (pre-scaled by d^-1/2)
  float* sdO = sQ + kBwdKT * DP;               // [QT][DP]
  float* sKk = sdO + kBwdKT * DP;              // [64][DP] keys owned by the CTA
  float* sVk = sKk + kBwdBlock * DP;           // [64][DP]
  float* sdK = sVk + kBwdBlock * DP;           // [64][DP]
  float* sdV = sdK + kBwdBlock * DP;           // [64][DP]
  float* sBias = sdV + kBwdBlock * DP;         // [warps][QT]
  float* sLse = sBias + kBwdWarps * kBwdKT;    // [QT]
  float* sDelta = sLse + kBwdKT;               // [QT]
  float* sLut = sDelta + kBwdKT;               // [lut_size]

  const int g = blockIdx.z, h = blockIdx.y;
  const int n0 = a.node_off[g];
  const int n = a.node_off[g + 1] - n0;
  const int j0 = blockIdx.x * kBwdBlock;
  if (j0 >= n) return;
  const int nk = min(kBwdBlock, n - j0);
  const int ld = (n + 15) & ~15;
  const int C = a.hid, C3 = 3 * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const T* qkv = (const T*)a.qkv + (int64_t)n0 * C3;
  const T* dout = (const T*)a.d_out + (int64_t)n0 * C;
  const uint16_t* pair = a.pair + a.mat_off[g];
  const float scale = rsqrtf((float)D);
  // pair[i][j] = spd_ij * V + spd_ji, so the column j of the bias is row j of `pair` with its two digits swapped
  int V = 1;
  while (V * V < a.lut_size) ++V;

  pdl_launch_dependents();
  for (int i = threadIdx.x; i < a.lut_size; i += blockDim.x) sLut[i] = a.lut[(int64_t)h * a.lut_size + i];
  pdl_wait();
  bwd_load_tile<T, D>(sKk, qkv + (int64_t)j0 * C3 + C + h * D, C3, nk, 1.f);
  bwd_load_tile<T, D>(sVk, qkv + (int64_t)j0 * C3 + 2 * C + h * D, C3, nk, 1.f);
  for (int i = threadIdx.x; i < 2 * kBwdBlock * DP; i += blockDim.x) sdK[i] = 0.f;     // sdK and sdV are adjacent

  float* myBias = sBias + warp * kBwdKT;
  for (int i0 = 0; i0 < n; i0 += kBwdKT) {
    const int it = min(kBwdKT, n - i0);
    __syncthreads();
    bwd_load_tile<T, D>(sQ, qkv + (int64_t)i0 * C3 + h * D, C3, it, scale);
    bwd_load_tile<T, D>(sdO, dout + (int64_t)i0 * C + h * D, C, it, 1.f);
    for (int i = threadIdx.x; i < it; i += blockDim.x) {
      sLse[i] = a.lse[(int64_t)h * a.total_nodes + n0 + i0 + i];
      sDelta[i] = a.delta[(int64_t)h * a.total_nodes + n0 + i0 + i];
    }
    __syncthreads();
    for (int r = 0; r < KPW; ++r) {
      const int lk = r * kBwdWarps + warp;
      if (lk >= nk) break;
      const int kj = j0 + lk;
      for (int i = lane; i < it; i += 32) {
        const int pt = pair[(int64_t)kj * ld + i0 + i];            // = spd_ji * V + spd_ij
        myBias[i] = sLut[(pt % V) * V + pt / V];
      }
      float k[D], v[D], dk[D], dv[D];
#pragma unroll
      for (int d = 0; d < D; ++d) { k[d] = sKk[lk * DP + d]; v[d] = sVk[lk * DP + d]; dk[d] = 0.f; dv[d] = 0.f; }
      __syncwarp();
#pragma unroll 2
      for (int u = 0; u < U; ++u) {
        const int i = lane + 32 * u;
        if (i < it) {
          float s_ = myBias[i], dp = 0.f;
#pragma unroll
          for (int d = 0; d < D; ++d) {
            s_ = fmaf(sQ[i * DP + d], k[d], s_);
            dp = fmaf(sdO[i * DP + d], v[d], dp);
          }
          const float p = __expf(s_ - sLse[i]);
          const float ds = p * (dp - sDelta[i]);
#pragma unroll
          for (int d = 0; d < D; ++d) {
            dv[d] = fmaf(p, sdO[i * DP + d], dv[d]);
            dk[d] = fmaf(ds, sQ[i * DP + d], dk[d]);               // sQ already carries d^-1/2
          }
        }
      }
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float vk = warp_sum(dk[d]);
        const float vv = warp_sum(dv[d]);
        if (lane == 0) {
          sdK[lk * DP + d] += vk;
          sdV[lk * DP + d] += vv;
        }
      }
      __syncwarp();
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < nk * D; idx += blockDim.x) {
    const int lk = idx / D, d = idx - lk * D;
    T* dst = (T*)a.d_qkv + (int64_t)(n0 + j0 + lk) * C3 + h * D + d;
    dst[C] = from_float<T>(sdK[lk * DP + d]);
    dst[2 * C] = from_float<T>(sdV[lk * DP + d]);
  }
}

template <typename T, int D>
static int launch_attention_bwd(const ghn3_attention_bwd_args* a, cudaStream_t stream) {
  constexpr int DP = D + 1;
  const size_t smem_a = sizeof(float) * (2 * kBwdKT * DP + 3 * kBwdBlock * DP + 3 * kBwdBlock + kBwdWarps * kBwdKT +
                                         2 * a->lut_size) + sizeof(uint16_t) * kBwdWarps * kBwdKT;
  const size_t smem_b = sizeof(float) * (2 * kBwdKT * DP + 4 * kBwdBlock * DP + kBwdWarps * kBwdKT + 2 * kBwdKT +
                                         a->lut_size);
  GHN3_REQUIRE(smem_a <= 220 * 1024 && smem_b <= 220 * 1024, "ghn3_attention_bwd: look-up table too large");
  GHN3_CUDA(cudaFuncSetAttribute(attention_bwd_dq_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
  GHN3_CUDA(cudaFuncSetAttribute(attention_bwd_dkv_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
  const dim3 grid((unsigned)ceil_div(a->max_nodes, kBwdBlock), (unsigned)a->heads, (unsigned)a->n_graphs);
  GHN3_CUDA(launch_pdl(attention_bwd_dq_kernel<T, D>, grid, dim3(256), smem_a, stream, *a));
  GHN3_LAUNCH_CHECK("attention_bwd_dq_kernel");
  GHN3_CUDA(launch_pdl(attention_bwd_dkv_kernel<T, D>, grid, dim3(256), smem_b, stream, *a));
  GHN3_LAUNCH_CHECK("attention_bwd_dkv_kernel");
  return GHN3_OK;
}

int attention_bwd_mma_impl(const ghn3_attention_bwd_args* a, cudaStream_t stream);

int attention_bwd_impl(const ghn3_attention_bwd_args* a, cudaStream_t stream) {
  GHN3_REQUIRE(a != nullptr, "ghn3_attention_bwd: null args");
  GHN3_REQUIRE(a->heads > 0 && a->hid % a->heads == 0, "ghn3_attention_bwd: hid must be divisible by heads");
  GHN3_REQUIRE(a->lse != nullptr && a->delta != nullptr, "ghn3_attention_bwd: lse / delta workspaces are required");
  if (a->n_graphs <= 0 || a->max_nodes <= 0) return GHN3_OK;
  if (a->dtype == GHN3_BF16 && a->fwd_lse2 != nullptr) return attention_bwd_mma_impl(a, stream);
  const int D = a->hid / a->heads;
  const bool bf = a->dtype == GHN3_BF16;
#define GHN3_ATTN_BWD_CASE(DV) \
  if (D == DV) return bf ? launch_attention_bwd<__nv_bfloat16, DV>(a, stream) : launch_attention_bwd<float, DV>(a, stream);
  GHN3_ATTN_BWD_CASE(4)
  GHN3_ATTN_BWD_CASE(8)
  GHN3_ATTN_BWD_CASE(16)
  GHN3_ATTN_BWD_CASE(24)
  GHN3_ATTN_BWD_CASE(32)
#undef GHN3_ATTN_BWD_CASE
  set_error("ghn3_attention_bwd: head dim %d is not supported (4, 8, 16, 24, 32)", D);
  return GHN3_ERR_UNSUPPORTED;
}

// ---------------------------------------------------------------------------------------------------------------
// Adjoint of scatter_kernel: every target element sends grad * d(finish)/dv back to the source element it was read
// from (tiled copies add up). Same descriptor table and index arithmetic as the forward kernel.
__device__ __forceinline__ uint32_t fdiv_b(uint32_t n, uint32_t mul, uint32_t sh) {
  return mul ? (__umulhi(n, mul) >> sh) : n;
}

__global__ void __launch_bounds__(256) scatter_bwd_kernel(const ghn3_scatter_desc* __restrict__ descs,
                                                          const int32_t* __restrict__ chunk_desc,
                                                          const float* const* __restrict__ grads,
                                                          float* const* __restrict__ d_src) {
  const int di = __ldg(chunk_desc + blockIdx.x);
  const float* grad = grads[di];
  float* ds = d_src[di];
  if (grad == nullptr || ds == nullptr) return;
  const ghn3_scatter_desc d = descs[di];
  const uint32_t base = (uint32_t)(((int64_t)blockIdx.x - d.chunk0) * GHN3_SCATTER_CHUNK);
  const uint32_t end = (uint32_t)min((int64_t)base + GHN3_SCATTER_CHUNK, d.numel);
  const int mode = d.mode;
  for (uint32_t e = base + threadIdx.x; e < end; e += 256) {
    uint32_t r = fdiv_b(e, d.m_t3, d.s_t3);
    const int x = (int)(e - r * d.t3);
    uint32_t r2 = fdiv_b(r, d.m_t2, d.s_t2);
    const int y = (int)(r - r2 * d.t2);
    const uint32_t a = fdiv_b(r2, d.m_t1, d.s_t1);
    const uint32_t b = r2 - a * d.t1;
    const uint32_t am = a - fdiv_b(a, d.m_so, d.s_so) * d.so;
    const uint32_t bm = b - fdiv_b(b, d.m_si, d.s_si) * d.si;
    const int64_t col = (int64_t)am * d.ca + bm;
    const float gval = grad[e];
    if (mode == 3) {
      const float sy = fmaxf(((float)y + 0.5f) * ((float)d.kh_src / (float)d.t2) - 0.5f, 0.f);
      const float sx = fmaxf(((float)x + 0.5f) * ((float)d.kw_src / (float)d.t3) - 0.5f, 0.f);
      const int y0 = (int)sy, x0 = (int)sx;
      const int y1 = min(y0 + 1, d.kh_src - 1), x1 = min(x0 + 1, d.kw_src - 1);
      const float ly = sy - (float)y0, lx = sx - (float)x0;
      const int64_t rb = (int64_t)a * d.ra;
      const float gs = gval * d.scale;
      atomicAdd(ds + (rb + y0 * d.kw_src + x0) * d.ld + col, gs * (1.f - ly) * (1.f - lx));
      atomicAdd(ds + (rb + y0 * d.kw_src + x1) * d.ld + col, gs * (1.f - ly) * lx);
      atomicAdd(ds + (rb + y1 * d.kw_src + x0) * d.ld + col, gs * ly * (1.f - lx));
      atomicAdd(ds + (rb + y1 * d.kw_src + x1) * d.ld + col, gs * ly * lx);
    } else {
      const int64_t row = (int64_t)a * d.ra + (int64_t)(y + d.cy) * d.kw_src + (x + d.cx);
      const int64_t si = row * d.ld + col;
      float dv;
      if (mode == 1) {
        const float s = 1.0f / (1.0f + expf(-(0.5f * d.src[si])));
        dv = s * (1.f - s);                       // d/dv 2*sigmoid(v/2)
      } else if (mode == 2) {
        const float t = tanhf(0.2f * d.src[si]);
        dv = 0.2f * (1.f - t * t);
      } else {
        dv = d.scale;
      }
      atomicAdd(ds + si, gval * dv);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Adjoint of node_features_kernel: scatter-add of dx rows into the six embedding tables
__global__ void __launch_bounds__(256) node_features_bwd_kernel(const ghn3_node_features_bwd_args a) {
  const int node = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (node >= a.total_nodes) return;
  const int C = a.hid, Q = C >> 2;
  const int op = a.op[node];
  const int4 si = *(const int4*)(a.shape_idx + 4 * (int64_t)node);
  const int din = a.deg_in[node], dout = a.deg_out[node], d0 = a.dist0[node];
  for (int c = lane; c < C; c += 32) {
    const float g = a.dx[(int64_t)node * C + c];
    const int q = c / Q, cc = c - q * Q;
    const int sidx = q == 0 ? si.x : (q == 1 ? si.y : (q == 2 ? si.z : si.w));
    float* stab = q < 2 ? a.d_embed_ch : a.d_embed_sp;
    atomicAdd(a.d_embed_op + (int64_t)op * C + c, g);
    atomicAdd(stab + (int64_t)sidx * Q + cc, g);
    atomicAdd(a.d_cent_in + (int64_t)din * C + c, g);
    atomicAdd(a.d_cent_out + (int64_t)dout * C + c, g);
    atomicAdd(a.d_dist_embed + (int64_t)d0 * C + c, g);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Adjoint of the edge-bias look-up table (edge_lut_stage1/2). P = stage-1 projections [2][V][C] (recomputed by the
// caller with the forward kernel), dP zero-initialised by the entry point.
__global__ void __launch_bounds__(256) edge_lut_bwd_stage2(int C, int V, int H, const float* __restrict__ P,
                                                           const float* __restrict__ b1, const float* __restrict__ W2,
                                                           const float* __restrict__ dlut, float* __restrict__ dP,
                                                           float* __restrict__ db1, float* __restrict__ dW2,
                                                           float* __restrict__ db2) {
  extern __shared__ float lb_smem[];
  float* sW2 = lb_smem;              // [H][C] gradient accumulator
  float* sB1 = sW2 + H * C;          // [C]
  float* sB2 = sB1 + C;              // [H]
  for (int i = threadIdx.x; i < H * C + C + H; i += blockDim.x) lb_smem[i] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int ab = blockIdx.x * 8 + warp; ab < V * V; ab += gridDim.x * 8) {
    const int av = ab / V, bv = ab % V;
    const float* pa = P + (int64_t)av * C;
    const float* pb = P + (int64_t)(V + bv) * C;
    const float dl_lane = lane < H ? dlut[(int64_t)lane * V * V + ab] : 0.f;     // H <= 32
    if (lane < H && dl_lane != 0.f) atomicAdd(sB2 + lane, dl_lane);
    for (int c = lane; c < C; c += 32) {
      const float t = fmaxf((pa[c] + pb[c]) + b1[c], 0.f);
      float dt = 0.f;
      for (int h = 0; h < H; ++h) {
        const float dl = __shfl_sync(0xffffffffu, dl_lane, h);
        dt = fmaf(dl, W2[(int64_t)h * C + c], dt);
        if (t > 0.f) atomicAdd(sW2 + h * C + c, dl * t);
      }
      if (t > 0.f && dt != 0.f) {
        atomicAdd(sB1 + c, dt);
        atomicAdd(dP + (int64_t)av * C + c, dt);
        atomicAdd(dP + (int64_t)(V + bv) * C + c, dt);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H * C; i += blockDim.x)
    if (sW2[i] != 0.f) atomicAdd(dW2 + i, sW2[i]);
  for (int i = threadIdx.x; i < C; i += blockDim.x)
    if (sB1[i] != 0.f) atomicAdd(db1 + i, sB1[i]);
  for (int i = threadIdx.x; i < H; i += blockDim.x)
    if (sB2[i] != 0.f) atomicAdd(db2 + i, sB2[i]);
}

// dW1[c][side*C + k] += sum_v dP[side][v][c] * E[v+2][k]
__global__ void __launch_bounds__(256) edge_lut_bwd_w1(int C, int V, const float* __restrict__ E,
                                                       const float* __restrict__ dP, float* __restrict__ dW1) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= (int64_t)C * 2 * C) return;
  const int c = (int)(o / (2 * C)), sk = (int)(o % (2 * C));
  const int side = sk / C, k = sk - side * C;
  float acc = 0.f;
  for (int v = 0; v < V; ++v) acc = fmaf(dP[((int64_t)side * V + v) * C + c], E[(int64_t)(v + 2) * C + k], acc);
  dW1[o] += acc;
}

// dE[v+2][k] += sum_side sum_c dP[side][v][c] * W1[c][side*C + k]
__global__ void __launch_bounds__(256) edge_lut_bwd_embed(int C, int V, const float* __restrict__ W1,
                                                          const float* __restrict__ dP, float* __restrict__ dE) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= V * C) return;
  const int v = o / C, k = o - v * C;
  float acc = 0.f;
  for (int side = 0; side < 2; ++side)
    for (int c = 0; c < C; ++c)
      acc = fmaf(dP[((int64_t)side * V + v) * C + c], W1[(int64_t)c * 2 * C + side * C + k], acc);
  dE[(int64_t)(v + 2) * C + k] += acc;
}


// ---------------------------------------------------------------------------------------------------------------
// Decoder fc stage backward (forward: one GEMM problem per decoder-grid position, reference ghn3/nn.py:738-745).
// Works on the ORIGINAL fp32 parameter layout fc.weight[(j*S*S + pos)][c], fc.bias[j*S*S + pos]; the (node, position)
// rows of dh0 are found through the forward row map.
//   d_w[(j, pos)][c] += sum_k dh0[row(k)][j] * x[a_row0 + k][c];  d_b[(j, pos)] += sum_k dh0[row(k)][j]
__global__ void __launch_bounds__(256) fc_wgrad_kernel(const ghn3_fc_bwd_args a) {
  __shared__ float sG[16][65];
  __shared__ float sX[16][65];
  const ghn3_gemm_problem p = a.problems[blockIdx.z];
  const int C = a.hid, J = a.n_out, SS = a.grid_positions;
  const int pos = p.b_row0 / J;
  const int j0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const int tj = threadIdx.x >> 4, tc = threadIdx.x & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;
  for (int k0 = 0; k0 < p.m; k0 += 16) {
    const int kt = min(16, p.m - k0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < 16 * 64; idx += 256) {
      const int kk = idx >> 6, e = idx & 63;
      float g = 0.f, x = 0.f;
      if (kk < kt) {
        const int64_t row = a.rowmap[p.d_off + k0 + kk];
        if (j0 + e < J) g = ld_f(a.dh0, row * J + j0 + e, a.dtype);
        if (c0 + e < C) x = ld_f(a.dec_in, (int64_t)(p.a_row0 + k0 + kk) * C + c0 + e, a.dtype);
      }
      sG[kk][e] = g;
      sX[kk][e] = x;
    }
    __syncthreads();
    for (int kk = 0; kk < kt; ++kk) {
      float g[4], x[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { g[i] = sG[kk][tj * 4 + i]; x[i] = sX[kk][tc * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(g[i], x[j], acc[i][j]);
      if (blockIdx.y == 0 && threadIdx.x < 64) bsum += sG[kk][threadIdx.x];
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = j0 + tj * 4 + i;
    if (j >= J) continue;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int c = c0 + tc * 4 + jj;
      if (c < C && acc[i][jj] != 0.f) atomicAdd(a.d_fc_w + ((int64_t)j * SS + pos) * C + c, acc[i][jj]);
    }
  }
  if (blockIdx.y == 0 && threadIdx.x < 64 && j0 + threadIdx.x < J && bsum != 0.f)
    atomicAdd(a.d_fc_b + (int64_t)(j0 + threadIdx.x) * SS + pos, bsum);
}

//   d_x[a_row0 + k][c] += sum_j dh0[row(k)][j] * w[(j, pos)][c]
__global__ void __launch_bounds__(256) fc_dgrad_kernel(const ghn3_fc_bwd_args a) {
  __shared__ float sW[32][64];
  __shared__ float sG[16][33];
  const ghn3_gemm_problem p = a.problems[blockIdx.z];
  const int k0 = blockIdx.y * 16;
  if (k0 >= p.m) return;
  const int kt = min(16, p.m - k0);
  const int C = a.hid, J = a.n_out, SS = a.grid_positions;
  const int pos = p.b_row0 / J;
  const int c0 = blockIdx.x * 64;
  const int tr = threadIdx.x >> 4, tc = threadIdx.x & 15;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j0 = 0; j0 < J; j0 += 32) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < 32 * 64; idx += 256) {
      const int jj = idx >> 6, e = idx & 63;
      float w = 0.f;
      if (j0 + jj < J && c0 + e < C) w = a.fc_w[((int64_t)(j0 + jj) * SS + pos) * C + c0 + e];
      sW[jj][e] = w;
    }
    for (int idx = threadIdx.x; idx < 16 * 32; idx += 256) {
      const int kk = idx >> 5, jj = idx & 31;
      float g = 0.f;
      if (kk < kt && j0 + jj < J) g = ld_f(a.dh0, (int64_t)a.rowmap[p.d_off + k0 + kk] * J + j0 + jj, a.dtype);
      sG[kk][jj] = g;
    }
    __syncthreads();
#pragma unroll 8
    for (int jj = 0; jj < 32; ++jj) {
      const float g = sG[tr][jj];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(g, sW[jj][tc * 4 + i], acc[i]);
    }
  }
  if (tr < kt) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + tc * 4 + i;
      if (c < C && acc[i] != 0.f) atomicAdd(a.d_dec_in + (int64_t)(p.a_row0 + k0 + tr) * C + c, acc[i]);
    }
  }
}

// Adjoint of relu_transpose_kernel: d_src[z*src_bs + r*ld + c] += (src > 0) * d_rt[(z*cols + c)*rows + r]
__global__ void __launch_bounds__(256) relu_transpose_bwd_kernel(const ghn3_relu_transpose_bwd_args a) {
  const int64_t total = (int64_t)a.batch * a.rows * a.cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % a.cols);
    const int r = (int)((i / a.cols) % a.rows);
    const int64_t z = i / ((int64_t)a.cols * a.rows);
    const int64_t si = z * a.src_bs + (int64_t)r * a.ld + c;
    if (a.src[si] > 0.f) a.d_src[si] += a.d_rt[(z * a.cols + c) * a.rows + r];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Decoder conv.2 backward, operand preparation: the compact per-class gradients dwout_c [rows_c][o'*i'] are expanded
// into the full max_shape column space, x[row0 + r][col(c)] and xt[col(c)][row0 + r] with col(c) = (c / i') * ms1 +
// c % i' (zeros elsewhere, set by the caller), so that ONE dgrad GEMM and ONE wgrad GEMM cover every column class;
// the zeros cost tensor-core flops, not HBM passes over the 1.8 GB weight gradient. Also accumulates the bias
// gradient d_bias[col(c)] += sum_r dwout_c[r][c].
__global__ void __launch_bounds__(256) expand_kernel(const ghn3_expand_args a) {
  __shared__ float tile[32][33];
  __shared__ int s_seg;
  if (threadIdx.x == 0) {
    int lo = 0, hi = a.n_segs;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (a.segs[mid].tile0 <= (int)blockIdx.x) lo = mid; else hi = mid;
    }
    s_seg = lo;
  }
  __syncthreads();
  const ghn3_expand_seg sg = a.segs[s_seg];
  const int lt = blockIdx.x - sg.tile0;
  const int r0 = (lt / sg.tiles_c) * 32, c0 = (lt % sg.tiles_c) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* src = a.src + sg.src_off;
  auto col = [&](int c) -> int64_t {
    return sg.group > 0 ? (int64_t)(c / sg.group) * a.group_stride + c % sg.group : c;
  };
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    float v = 0.f;
    if (r < sg.rows && c < sg.ld) {
      v = src[(int64_t)r * sg.ld + c];
      st_f(a.x, (int64_t)(sg.row0 + r) * a.ld_x + col(c), a.dtype, v);
    }
    tile[ty + 8 * i][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, r = r0 + tx;
    if (c < sg.ld && r < sg.rows) st_f(a.xt, col(c) * a.ld_xt + sg.row0 + r, a.dtype, tile[tx][ty + 8 * i]);
  }
  if (a.d_bias != nullptr && ty == 0 && c0 + tx < sg.ld) {
    float t = 0.f;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) t += tile[r][tx];
    if (t != 0.f) atomicAdd(a.d_bias + col(c0 + tx), t);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Optimizer step of the training path: global-norm gradient clipping + decoupled-weight-decay Adam (AdamW) in ONE
// pass over the flat gradient buffer (reference: nn.utils.clip_grad_norm_ + torch.optim.AdamW.step, trainer.py:
// 343-379). Gradients and both moment buffers are flat (same layout); parameters stay separate tensors and are
// reached through a pointer table. The clipping coefficient is computed on the device from the squared norm.
__global__ void __launch_bounds__(256) sumsq_flat_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ out) {
  float acc = 0.f;
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = ((const float4*)g)[i];
    acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = g[(n4 << 2) + threadIdx.x];
    acc += v * v;
  }
  double d = (double)warp_sum(acc);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < 8; ++i) t += part[i];
    atomicAdd(out, t);
  }
}

__global__ void __launch_bounds__(256) adamw_kernel(const ghn3_adamw_args a) {
  const int64_t chunk = (int64_t)blockIdx.x + a.chunk_begin;
  const int t = a.chunk_tensor[chunk];
  const int64_t off = a.offsets[t];
  const int64_t n = a.numels[t];
  int64_t c0 = (chunk - a.chunk0[t]) * GHN3_ADAMW_CHUNK;
  int64_t c1 = min(c0 + (int64_t)GHN3_ADAMW_CHUNK, n);
  if (a.range_hi > 0) {                    // this rank's shard of the flat buffers (bounds are multiples of 4)
    c0 = max(c0, a.range_lo - off);
    c1 = min(c1, a.range_hi - off);
    if (c0 >= c1) return;
  }
  float* __restrict__ p = a.params[t];
  const float* __restrict__ g = a.grads + off;
  float* __restrict__ m = a.exp_avg + off;
  float* __restrict__ v = a.exp_avg_sq + off;
  float coef = 1.f;
  if (a.skipped != nullptr) {
    // non-finite guard: fminf(1, NaN) = 1 and inf * 0 = NaN would otherwise poison the parameters and both moments
    const double ss = *a.sumsq;
    const bool bad = !(ss == ss) || ss > 1e300 || (a.loss != nullptr && !isfinite(*a.loss));
    if (bad) {
      if (blockIdx.x == 0 && threadIdx.x == 0 && a.sumsq_ready != 2) atomicAdd(a.skipped, 1);   // once per step
      return;
    }
  }
  if (a.max_norm > 0.f) {
    const float norm = sqrtf((float)*a.sumsq);
    coef = fminf(1.f, a.max_norm / (norm + 1e-6f));        // nn.utils.clip_grad_norm_
  }
  const float b1 = a.beta1, b2 = a.beta2;
  const float decay = 1.f - a.lr * a.weight_decay;
  const float step_size = a.lr / a.bias_correction1;
  const float inv_bc2_sqrt = rsqrtf(a.bias_correction2);
  const bool vec = (((uintptr_t)p) & 15) == 0;             // flat buffers are 16-byte aligned per tensor
  if (vec) {
    for (int64_t i = c0 + threadIdx.x * 4; i + 4 <= c1; i += 1024) {
      float4 pv = *(float4*)(p + i);
      const float4 gv = *(const float4*)(g + i);
      float4 mv = *(float4*)(m + i), vv = *(float4*)(v + i);
      float* pp = &pv.x; const float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gr = gp[k] * coef;
        mp[k] = b1 * mp[k] + (1.f - b1) * gr;
        vp[k] = b2 * vp[k] + (1.f - b2) * gr * gr;
        const float denom = sqrtf(vp[k]) * inv_bc2_sqrt + a.eps;
        pp[k] = pp[k] * decay - step_size * (mp[k] / denom);
      }
      *(float4*)(p + i) = pv;
      *(float4*)(m + i) = mv;
      *(float4*)(v + i) = vv;
    }
  }
  // tail (or everything, for an unaligned parameter)
  const int64_t t0 = vec ? c0 + ((c1 - c0) & ~(int64_t)3) : c0;
  for (int64_t i = t0 + threadIdx.x; i < c1; i += 256) {
    const float gr = g[i] * coef;
    const float mm = b1 * m[i] + (1.f - b1) * gr;
    const float vv = b2 * v[i] + (1.f - b2) * gr * gr;
    m[i] = mm;
    v[i] = vv;
    p[i] = p[i] * decay - step_size * (mm / (sqrtf(vv) * inv_bc2_sqrt + a.eps));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Frobenius-norm regulariser of the predicted parameters (ghn3_segnorm): one pass for the per-tensor sums of squares,
// one tiny kernel for the total, one pass for the gradient -- instead of one torch.norm (+ its backward) per tensor.
__global__ void __launch_bounds__(256) segnorm_sumsq_kernel(const ghn3_segnorm_args a) {
  const int s = a.chunk_seg[blockIdx.x];
  const int64_t n = a.seg_numel[s];
  const int64_t c0 = ((int64_t)blockIdx.x - a.chunk0[s]) * GHN3_ADAMW_CHUNK;
  const int64_t c1 = min(c0 + (int64_t)GHN3_ADAMW_CHUNK, n);
  const float* __restrict__ v = a.src + a.seg_off[s];
  float acc = 0.f;
  for (int64_t i = c0 + threadIdx.x; i < c1; i += 256) acc = fmaf(v[i], v[i], acc);
  double d = (double)warp_sum(acc);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < 8; ++i) t += part[i];
    atomicAdd(a.sumsq + s, t);
  }
}

__global__ void __launch_bounds__(256) segnorm_total_kernel(const ghn3_segnorm_args a) {
  double acc = 0;
  for (int s = threadIdx.x; s < a.n_segs; s += 256) acc += sqrt(a.sumsq[s]);
  __shared__ double part[256];
  part[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *a.total = (float)(part[0] * (double)a.coef);
}

__global__ void __launch_bounds__(256) segnorm_grad_kernel(const ghn3_segnorm_args a) {
  const int s = a.chunk_seg[blockIdx.x];
  const int64_t n = a.seg_numel[s];
  const int64_t c0 = ((int64_t)blockIdx.x - a.chunk0[s]) * GHN3_ADAMW_CHUNK;
  const int64_t c1 = min(c0 + (int64_t)GHN3_ADAMW_CHUNK, n);
  const double ss = a.sumsq[s];
  if (!(ss > 0)) return;                                  // d||p|| / dp is taken as 0 at p = 0
  const float k = (a.gscale ? *a.gscale : 1.f) * a.coef * (float)(1.0 / sqrt(ss));
  const float* __restrict__ v = a.src + a.seg_off[s];
  float* __restrict__ g = a.grad + a.seg_off[s];
  for (int64_t i = c0 + threadIdx.x; i < c1; i += 256) g[i] = fmaf(k, v[i], g[i]);
}

}  // namespace ghn3

using namespace ghn3;

extern "C" int ghn3_segnorm(const ghn3_segnorm_args* a, ghn3_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHN3_REQUIRE(a != nullptr, "ghn3_segnorm: null args");
  if (a->n_chunks <= 0 || a->n_segs <= 0) return GHN3_OK;
  GHN3_REQUIRE(a->src && a->seg_off && a->seg_numel && a->chunk0 && a->chunk_seg && a->sumsq, "ghn3_segnorm: null pointer");
  if (a->mode == 0) {
    GHN3_REQUIRE(a->total != nullptr, "ghn3_segnorm: mode 0 needs `total`");
    GHN3_CUDA(cudaMemsetAsync(a->sumsq, 0, sizeof(double) * (size_t)a->n_segs, stream));
    segnorm_sumsq_kernel<<<(unsigned)a->n_chunks, 256, 0, stream>>>(*a);
    GHN3_LAUNCH_CHECK("segnorm_sumsq_kernel");
    segnorm_total_kernel<<<1, 256, 0, stream>>>(*a);
    GHN3_LAUNCH_CHECK("segnorm_total_kernel");
  } else {
    GHN3_REQUIRE(a->grad != nullptr, "ghn3_segnorm: mode 1 needs `grad`");
    segnorm_grad_kernel<<<(unsigned)a->n_chunks, 256, 0, stream>>>(*a);
    GHN3_LAUNCH_CHECK("segnorm_grad_kernel");
  }
  return GHN3_OK;
}

extern "C" int ghn3_transpose(const ghn3_transpose_args* a, ghn3_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHN3_REQUIRE(a != nullptr && a->src != nullptr && a->dst != nullptr, "ghn3_transpose: null args");
  if (a->rows <= 0 || a->cols <= 0) return GHN3_OK;
  GHN3_REQUIRE(a->ld_dst >= a->rows && a->ld_src >= a->cols, "ghn3_transpose: leading dimensions too small");
  const dim3 grid((unsigned)ceil_div(a->cols, 32), (unsigned)ceil_div(a->rows, 32));
  GHN3_REQUIRE(grid.y < 65536, "ghn3_transpose: too many rows");
  GHN3_CUDA(launch_pdl(transpose_kernel, grid, dim3(256), 0, stream, *a));
  GHN3_LAUNCH_CHECK("transpose_kernel");
  return GHN3_OK;
}

extern "C" int ghn3_elementwise(const ghn3_elementwise_args* a, ghn3_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHN3_REQUIRE(a != nullptr && a->a != nullptr && a->out != nullptr, "ghn3_elementwise: null args");
  GHN3_REQUIRE(a->op >= GHN3_EW_COPY && a->op <= GHN3_EW_ADD, "ghn3_elementwise: bad op");
  GHN3_REQUIRE(a->op == GHN3_EW_COPY || a->op == GHN3_EW_GELU || a->b != nullptr, "ghn3_elementwise: second operand missing");
  if (a->n <= 0) return GHN3_OK;
  const int vec = (a->n % 4 == 0) && ((((uintptr_t)a->a) | ((uintptr_t)a->out) | ((uintptr_t)a->b)) & 15) == 0;
  const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(vec ? a->n / 4 : a->n, 256), (int64_t)num_sms() * 16);
  GHN3_CUDA(launch_pdl(elementwise_kernel, dim3(blocks), dim3(256), 0, stream, *a, vec));
  GHN3_LAUNCH_CHECK("elementwise_kernel");
  return GHN3_OK;
}

extern "C" int ghn3_colsum(const ghn3_colsum_args* a, ghn3_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHN3_REQUIRE(a != nullptr && a->src != nullptr && a->dst != nullptr, "ghn3_colsum: null args");
  if (a->rows <= 0 || a->cols <= 0) return GHN3_OK;
  const dim3 grid((unsigned)ceil_div(a->cols, 32), (unsigned)ceil_div(a->rows, 256));
  GHN3_REQUIRE(grid.y < 65536, "ghn3_colsum: too many rows");
  GHN3_CUDA(launch_pdl(colsum_kernel, grid, dim3(256), 0, stream, *a));
  GHN3_LAUNCH_CHECK("colsum_kernel");
  return GHN3_OK;
}

extern "C" int ghn3_layernorm_bwd(const ghn3_layernorm_bwd_args* a, ghn3_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHN3_REQUIRE(a != nullptr, "ghn3_layernorm_bwd: null args");
  GHN3_REQUIRE(a->hid > 0 && a->hid <= 1024, "ghn3_layernorm_bwd: hid must be <= 1024");
  GHN3_REQUIRE(a->dy && a->dx && (a->gamma == nullptr || (a->x && a->dgamma && a->dbeta)),
               "ghn3_layernorm_bwd: null pointer");
  if (a->rows <= 0) return GHN3_OK;
  const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(a->rows, 8), (int64_t)num_sms() * 3);
  const size_t smem = sizeof(float) * 2 * a->hid;
  const int t = (a->hid + 31) / 32;
  if (t <= 4) GHN3_CUDA(launch_pdl(layernorm_bwd_kernel<4>, dim3(blocks), dim3(256), smem, stream, *a));
  else if (t <= 8) GHN3_CUDA(launch_pdl(layernorm_bwd_kernel<8>, dim3(blocks), dim3(256), smem, stream, *a));
  else if (t <= 12) GHN3_CUDA(launch_pdl(layernorm_bwd_kernel<12>, dim3(blocks), dim3(256), smem, stream, *a));
  else if (t <= 16) GHN3_CUDA(launch_pdl(layernorm_bwd_kernel<16>, dim3(blocks), dim3(256), smem, stream, *a));
  else GHN3_CUDA(launch_pdl(layernorm_bwd_kernel<32>, dim3(blocks), dim3(256), smem, stream, *a));
  GHN3_LAUNCH_CHECK("layernorm_bwd_kernel");
  return GHN3_OK;
}

extern "C" int ghn3_attention_bwd(const ghn3_attention_bwd_args* a, ghn3_stream_t stream) {
  return attention_bwd_impl(a, (cudaStream_t)stream);
}

extern "C" int ghn3_scatter_bwd(const ghn3_scatter_bwd_args* a, ghn3_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHN3_REQUIRE(a != nullptr, "ghn3_scatter_bwd: null args");
  if (a->n_descs <= 0 || a->n_chunks <= 0) return GHN3_OK;
  GHN3_REQUIRE(a->chunk_desc != nullptr && a->grads != nullptr && a->d_src != nullptr && a->descs != nullptr,
               "ghn3_scatter_bwd: descriptor, chunk, gradient and source-gradient tables are required");
  GHN3_REQUIRE(a->n_chunks < (int64_t)2147483647, "ghn3_scatter_bwd: too many chunks");
  scatter_bwd_kernel<<<(unsigned)a->n_chunks, 256, 0, stream>>>(a->descs, a->chunk_desc, a->grads, a->d_src);
  GHN3_LAUNCH_CHECK("scatter_bwd_kernel");
  return GHN3_OK;
}

extern "C" int ghn3_node_features_bwd(const ghn3_node_features_bwd_args* a, ghn3_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHN3_REQUIRE(a != nullptr && a->dx != nullptr, "ghn3_node_features_bwd: null args");
  GHN3_REQUIRE(a->hid > 0 && a->hid % 4 == 0, "ghn3_node_features_bwd: hid must be a multiple of 4");
  if (a->total_nodes <= 0) return GHN3_OK;
  node_features_bwd_kernel<<<(unsigned)ceil_div(a->total_nodes, 8), 256, 0, stream>>>(*a);
  GHN3_LAUNCH_CHECK("node_features_bwd_kernel");
  return GHN3_OK;
}

extern "C" int ghn3_edge_lut_bwd(const ghn3_edge_lut_bwd_args* a, ghn3_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHN3_REQUIRE(a != nullptr && a->d_lut != nullptr && a->workspace != nullptr, "ghn3_edge_lut_bwd: null args");
  GHN3_REQUIRE(a->heads > 0 && a->heads <= 32, "ghn3_edge_lut_bwd: at most 32 heads");
  const int C = a->hid, V = a->vmax + 1, H = a->heads;
  // workspace: P [2][V][C] | dP [2][V][C]
  float* P = a->workspace;
  float* dP = P + (int64_t)2 * V * C;
  ghn3_edge_lut_args f = {};
  f.hid = C; f.heads = H; f.vmax = a->vmax;
  f.edge_embed = a->edge_embed; f.w1 = a->w1; f.b1 = a->b1; f.w2 = a->w2; f.b2 = nullptr; f.workspace = P; f.lut = nullptr;
  int rc = ghn3_edge_lut(&f, stream_);          // lut == NULL: stage 1 only
  if (rc != GHN3_OK) return rc;
  GHN3_CUDA(cudaMemsetAsync(dP, 0, sizeof(float) * 2 * V * C, stream));
  const size_t smem = sizeof(float) * ((size_t)H * C + C + H);
  GHN3_REQUIRE(smem <= 160 * 1024, "ghn3_edge_lut_bwd: heads * hid too large");
  GHN3_CUDA(cudaFuncSetAttribute(edge_lut_bwd_stage2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div((int64_t)V * V, 8), num_sms());
  edge_lut_bwd_stage2<<<blocks, 256, smem, stream>>>(C, V, H, P, a->b1, a->w2, a->d_lut, dP, a->d_b1, a->d_w2, a->d_b2);
  GHN3_LAUNCH_CHECK("edge_lut_bwd_stage2");
  edge_lut_bwd_w1<<<(unsigned)ceil_div((int64_t)C * 2 * C, 256), 256, 0, stream>>>(C, V, a->edge_embed, dP, a->d_w1);
  GHN3_LAUNCH_CHECK("edge_lut_bwd_w1");
  edge_lut_bwd_embed<<<(unsigned)ceil_div((int64_t)V * C, 256), 256, 0, stream>>>(C, V, a->w1, dP, a->d_edge_embed);
  GHN3_LAUNCH_CHECK("edge_lut_bwd_embed");
  return GHN3_OK;
}

extern "C" int ghn3_fc_bwd(const ghn3_fc_bwd_args* a, ghn3_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHN3_REQUIRE(a != nullptr, "ghn3_fc_bwd: null args");
  if (a->n_problems <= 0) return GHN3_OK;
  GHN3_REQUIRE(a->problems && a->rowmap && a->dh0 && a->dec_in && a->fc_w && a->d_fc_w && a->d_fc_b && a->d_dec_in,
               "ghn3_fc_bwd: null pointer");
  GHN3_REQUIRE(a->n_problems < 65536 && a->max_m > 0, "ghn3_fc_bwd: too many problems / max_m missing");
  const dim3 gw((unsigned)ceil_div(a->n_out, 64), (unsigned)ceil_div(a->hid, 64), (unsigned)a->n_problems);
  fc_wgrad_kernel<<<gw, 256, 0, stream>>>(*a);
  GHN3_LAUNCH_CHECK("fc_wgrad_kernel");
  const dim3 gd((unsigned)ceil_div(a->hid, 64), (unsigned)ceil_div(a->max_m, 16), (unsigned)a->n_problems);
  fc_dgrad_kernel<<<gd, 256, 0, stream>>>(*a);
  GHN3_LAUNCH_CHECK("fc_dgrad_kernel");
  return GHN3_OK;
}

extern "C" int ghn3_relu_transpose_bwd(const ghn3_relu_transpose_bwd_args* a, ghn3_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHN3_REQUIRE(a != nullptr && a->src && a->d_src && a->d_rt, "ghn3_relu_transpose_bwd: null args");
  const int64_t total = (int64_t)a->batch * a->rows * a->cols;
  if (total <= 0) return GHN3_OK;
  const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(total, 256), (int64_t)num_sms() * 16);
  relu_transpose_bwd_kernel<<<blocks, 256, 0, stream>>>(*a);
  GHN3_LAUNCH_CHECK("relu_transpose_bwd_kernel");
  return GHN3_OK;
}

extern "C" int ghn3_expand_cols(const ghn3_expand_args* a, ghn3_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHN3_REQUIRE(a != nullptr, "ghn3_expand_cols: null args");
  if (a->n_segs <= 0 || a->n_tiles <= 0) return GHN3_OK;
  GHN3_REQUIRE(a->segs && a->src && a->x && a->xt, "ghn3_expand_cols: null pointer");
  expand_kernel<<<(unsigned)a->n_tiles, 256, 0, stream>>>(*a);
  GHN3_LAUNCH_CHECK("expand_kernel");
  return GHN3_OK;
}

extern "C" int ghn3_adamw(const ghn3_adamw_args* a, ghn3_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHN3_REQUIRE(a != nullptr, "ghn3_adamw: null args");
  if (a->n_chunks <= 0) return GHN3_OK;
  GHN3_REQUIRE(a->params && a->grads && a->exp_avg && a->exp_avg_sq && a->offsets && a->numels && a->chunk0 &&
                   a->chunk_tensor, "ghn3_adamw: null pointer");
  GHN3_REQUIRE(a->bias_correction1 > 0.f && a->bias_correction2 > 0.f, "ghn3_adamw: bias corrections must be positive");
  GHN3_REQUIRE(a->range_hi == 0 || (a->range_lo >= 0 && a->range_lo < a->range_hi && a->range_hi <= a->total &&
                                    a->range_lo % 4 == 0 && a->range_hi % 4 == 0 && a->chunk_begin >= 0),
               "ghn3_adamw: bad shard range [%lld, %lld)", (long long)a->range_lo, (long long)a->range_hi);
  if (a->sumsq_ready < 0) {
    // |g|^2 of this rank's slice only (-1: *sumsq is cleared first, -2: accumulated); the caller sums over the ranks
    GHN3_REQUIRE(a->sumsq != nullptr && a->range_hi > 0, "ghn3_adamw: the slice norm needs sumsq and a shard range");
    if (a->sumsq_ready == -1) GHN3_CUDA(cudaMemsetAsync(a->sumsq, 0, sizeof(double), stream));
    sumsq_flat_kernel<<<(unsigned)(num_sms() * 8), 256, 0, stream>>>(a->grads + a->range_lo,
                                                                    a->range_hi - a->range_lo, a->sumsq);
    GHN3_LAUNCH_CHECK("sumsq_flat_kernel");
    return GHN3_OK;
  }
  if ((a->max_norm > 0.f || a->skipped != nullptr) && !a->sumsq_ready) {
    GHN3_REQUIRE(a->sumsq != nullptr, "ghn3_adamw: clipping / the non-finite guard need the sumsq scratch double");
    GHN3_REQUIRE(a->range_hi == 0, "ghn3_adamw: a sharded step needs the global |g|^2 from the caller (sumsq_ready)");
    GHN3_CUDA(cudaMemsetAsync(a->sumsq, 0, sizeof(double), stream));
    sumsq_flat_kernel<<<(unsigned)(num_sms() * 8), 256, 0, stream>>>(a->grads, a->total, a->sumsq);
    GHN3_LAUNCH_CHECK("sumsq_flat_kernel");
  }
  adamw_kernel<<<(unsigned)a->n_chunks, 256, 0, stream>>>(*a);
  GHN3_LAUNCH_CHECK("adamw_kernel");
  return GHN3_OK;
}
