// LayerNorm, fused bias-aware attention, small strided SIMT GEMM and dtype conversion (sm_100a).
//   ghn3_layernorm   reference ghn3/graphormer.py:239,241 and ghn3/nn.py:262-263
//   ghn3_attention   reference ghn3/graphormer.py:121-140 (QK^T * d^-1/2 + edge bias, softmax, PV), flash-style:
//                    no (B,H,N,N) logits and no (B,N,N,H) bias tensor are ever materialised
//   ghn3_gemm_simt   classification heads (ghn3/nn.py:757-758, 294) whose operands are transposed views
#include <atomic>
#include <cstdlib>

#include "common.cuh"

namespace ghn3 {

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row kept in registers (C <= 1024), two-pass mean / variance in fp32
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_kernel(const ghn3_layernorm_args a) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();
  pdl_wait();
  if (row >= a.rows) return;
  const int C = a.hid, C4 = C >> 2;
  const float4* x4 = (const float4*)(a.x + (int64_t)row * C);
  float4 v[8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int f = lane + 32 * i;
    if (f < C4) {
      v[i] = x4[f];
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const bool identity = a.gamma == nullptr;        // layernorm=False GHNs: conversion / row scatter only
  const float mean = warp_sum(sum) / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int f = lane + 32 * i;
    if (f < C4) {
      const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
      sq += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(sq) / (float)C + 1e-5f);
  const int orow = a.dst_row ? a.dst_row[row] : row;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int f = lane + 32 * i;
    if (f < C4) {
      float4 y = v[i];
      if (!identity) {
        const float4 g = ((const float4*)a.gamma)[f];
        const float4 b = ((const float4*)a.beta)[f];
        y.x = (v[i].x - mean) * rstd * g.x + b.x;
        y.y = (v[i].y - mean) * rstd * g.y + b.y;
        y.z = (v[i].z - mean) * rstd * g.z + b.z;
        y.w = (v[i].w - mean) * rstd * g.w + b.w;
      }
      if (a.out_f32) ((float4*)(a.out_f32 + (int64_t)row * C))[f] = y;
      if (orow >= 0 && a.out) {
        if (a.out_dtype == GHN3_BF16) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(y.x, y.y), hi = __floats2bfloat162_rn(y.z, y.w);
          uint2 pk;
          pk.x = *(uint32_t*)&lo;
          pk.y = *(uint32_t*)&hi;
          ((uint2*)((__nv_bfloat16*)a.out + (int64_t)orow * C))[f] = pk;
        } else {
          if (a.out_dtype == GHN3_TF32) {
            y.x = round_tf32(y.x); y.y = round_tf32(y.y); y.z = round_tf32(y.z); y.w = round_tf32(y.w);
          }
          ((float4*)((float*)a.out + (int64_t)orow * C))[f] = y;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Attention. CTA = (graph, head, 16 queries); 8 warps, 2 queries per warp. Keys are processed in tiles of KT = 256.
// Per tile the CTA stages, with independent 16-byte loads, K and V of the head (fp32, rows padded to D+1 words:
// conflict-free when lanes walk over keys) and the edge bias of its 16 x 256 (query, key) pairs, already looked up
// in the head's LUT (bias = lut[h][pair[i][j]] * log2e). Lane l scores keys l, l+32, ... for BOTH queries of its
// warp from one read of K / V; the tile maximum and the normaliser are combined with warp shuffles; every lane
// keeps a partial output over its own keys which is reduced across the warp once at the end.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kAttnKT = 256;
constexpr int kAttnQPerWarp = 2;
constexpr int kAttnWarps = 8;
constexpr int kAttnQT = kAttnQPerWarp * kAttnWarps;

template <typename T, int D>
__global__ void __launch_bounds__(kAttnWarps * 32, (D <= 24) ? 2 : 1) attention_kernel(const ghn3_attention_args a) {
  extern __shared__ float attn_smem[];
  constexpr int DP = D + 1;
  constexpr int R = kAttnKT / 32;
  float* sK = attn_smem;                      // [KT][DP]
  float* sV = sK + kAttnKT * DP;              // [KT][DP]
  float* sBias = sV + kAttnKT * DP;           // [QT][KT]
  float* sLut = sBias + kAttnQT * kAttnKT;    // [lut_size]

  const int g = blockIdx.z, h = blockIdx.y;
  const int n0 = a.node_off[g];
  const int n = a.node_off[g + 1] - n0;
  const int q0 = blockIdx.x * kAttnQT;
  if (q0 >= n) return;
  const int ld = (n + 15) & ~15;
  const int C = a.hid, C3 = 3 * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const T* qkv = (const T*)a.qkv + (int64_t)n0 * C3;
  const uint16_t* pair = a.pair + a.mat_off[g];
  const float scale_log2 = rsqrtf((float)D) * 1.44269504088896340736f;
  constexpr float kLog2e = 1.44269504088896340736f;

  pdl_launch_dependents();
  for (int i = threadIdx.x; i < a.lut_size; i += blockDim.x) sLut[i] = __ldg(a.lut + (int64_t)h * a.lut_size + i) * kLog2e;
  pdl_wait();

  float q[kAttnQPerWarp][D], acc[kAttnQPerWarp][D], m[kAttnQPerWarp], l[kAttnQPerWarp];
  int qi[kAttnQPerWarp];
#pragma unroll
  for (int t = 0; t < kAttnQPerWarp; ++t) {
    qi[t] = q0 + warp * kAttnQPerWarp + t;
    const bool ok = qi[t] < n;
    const T* qp = qkv + (int64_t)(ok ? qi[t] : 0) * C3 + h * D;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      q[t][d] = ok ? to_float(qp[d]) * scale_log2 : 0.f;
      acc[t][d] = 0.f;
    }
    m[t] = -INFINITY;
    l[t] = 0.f;
  }
  const bool any_q = qi[0] < n;               // warp-uniform (qi[1] valid implies qi[0] valid)
  const int nq = min(kAttnQT, n - q0);

  for (int k0 = 0; k0 < n; k0 += kAttnKT) {
    const int kt = min(kAttnKT, n - k0);
    __syncthreads();   // previous tile fully consumed; orders the LUT fill before the first bias staging
    {
      // 16-byte (8-byte for D*sizeof(T) == 8) vector loads of the K and V head slices, all independent
      constexpr int ROW_BYTES = D * (int)sizeof(T);
      constexpr int VB = (ROW_BYTES % 16 == 0) ? 16 : 8;
      constexpr int VPR = ROW_BYTES / VB;
      constexpr int EPV = VB / (int)sizeof(T);
      const char* kbase = (const char*)(qkv + (int64_t)k0 * C3 + C + h * D);
      const size_t row_stride = (size_t)C3 * sizeof(T), v_off = (size_t)C * sizeof(T);
      for (int idx = threadIdx.x; idx < kt * VPR; idx += blockDim.x) {
        const int j = idx / VPR, c = idx - j * VPR;
        const char* src = kbase + (size_t)j * row_stride + c * VB;
        union { uint4 u4; uint2 u2; T e[EPV]; } kv, vv;
        if constexpr (VB == 16) {
          kv.u4 = __ldg((const uint4*)src);
          vv.u4 = __ldg((const uint4*)(src + v_off));
        } else {
          kv.u2 = __ldg((const uint2*)src);
          vv.u2 = __ldg((const uint2*)(src + v_off));
        }
#pragma unroll
        for (int e = 0; e < EPV; ++e) {
          sK[j * DP + c * EPV + e] = to_float(kv.e[e]);
          sV[j * DP + c * EPV + e] = to_float(vv.e[e]);
        }
      }
      // edge bias of the (query, key) pairs of this tile: 8 pair indices per 16-byte load -> 8 LUT values
      const int kt8 = (kt + 7) >> 3;           // rows are padded to ld (multiple of 16), so this never overruns
      for (int idx = threadIdx.x; idx < nq * kt8; idx += blockDim.x) {
        const int qq = idx / kt8, c = idx - qq * kt8;
        union { uint4 u4; uint16_t e[8]; } pv;
        pv.u4 = __ldg((const uint4*)(pair + (int64_t)(q0 + qq) * ld + k0 + c * 8));
#pragma unroll
        for (int e = 0; e < 8; ++e) sBias[qq * kAttnKT + c * 8 + e] = sLut[pv.e[e]];
      }
    }
    __syncthreads();
    if (any_q) {
      const float* b0 = sBias + (warp * kAttnQPerWarp) * kAttnKT;
      float s[kAttnQPerWarp][R];
      float tmax[kAttnQPerWarp];
#pragma unroll
      for (int t = 0; t < kAttnQPerWarp; ++t) tmax[t] = -INFINITY;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int j = lane + 32 * r;
        if (j < kt) {
          const float* kp = sK + j * DP;
          float dot[kAttnQPerWarp];
#pragma unroll
          for (int t = 0; t < kAttnQPerWarp; ++t) dot[t] = b0[t * kAttnKT + j];
#pragma unroll
          for (int d = 0; d < D; ++d) {
            const float kd = kp[d];
#pragma unroll
            for (int t = 0; t < kAttnQPerWarp; ++t) dot[t] = fmaf(q[t][d], kd, dot[t]);
          }
#pragma unroll
          for (int t = 0; t < kAttnQPerWarp; ++t) {
            s[t][r] = dot[t];
            tmax[t] = fmaxf(tmax[t], dot[t]);
          }
        } else {
#pragma unroll
          for (int t = 0; t < kAttnQPerWarp; ++t) s[t][r] = -INFINITY;
        }
      }
      float m_new[kAttnQPerWarp];
#pragma unroll
      for (int t = 0; t < kAttnQPerWarp; ++t) {
        m_new[t] = fmaxf(m[t], warp_max(tmax[t]));
        const float corr = exp2f(m[t] - m_new[t]);
        m[t] = m_new[t];
        l[t] *= corr;
#pragma unroll
        for (int d = 0; d < D; ++d) acc[t][d] *= corr;
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int j = lane + 32 * r;
        if (j < kt) {
          float p[kAttnQPerWarp];
#pragma unroll
          for (int t = 0; t < kAttnQPerWarp; ++t) {
            p[t] = exp2f(s[t][r] - m_new[t]);
            l[t] += p[t];
          }
          const float* vp = sV + j * DP;
#pragma unroll
          for (int d = 0; d < D; ++d) {
            const float vd = vp[d];
#pragma unroll
            for (int t = 0; t < kAttnQPerWarp; ++t) acc[t][d] = fmaf(p[t], vd, acc[t][d]);
          }
        }
      }
    }
  }

#pragma unroll
  for (int t = 0; t < kAttnQPerWarp; ++t) {
    if (qi[t] >= n) continue;
    const float inv = 1.f / warp_sum(l[t]);
    float mine = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const float tot = warp_sum(acc[t][d]);
      if (lane == d) mine = tot * inv;
    }
    if (lane < D) {
      T* op = (T*)a.out + (int64_t)(n0 + qi[t]) * C + h * D + lane;
      if (sizeof(T) == 4 && a.dtype == GHN3_TF32) mine = round_tf32(mine);
      *op = from_float<T>(mine);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Attention on the tensor cores (bf16 storage): mma.sync m16n8k16, flash-attention-2 register layout.
// CTA = (graph, head, 64 queries); 4 warps x 16 queries. K (key-major, dims zero-padded to a multiple of 16) and
// V^T (dim-major) of the head are staged per 256-key tile as bf16 with row strides chosen so that every
// B-fragment load is bank-conflict free. Logits start from the edge bias lut[h][pair[i][j]] (pair indices are read
// as 32-bit words straight from global memory, issued before the QK^T MMAs so their latency overlaps), softmax is
// the usual online form with quad shuffles, P is re-packed in registers as the A operand of P.V.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMmaKT = 256;      // keys staged per outer iteration
constexpr int kMmaWarps = 8;
constexpr int kMmaQT = 16 * kMmaWarps;       // queries per CTA

template <int D>
__global__ void __launch_bounds__(kMmaWarps * 32) attention_mma_kernel(const ghn3_attention_args a) {
  constexpr int DK = (D + 15) / 16 * 16;   // k extent of Q.K^T
  constexpr int DS = DK + 8;               // K row stride (bf16 elements)
  constexpr int DN = (D + 7) / 8 * 8;      // n extent of P.V
  constexpr int VS = kMmaKT + 8;           // V^T row stride (bf16 elements)
  constexpr int NT2 = DN / 8;
  constexpr int KK = DK / 16;
  constexpr int NTHREADS = kMmaWarps * 32;
  constexpr int ROW_BYTES = D * 2;
  constexpr int VB = (ROW_BYTES % 16 == 0) ? 16 : 8;
  constexpr int VPR = ROW_BYTES / VB;      // vectors per K / V row
  constexpr int EPV = VB / 2;
  constexpr int NV = (kMmaKT * VPR + NTHREADS - 1) / NTHREADS;   // staging vectors per thread and operand
  extern __shared__ __align__(16) uint8_t attn_mma_smem[];
  __nv_bfloat16* sK = (__nv_bfloat16*)attn_mma_smem;            // [KT][DS]
  __nv_bfloat16* sVt = sK + kMmaKT * DS;                        // [DN][VS]
  float* sLut = (float*)(sVt + DN * VS);                        // [lut_size]

  const int g = blockIdx.z, h = blockIdx.y;
  const int n0 = a.node_off[g];
  const int n = a.node_off[g + 1] - n0;
  const int q0 = blockIdx.x * kMmaQT;
  if (q0 >= n) return;
  const int ld = (n + 15) & ~15;
  const int C = a.hid, C3 = 3 * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const __nv_bfloat16* qkv = (const __nv_bfloat16*)a.qkv + (int64_t)n0 * C3;
  const uint16_t* pair = a.pair + a.mat_off[g];
  const float scale_log2 = rsqrtf((float)D) * 1.44269504088896340736f;
  constexpr float kLog2e = 1.44269504088896340736f;

  pdl_launch_dependents();
  for (int i = threadIdx.x; i < a.lut_size; i += NTHREADS) sLut[i] = __ldg(a.lut + (int64_t)h * a.lut_size + i) * kLog2e;
  // zero the padded dims of K once (they are never overwritten) and the padded dims of V^T
  if (DK > D) {
    for (int idx = threadIdx.x; idx < kMmaKT * (DK - D); idx += NTHREADS) {
      const int j = idx / (DK - D), d = D + idx % (DK - D);
      sK[j * DS + d] = __float2bfloat16_rn(0.f);
    }
  }
  if (DN > D) {
    for (int idx = threadIdx.x; idx < (DN - D) * VS; idx += NTHREADS) sVt[D * VS + idx] = __float2bfloat16_rn(0.f);
  }
  pdl_wait();       // everything above only touched the LUT and shared memory

  // K / V staging: every thread owns NV (row, vector) slots of a tile; all loads of a tile are issued back to back
  // into registers (one memory round trip), and the NEXT tile is fetched while the current one is being consumed.
  uint4 kreg[NV], vreg[NV];
  const size_t row_stride = (size_t)C3 * 2, v_off = (size_t)C * 2;
  auto load_tile = [&](int k0) {
    const int kt = min(kMmaKT, n - k0);
    const char* kbase = (const char*)(qkv + (int64_t)k0 * C3 + C + h * D);
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      const int idx = threadIdx.x + u * NTHREADS;
      const int j = idx / VPR, c = idx - j * VPR;
      kreg[u] = make_uint4(0, 0, 0, 0);
      vreg[u] = make_uint4(0, 0, 0, 0);
      if (j < kt) {
        const char* src = kbase + (size_t)j * row_stride + c * VB;
        if constexpr (VB == 16) {
          kreg[u] = __ldg((const uint4*)src);
          vreg[u] = __ldg((const uint4*)(src + v_off));
        } else {
          const uint2 t0 = __ldg((const uint2*)src), t1 = __ldg((const uint2*)(src + v_off));
          kreg[u].x = t0.x; kreg[u].y = t0.y;
          vreg[u].x = t1.x; vreg[u].y = t1.y;
        }
      }
    }
  };
  auto store_tile = [&](int k0) {
    const int kt64 = (min(kMmaKT, n - k0) + 63) & ~63;      // rows up to the 64-key chunk boundary get zeros
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      const int idx = threadIdx.x + u * NTHREADS;
      const int j = idx / VPR, c = idx - j * VPR;
      if (j < kt64) {
        if constexpr (VB == 16) *(uint4*)(sK + j * DS + c * EPV) = kreg[u];
        else *(uint2*)(sK + j * DS + c * EPV) = make_uint2(kreg[u].x, kreg[u].y);
        const __nv_bfloat16* ve = (const __nv_bfloat16*)&vreg[u];
#pragma unroll
        for (int e = 0; e < EPV; ++e) sVt[(c * EPV + e) * VS + j] = ve[e];
      }
    }
  };
  load_tile(0);

  // Q fragments of this warp's 16 queries, pre-scaled by d^-1/2 * log2(e)
  const int r0 = q0 + warp * 16 + gq, r1 = r0 + 8;
  const bool ok0 = r0 < n, ok1 = r1 < n;
  uint32_t aq[KK][4];
#pragma unroll
  for (int kk = 0; kk < KK; ++kk) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int d = kk * 16 + half * 8 + 2 * tq;
      float v00 = 0.f, v01 = 0.f, v10 = 0.f, v11 = 0.f;
      if (d < D) {
        if (ok0) {
          const __nv_bfloat162 t = *(const __nv_bfloat162*)(qkv + (int64_t)r0 * C3 + h * D + d);
          v00 = __low2float(t) * scale_log2; v01 = __high2float(t) * scale_log2;
        }
        if (ok1) {
          const __nv_bfloat162 t = *(const __nv_bfloat162*)(qkv + (int64_t)r1 * C3 + h * D + d);
          v10 = __low2float(t) * scale_log2; v11 = __high2float(t) * scale_log2;
        }
      }
      aq[kk][half * 2 + 0] = pack_bf16(v00, v01);
      aq[kk][half * 2 + 1] = pack_bf16(v10, v11);
    }
  }
  const uint16_t* prow0 = pair + (int64_t)(ok0 ? r0 : q0) * ld;
  const uint16_t* prow1 = pair + (int64_t)(ok1 ? r1 : q0) * ld;
  const bool warp_active = (q0 + warp * 16) < n;     // warp-uniform

  // edge-bias indices of a 16 x 64 block: two keys per 32-bit load, fetched one block ahead of their use
  uint32_t pw0[8], pw1[8];
  auto load_pairs = [&](int kbase) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int col = kbase + nt * 8 + 2 * tq;
      pw0[nt] = 0; pw1[nt] = 0;
      if (warp_active && col < n) {
        pw0[nt] = __ldg((const uint32_t*)(prow0 + col));
        pw1[nt] = __ldg((const uint32_t*)(prow1 + col));
      }
    }
  };
  load_pairs(0);

  float o[NT2][4];
#pragma unroll
  for (int i = 0; i < NT2; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int k0 = 0; k0 < n; k0 += kMmaKT) {
    const int kt = min(kMmaKT, n - k0);
    __syncthreads();                       // the previous tile has been consumed by every warp
    store_tile(k0);
    __syncthreads();
    if (k0 + kMmaKT < n) load_tile(k0 + kMmaKT);       // in flight while this tile is consumed
    if (!warp_active) continue;
    for (int c0 = 0; c0 < kt; c0 += 64) {
      uint32_t cw0[8], cw1[8];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) { cw0[nt] = pw0[nt]; cw1[nt] = pw1[nt]; }
      if (k0 + c0 + 64 < n) load_pairs(k0 + c0 + 64);  // next block's indices, consumed one iteration later
      float s[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < KK; ++kk) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const __nv_bfloat16* kp = sK + (c0 + nt * 8 + gq) * DS + kk * 16 + 2 * tq;
          mma_bf16_16816(s[nt], aq[kk], *(const uint32_t*)kp, *(const uint32_t*)(kp + 8));
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int col = k0 + c0 + nt * 8 + 2 * tq;
        const bool v0 = col < n, v1 = col + 1 < n;
        s[nt][0] = v0 ? s[nt][0] + sLut[cw0[nt] & 0xFFFFu] : -INFINITY;
        s[nt][1] = v1 ? s[nt][1] + sLut[cw0[nt] >> 16] : -INFINITY;
        s[nt][2] = v0 ? s[nt][2] + sLut[cw1[nt] & 0xFFFFu] : -INFINITY;
        s[nt][3] = v1 ? s[nt][3] + sLut[cw1[nt] >> 16] : -INFINITY;
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
      const float corr0 = exp2f(m0 - mn0), corr1 = exp2f(m1 - mn1);
      m0 = mn0; m1 = mn1;
      l0 *= corr0; l1 *= corr1;
#pragma unroll
      for (int i = 0; i < NT2; ++i) { o[i][0] *= corr0; o[i][1] *= corr0; o[i][2] *= corr1; o[i][3] *= corr1; }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = exp2f(s[nt][0] - mn0); s[nt][1] = exp2f(s[nt][1] - mn0);
        s[nt][2] = exp2f(s[nt][2] - mn1); s[nt][3] = exp2f(s[nt][3] - mn1);
        l0 += s[nt][0] + s[nt][1];
        l1 += s[nt][2] + s[nt][3];
      }
#pragma unroll
      for (int k2 = 0; k2 < 4; ++k2) {
        uint32_t ap[4];
        ap[0] = pack_bf16(s[2 * k2][0], s[2 * k2][1]);
        ap[1] = pack_bf16(s[2 * k2][2], s[2 * k2][3]);
        ap[2] = pack_bf16(s[2 * k2 + 1][0], s[2 * k2 + 1][1]);
        ap[3] = pack_bf16(s[2 * k2 + 1][2], s[2 * k2 + 1][3]);
#pragma unroll
        for (int i = 0; i < NT2; ++i) {
          const __nv_bfloat16* vp = sVt + (i * 8 + gq) * VS + c0 + k2 * 16 + 2 * tq;
          mma_bf16_16816(o[i], ap, *(const uint32_t*)vp, *(const uint32_t*)(vp + 8));
        }
      }
    }
  }
  if (!warp_active) return;
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  if (a.lse2 != nullptr && tq == 0) {          // softmax statistics for the backward pass (log2 domain)
    if (ok0) a.lse2[(int64_t)h * a.total_nodes + n0 + r0] = m0 + log2f(l0);
    if (ok1) a.lse2[(int64_t)h * a.total_nodes + n0 + r1] = m1 + log2f(l1);
  }
  __nv_bfloat16* out = (__nv_bfloat16*)a.out;
#pragma unroll
  for (int i = 0; i < NT2; ++i) {
    const int d = i * 8 + 2 * tq;
    if (d < D) {
      if (ok0) *(uint32_t*)(out + (int64_t)(n0 + r0) * C + h * D + d) = pack_bf16(o[i][0] * i0, o[i][1] * i0);
      if (ok1) *(uint32_t*)(out + (int64_t)(n0 + r1) * C + h * D + d) = pack_bf16(o[i][2] * i1, o[i][3] * i1);
    }
  }
}

template <int D>
static int launch_attention_mma(const ghn3_attention_args* a, cudaStream_t stream) {
  constexpr int DK = (D + 15) / 16 * 16, DN = (D + 7) / 8 * 8;
  const int smem = (kMmaKT * (DK + 8) + DN * (kMmaKT + 8)) * 2 + a->lut_size * (int)sizeof(float);
  static bool configured = false;
  if (!configured) {
    GHN3_CUDA(cudaFuncSetAttribute(attention_mma_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  if (smem > 200 * 1024) {
    set_error("ghn3_attention: lut too large for shared memory");
    return GHN3_ERR_UNSUPPORTED;
  }
  const dim3 grid((unsigned)ceil_div(a->max_nodes, kMmaQT), (unsigned)a->heads, (unsigned)a->n_graphs);
  GHN3_CUDA(launch_pdl(attention_mma_kernel<D>, grid, dim3(kMmaWarps * 32), (size_t)smem, stream, *a));
  GHN3_LAUNCH_CHECK("attention_mma_kernel");
  return GHN3_OK;
}

template <typename T, int D>
static int launch_attention(const ghn3_attention_args* a, cudaStream_t stream) {
  const int smem = (2 * kAttnKT * (D + 1) + kAttnQT * kAttnKT + a->lut_size) * (int)sizeof(float);
  static bool configured = false;
  if (!configured) {
    GHN3_CUDA(cudaFuncSetAttribute(attention_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  if (smem > 200 * 1024) {
    set_error("ghn3_attention: lut too large for shared memory");
    return GHN3_ERR_UNSUPPORTED;
  }
  const dim3 grid((unsigned)ceil_div(a->max_nodes, kAttnQT), (unsigned)a->heads, (unsigned)a->n_graphs);
  GHN3_CUDA(launch_pdl(attention_kernel<T, D>, grid, dim3(kAttnWarps * 32), (size_t)smem, stream, *a));
  GHN3_LAUNCH_CHECK("attention_kernel");
  return GHN3_OK;
}

int attention_split_impl(const ghn3_attention_args* a, cudaStream_t stream);
int attention_tc_impl(const ghn3_attention_args* a, cudaStream_t stream);

static std::atomic<int> g_attn_tc_min{getenv("GHN3_ATTN_TC_MIN") ? atoi(getenv("GHN3_ATTN_TC_MIN")) : 2048};

int attention_impl(const ghn3_attention_args* a, cudaStream_t stream) {
  GHN3_REQUIRE(a != nullptr, "ghn3_attention: null args");
  GHN3_REQUIRE(a->heads > 0 && a->hid % a->heads == 0, "ghn3_attention: hid must be divisible by heads");
  GHN3_REQUIRE(a->dtype >= GHN3_BF16 && a->dtype <= GHN3_F32, "ghn3_attention: bad dtype");
  if (a->n_graphs <= 0 || a->max_nodes <= 0) return GHN3_OK;
  const int D = a->hid / a->heads;
  const bool bf = a->dtype == GHN3_BF16;
  if (bf) {
    // large graphs: the tcgen05 kernel (attention_tcgen05.cu); below the threshold the mma.sync kernel's shorter
    // prologue and larger CTA count win (measured cross-over between 1024 and 2048 nodes). GHN3_ATTN_TC_MIN overrides the threshold (0 = always, a huge value = never).
    if (a->max_nodes >= g_attn_tc_min.load(std::memory_order_relaxed)) {
      const int rc = attention_tc_impl(a, stream);
      if (rc != GHN3_ERR_UNSUPPORTED) return rc;
    }
  }
  if (!bf) {
    // fp32 storage: split-bf16 tensor-core kernel (attention_split.cu); the CUDA-core kernel below remains for the
    // head dims it does not cover and as a cross-check (GHN3_NO_SPLIT_ATTN=1)
    static const bool no_split = getenv("GHN3_NO_SPLIT_ATTN") != nullptr;
    if (!no_split) {
      const int rc = attention_split_impl(a, stream);
      if (rc != GHN3_ERR_UNSUPPORTED) return rc;
    }
  }
#define GHN3_ATTN_CASE(DV)                                                     \
  if (D == DV) return bf ? launch_attention_mma<DV>(a, stream) : launch_attention<float, DV>(a, stream);
  GHN3_ATTN_CASE(4)
  GHN3_ATTN_CASE(8)
  GHN3_ATTN_CASE(16)
  GHN3_ATTN_CASE(24)
  GHN3_ATTN_CASE(32)
#undef GHN3_ATTN_CASE
  set_error("ghn3_attention: head dim %d is not supported (4, 8, 16, 24, 32)", D);
  return GHN3_ERR_UNSUPPORTED;
}

int layernorm_impl(const ghn3_layernorm_args* a, cudaStream_t stream) {
  GHN3_REQUIRE(a != nullptr, "ghn3_layernorm: null args");
  GHN3_REQUIRE(a->hid > 0 && a->hid % 4 == 0 && a->hid <= 1024, "ghn3_layernorm: hid must be a multiple of 4, <= 1024");
  GHN3_REQUIRE(a->out_dtype >= GHN3_BF16 && a->out_dtype <= GHN3_F32, "ghn3_layernorm: bad out_dtype");
  if (a->rows <= 0) return GHN3_OK;
  GHN3_CUDA(launch_pdl(layernorm_kernel, dim3((unsigned)ceil_div(a->rows, 8)), dim3(256), 0, stream, *a));
  GHN3_LAUNCH_CHECK("layernorm_kernel");
  return GHN3_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Strided fp32 GEMM on CUDA cores: 64 x 64 tile, 256 threads, 4 x 4 outputs per thread, K step 16
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gemm_simt_kernel(const ghn3_gemm_simt_args a) {
  __shared__ float sA[16][65];
  __shared__ float sB[16][65];
  const int bz = blockIdx.z;
  const float* A = a.a + (int64_t)bz * a.a_bs;
  float* Dp = a.d + (int64_t)bz * a.d_bs;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const bool a_k_fast = a.sak == 1, b_k_fast = a.sbk == 1;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < a.k; k0 += 16) {
    for (int idx = threadIdx.x; idx < 64 * 16; idx += 256) {
      // consecutive threads walk the operand's contiguous dimension (k if its stride is 1, else the row index)
      const int kka = a_k_fast ? (idx & 15) : (idx >> 6), ra = a_k_fast ? (idx >> 4) : (idx & 63);
      const int kkb = b_k_fast ? (idx & 15) : (idx >> 6), rb = b_k_fast ? (idx >> 4) : (idx & 63);
      float va = 0.f, vb = 0.f;
      if (k0 + kka < a.k && m0 + ra < a.m) {
        va = __ldg(A + (int64_t)(m0 + ra) * a.sam + (int64_t)(k0 + kka) * a.sak);
        if (a.relu_a) va = fmaxf(va, 0.f);
      }
      if (k0 + kkb < a.k && n0 + rb < a.n) vb = __ldg(a.b + (int64_t)(n0 + rb) * a.sbn + (int64_t)(k0 + kkb) * a.sbk);
      sA[kka][ra] = va;
      sB[kkb][rb] = vb;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float ra[4], rb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) ra[i] = sA[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) rb[j] = sB[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ra[i], rb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= a.m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= a.n) continue;
      float v = acc[i][j] + (a.bias ? a.bias[n] : 0.f);
      if (a.act == GHN3_ACT_RELU) v = fmaxf(v, 0.f);
      Dp[(int64_t)m * a.sdm + (int64_t)n * a.sdn] = v;
    }
  }
}

// Skinny case (m <= 8 rows, unit k strides): one warp per output column n, lanes over k, all rows at once.
__global__ void __launch_bounds__(256) gemm_skinny_kernel(const ghn3_gemm_simt_args a) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= a.n) return;
  const float* A = a.a + (int64_t)blockIdx.z * a.a_bs;
  float* Dp = a.d + (int64_t)blockIdx.z * a.d_bs;
  const float* brow = a.b + (int64_t)n * a.sbn;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int k = lane; k < a.k; k += 32) {
    const float w = __ldg(brow + k);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < a.m) {
        float v = __ldg(A + (int64_t)i * a.sam + k);
        if (a.relu_a) v = fmaxf(v, 0.f);
        acc[i] = fmaf(v, w, acc[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i < a.m) {
      float v = warp_sum(acc[i]);
      if (lane == 0) {
        v += a.bias ? a.bias[n] : 0.f;
        if (a.act == GHN3_ACT_RELU) v = fmaxf(v, 0.f);
        Dp[(int64_t)i * a.sdm + (int64_t)n * a.sdn] = v;
      }
    }
  }
}

// 32 x 32 tiles through shared memory: coalesced reads along c, coalesced writes along r
__global__ void __launch_bounds__(256) relu_transpose_kernel(const ghn3_relu_transpose_args a) {
  __shared__ float tile[32][33];
  const int z = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  pdl_launch_dependents();
  pdl_wait();
  const float* src = a.src + (int64_t)z * a.src_bs;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < a.rows && c < a.cols) ? fmaxf(src[(int64_t)r * a.ld + c], 0.f) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < a.cols && r < a.rows) {
      const float v = tile[tx][i];
      const int64_t o = ((int64_t)z * a.cols + c) * a.rows + r;
      if (a.dst_dtype == GHN3_BF16) ((__nv_bfloat16*)a.dst)[o] = __float2bfloat16_rn(v);
      else ((float*)a.dst)[o] = a.dst_dtype == GHN3_TF32 ? round_tf32(v) : v;
    }
  }
}

__global__ void convert_kernel(const float* __restrict__ src, void* __restrict__ dst, int64_t n, int dst_dtype) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = src[i];
    if (dst_dtype == GHN3_BF16) ((__nv_bfloat16*)dst)[i] = __float2bfloat16_rn(v);
    else ((float*)dst)[i] = dst_dtype == GHN3_TF32 ? round_tf32(v) : v;
  }
}

}  // namespace ghn3

using namespace ghn3;

extern "C" int ghn3_layernorm(const ghn3_layernorm_args* a, ghn3_stream_t stream) {
  return layernorm_impl(a, (cudaStream_t)stream);
}

extern "C" int ghn3_set_attention_tc_min(int min_nodes) {
  return g_attn_tc_min.exchange(min_nodes);
}

extern "C" int ghn3_attention(const ghn3_attention_args* a, ghn3_stream_t stream) {
  return attention_impl(a, (cudaStream_t)stream);
}

extern "C" int ghn3_gemm_simt(const ghn3_gemm_simt_args* a, ghn3_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHN3_REQUIRE(a != nullptr, "ghn3_gemm_simt: null args");
  if (a->m <= 0 || a->n <= 0) return GHN3_OK;
  const int batch = a->batch <= 0 ? 1 : a->batch;
  if (a->m <= 8 && a->sak == 1 && a->sbk == 1) {
    gemm_skinny_kernel<<<dim3((unsigned)ceil_div(a->n, 8), 1, (unsigned)batch), 256, 0, stream>>>(*a);
    GHN3_LAUNCH_CHECK("gemm_skinny_kernel");
    return GHN3_OK;
  }
  const dim3 grid((unsigned)ceil_div(a->n, 64), (unsigned)ceil_div(a->m, 64), (unsigned)batch);
  gemm_simt_kernel<<<grid, 256, 0, stream>>>(*a);
  GHN3_LAUNCH_CHECK("gemm_simt_kernel");
  return GHN3_OK;
}

extern "C" int ghn3_relu_transpose(const ghn3_relu_transpose_args* a, ghn3_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHN3_REQUIRE(a != nullptr, "ghn3_relu_transpose: null args");
  GHN3_REQUIRE(a->dst_dtype >= GHN3_BF16 && a->dst_dtype <= GHN3_F32, "ghn3_relu_transpose: bad dtype");
  if (a->rows <= 0 || a->cols <= 0 || a->batch <= 0) return GHN3_OK;
  const dim3 grid((unsigned)ceil_div(a->cols, 32), (unsigned)ceil_div(a->rows, 32), (unsigned)a->batch);
  GHN3_CUDA(launch_pdl(relu_transpose_kernel, grid, dim3(256), 0, stream, *a));
  GHN3_LAUNCH_CHECK("relu_transpose_kernel");
  return GHN3_OK;
}

extern "C" int ghn3_convert_f32(const float* src, void* dst, int64_t n, int32_t dst_dtype, ghn3_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHN3_REQUIRE(dst_dtype >= GHN3_BF16 && dst_dtype <= GHN3_F32, "ghn3_convert_f32: bad dtype");
  if (n <= 0) return GHN3_OK;
  const int blocks = (int)std::min<int64_t>(ceil_div(n, 256), 148 * 16);
  convert_kernel<<<blocks, 256, 0, stream>>>(src, dst, n, dst_dtype);
  GHN3_LAUNCH_CHECK("convert_kernel");
  return GHN3_OK;
}
