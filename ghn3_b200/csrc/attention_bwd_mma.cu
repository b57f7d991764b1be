// Attention backward on the tensor cores (bf16 storage): mma.sync m16n8k16 with the register layout of the forward
// kernel (dense_kernels.cu: attention_mma_kernel). Adjoint of ghn3/graphormer.py:121-140 with the edge bias:
//   S = Q K^T d^-1/2 + lut[h][pair];  P = softmax(S);  O = P V
//   delta_i = dO_i . O_i;  dP = dO V^T;  dS = P o (dP - delta);  dQ = dS K d^-1/2;  dK = dS^T Q d^-1/2;  dV = P^T dO
// Two kernels, both "forward shaped" (one S-like product pair, then P.V-like accumulations):
//   A  rows = queries, columns = keys:    S, dP -> dS -> dQ += dS K;   also delta and the running sum of dS over layers
//   B  rows = keys,    columns = queries: S^T, dP^T -> P^T, dS^T -> dV += P^T dO, dK += dS^T Q
// The softmax statistics come from the forward kernel (log2 domain, ghn3_attention_args.lse2). The edge-bias gradient
// is NOT binned per layer: dS is accumulated into ds_total[g][h][i][j] (plain read-modify-write, each element has one
// owner) and binned into the (vmax+1)^2 x H look-up table once per backward pass by ghn3_lut_bin -- the bias is shared
// by all layers (graphormer.py:126-130), so the per-layer shared-memory atomics of the CUDA-core path go away.
#include "common.cuh"

namespace ghn3 {

constexpr int kBmKT = 256;       // columns (keys for A, queries for B) staged per outer iteration
constexpr int kBmWarps = 8;
constexpr int kBmRows = 16 * kBmWarps;
constexpr int kBmChunk = 32;     // columns per inner step: small enough for 2 CTAs per SM (register budget)
constexpr int kBmNT = kBmChunk / 8, kBmK2 = kBmChunk / 16;
constexpr float kLog2e = 1.44269504088896340736f;
constexpr float kLn2 = 0.69314718055994530942f;

template <int D>
struct BmDims {
  static constexpr int DK = (D + 15) / 16 * 16;   // k extent of the S-like products
  static constexpr int DS = DK + 8;               // row stride of the [column][dim] tiles (bf16 elements)
  static constexpr int DN = (D + 7) / 8 * 8;      // n extent of the P.V-like products
  static constexpr int VS = kBmKT + 8;            // row stride of the [dim][column] tiles
  static constexpr int NT2 = DN / 8;
  static constexpr int KK = DK / 16;
};

// rows x D slice (row stride `ld` elements) -> dst [kBmKT][DS] (and dstT [DN][VS] if given), values * mul, zero
// padding for rows in [rows, roundup64(rows)) and dims in [D, DK / DN)
template <int D>
__device__ __forceinline__ void bm_stage(__nv_bfloat16* dst, __nv_bfloat16* dstT, const __nv_bfloat16* src, int64_t ld,
                                         int rows, float mul) {
  using X = BmDims<D>;
  constexpr int HP = X::DK / 2;                   // bf16 pairs per row, padded
  const int rows64 = (rows + 63) & ~63;
#pragma unroll 4                                   // four independent global loads in flight per thread
  for (int idx = threadIdx.x; idx < rows64 * HP; idx += blockDim.x) {
    const int j = idx / HP, d = (idx - j * HP) * 2;
    float v0 = 0.f, v1 = 0.f;
    if (j < rows && d < D) {
      const __nv_bfloat162 t = *(const __nv_bfloat162*)(src + (int64_t)j * ld + d);
      v0 = __low2float(t) * mul;
      v1 = __high2float(t) * mul;
    }
    const __nv_bfloat162 o = __floats2bfloat162_rn(v0, v1);
    *(__nv_bfloat162*)(dst + j * X::DS + d) = o;
    if (dstT != nullptr && d < X::DN) {
      dstT[d * X::VS + j] = __low2bfloat16(o);
      dstT[(d + 1) * X::VS + j] = __high2bfloat16(o);
    }
  }
}

// A fragments of 16 rows (r0 = base + gq, r1 = r0 + 8) of a [*, ld] bf16 matrix, values * mul
template <int D>
__device__ __forceinline__ void bm_row_frags(uint32_t (&fr)[BmDims<D>::KK][4], const __nv_bfloat16* src, int64_t ld,
                                             int64_t r0, int64_t r1, bool ok0, bool ok1, int tq, float mul) {
#pragma unroll
  for (int kk = 0; kk < BmDims<D>::KK; ++kk) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int d = kk * 16 + half * 8 + 2 * tq;
      float v00 = 0.f, v01 = 0.f, v10 = 0.f, v11 = 0.f;
      if (d < D) {
        if (ok0) {
          const __nv_bfloat162 t = *(const __nv_bfloat162*)(src + r0 * ld + d);
          v00 = __low2float(t) * mul; v01 = __high2float(t) * mul;
        }
        if (ok1) {
          const __nv_bfloat162 t = *(const __nv_bfloat162*)(src + r1 * ld + d);
          v10 = __low2float(t) * mul; v11 = __high2float(t) * mul;
        }
      }
      fr[kk][half * 2 + 0] = pack_bf16(v00, v01);
      fr[kk][half * 2 + 1] = pack_bf16(v10, v11);
    }
  }
}

template <int D>
__global__ void __launch_bounds__(kBmWarps * 32, 2) attention_bwd_mma_dq_kernel(const ghn3_attention_bwd_args a) {
  using X = BmDims<D>;
  extern __shared__ __align__(16) uint8_t bm_smem[];
  __nv_bfloat16* sK = (__nv_bfloat16*)bm_smem;                  // [KT][DS]
  __nv_bfloat16* sV = sK + kBmKT * X::DS;                       // [KT][DS]
  __nv_bfloat16* sKt = sV + kBmKT * X::DS;                      // [DN][VS]
  float* sLut = (float*)(sKt + X::DN * X::VS);                  // [lut_size], pre-multiplied by log2(e)

  const int g = blockIdx.z, h = blockIdx.y;
  const int n0 = a.node_off[g];
  const int n = a.node_off[g + 1] - n0;
  const int q0 = blockIdx.x * kBmRows;
  if (q0 >= n) return;
  const int ld = (n + 15) & ~15;
  const int C = a.hid, C3 = 3 * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const __nv_bfloat16* qkv = (const __nv_bfloat16*)a.qkv + (int64_t)n0 * C3;
  const __nv_bfloat16* dout = (const __nv_bfloat16*)a.d_out + (int64_t)n0 * C;
  const __nv_bfloat16* outp = (const __nv_bfloat16*)a.out + (int64_t)n0 * C;
  const uint16_t* pair = a.pair + a.mat_off[g];
  const float scale = rsqrtf((float)D);

  pdl_launch_dependents();
  for (int i = threadIdx.x; i < a.lut_size; i += blockDim.x) sLut[i] = a.lut[(int64_t)h * a.lut_size + i] * kLog2e;
  pdl_wait();

  const int r0 = q0 + warp * 16 + gq, r1 = r0 + 8;
  const bool ok0 = r0 < n, ok1 = r1 < n;
  const bool warp_active = (q0 + warp * 16) < n;
  uint32_t aq[X::KK][4], ado[X::KK][4];
  bm_row_frags<D>(aq, qkv + h * D, C3, r0, r1, ok0, ok1, tq, scale * kLog2e);     // same rounding as the forward
  bm_row_frags<D>(ado, dout + h * D, C, r0, r1, ok0, ok1, tq, 1.f);
  // delta = dO . O per row: every lane of a quad holds a quarter of the dims
  float dl0 = 0.f, dl1 = 0.f;
#pragma unroll
  for (int kk = 0; kk < X::KK; ++kk) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int d = kk * 16 + half * 8 + 2 * tq;
      if (d < D) {
        if (ok0) {
          const __nv_bfloat162 x = *(const __nv_bfloat162*)(dout + (int64_t)r0 * C + h * D + d);
          const __nv_bfloat162 y = *(const __nv_bfloat162*)(outp + (int64_t)r0 * C + h * D + d);
          dl0 += __low2float(x) * __low2float(y) + __high2float(x) * __high2float(y);
        }
        if (ok1) {
          const __nv_bfloat162 x = *(const __nv_bfloat162*)(dout + (int64_t)r1 * C + h * D + d);
          const __nv_bfloat162 y = *(const __nv_bfloat162*)(outp + (int64_t)r1 * C + h * D + d);
          dl1 += __low2float(x) * __low2float(y) + __high2float(x) * __high2float(y);
        }
      }
    }
  }
  dl0 += __shfl_xor_sync(0xffffffffu, dl0, 1);
  dl0 += __shfl_xor_sync(0xffffffffu, dl0, 2);
  dl1 += __shfl_xor_sync(0xffffffffu, dl1, 1);
  dl1 += __shfl_xor_sync(0xffffffffu, dl1, 2);
  float lse0 = 0.f, lse1 = 0.f;
  if (ok0) lse0 = a.fwd_lse2[(int64_t)h * a.total_nodes + n0 + r0];
  if (ok1) lse1 = a.fwd_lse2[(int64_t)h * a.total_nodes + n0 + r1];
  if (tq == 0) {
    if (ok0) a.delta[(int64_t)h * a.total_nodes + n0 + r0] = dl0;
    if (ok1) a.delta[(int64_t)h * a.total_nodes + n0 + r1] = dl1;
  }
  const uint16_t* prow0 = pair + (int64_t)(ok0 ? r0 : q0) * ld;
  const uint16_t* prow1 = pair + (int64_t)(ok1 ? r1 : q0) * ld;
  float* dsp = a.ds_total != nullptr ? a.ds_total + a.mat_off[g] * a.heads + (int64_t)h * n * ld : nullptr;
  // edge-bias indices of a 16 x 64 chunk (two columns per 32-bit word), fetched one chunk ahead of their use
  uint32_t nw0[kBmNT], nw1[kBmNT];
  auto load_pairs = [&](int cbase) {
#pragma unroll
    for (int nt = 0; nt < kBmNT; ++nt) {
      const int col = cbase + nt * 8 + 2 * tq;
      nw0[nt] = 0; nw1[nt] = 0;
      if (warp_active && col < n) {
        nw0[nt] = __ldg((const uint32_t*)(prow0 + col));
        nw1[nt] = __ldg((const uint32_t*)(prow1 + col));
      }
    }
  };
  load_pairs(0);

  float o[X::NT2][4];
#pragma unroll
  for (int i = 0; i < X::NT2; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;

  for (int k0 = 0; k0 < n; k0 += kBmKT) {
    const int kt = min(kBmKT, n - k0);
    __syncthreads();
    bm_stage<D>(sK, sKt, qkv + (int64_t)k0 * C3 + C + h * D, C3, kt, 1.f);
    bm_stage<D>(sV, nullptr, qkv + (int64_t)k0 * C3 + 2 * C + h * D, C3, kt, 1.f);
    __syncthreads();
    if (!warp_active) continue;
    for (int c0 = 0; c0 < kt; c0 += kBmChunk) {
      uint32_t pw0[kBmNT], pw1[kBmNT];
#pragma unroll
      for (int nt = 0; nt < kBmNT; ++nt) { pw0[nt] = nw0[nt]; pw1[nt] = nw1[nt]; }
      if (k0 + c0 + kBmChunk < n) load_pairs(k0 + c0 + kBmChunk);   // next chunk's indices: in flight during this chunk's MMAs
      // running dS sums of this chunk: all loads are issued here, ahead of the MMAs (issued one by one at their
      // read-modify-write site they serialise -- a store may alias the next load -- into 8 L2 round trips per chunk)
      float2 acc0[kBmNT], acc1[kBmNT];
      if (dsp != nullptr) {
#pragma unroll
        for (int nt = 0; nt < kBmNT; ++nt) {
          const int col = k0 + c0 + nt * 8 + 2 * tq;
          acc0[nt] = make_float2(0.f, 0.f);
          acc1[nt] = make_float2(0.f, 0.f);
          if (col < n) {
            if (ok0) acc0[nt] = __ldcg((const float2*)(dsp + (int64_t)r0 * ld + col));
            if (ok1) acc1[nt] = __ldcg((const float2*)(dsp + (int64_t)r1 * ld + col));
          }
        }
      }
      float s[kBmNT][4], dp[kBmNT][4];
#pragma unroll
      for (int nt = 0; nt < kBmNT; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
      }
#pragma unroll
      for (int kk = 0; kk < X::KK; ++kk) {
#pragma unroll
        for (int nt = 0; nt < kBmNT; ++nt) {
          const __nv_bfloat16* kp = sK + (c0 + nt * 8 + gq) * X::DS + kk * 16 + 2 * tq;
          mma_bf16_16816(s[nt], aq[kk], *(const uint32_t*)kp, *(const uint32_t*)(kp + 8));
          const __nv_bfloat16* vp = sV + (c0 + nt * 8 + gq) * X::DS + kk * 16 + 2 * tq;
          mma_bf16_16816(dp[nt], ado[kk], *(const uint32_t*)vp, *(const uint32_t*)(vp + 8));
        }
      }
#pragma unroll
      for (int nt = 0; nt < kBmNT; ++nt) {
        const int col = k0 + c0 + nt * 8 + 2 * tq;
        const bool v0 = col < n, v1 = col + 1 < n;
        const float p00 = v0 ? exp2f(s[nt][0] + sLut[pw0[nt] & 0xFFFFu] - lse0) : 0.f;
        const float p01 = v1 ? exp2f(s[nt][1] + sLut[pw0[nt] >> 16] - lse0) : 0.f;
        const float p10 = v0 ? exp2f(s[nt][2] + sLut[pw1[nt] & 0xFFFFu] - lse1) : 0.f;
        const float p11 = v1 ? exp2f(s[nt][3] + sLut[pw1[nt] >> 16] - lse1) : 0.f;
        s[nt][0] = p00 * (dp[nt][0] - dl0);
        s[nt][1] = p01 * (dp[nt][1] - dl0);
        s[nt][2] = p10 * (dp[nt][2] - dl1);
        s[nt][3] = p11 * (dp[nt][3] - dl1);
        if (dsp != nullptr && v0) {                      // running sum of dS over the layers (one owner per element)
          if (ok0)                                       // col even, ld multiple of 16: 8-byte aligned;
            __stcg((float2*)(dsp + (int64_t)r0 * ld + col),                  // s[nt][1] is 0 for an invalid column
                   make_float2(acc0[nt].x + s[nt][0], acc0[nt].y + s[nt][1]));
          if (ok1)
            __stcg((float2*)(dsp + (int64_t)r1 * ld + col),
                   make_float2(acc1[nt].x + s[nt][2], acc1[nt].y + s[nt][3]));
        }
      }
#pragma unroll
      for (int k2 = 0; k2 < kBmK2; ++k2) {
        uint32_t ap[4];
        ap[0] = pack_bf16(s[2 * k2][0], s[2 * k2][1]);
        ap[1] = pack_bf16(s[2 * k2][2], s[2 * k2][3]);
        ap[2] = pack_bf16(s[2 * k2 + 1][0], s[2 * k2 + 1][1]);
        ap[3] = pack_bf16(s[2 * k2 + 1][2], s[2 * k2 + 1][3]);
#pragma unroll
        for (int i = 0; i < X::NT2; ++i) {
          const __nv_bfloat16* kp = sKt + (i * 8 + gq) * X::VS + c0 + k2 * 16 + 2 * tq;
          mma_bf16_16816(o[i], ap, *(const uint32_t*)kp, *(const uint32_t*)(kp + 8));
        }
      }
    }
  }
  if (!warp_active) return;
  __nv_bfloat16* dq = (__nv_bfloat16*)a.d_qkv;
#pragma unroll
  for (int i = 0; i < X::NT2; ++i) {
    const int d = i * 8 + 2 * tq;
    if (d < D) {
      if (ok0) *(uint32_t*)(dq + (int64_t)(n0 + r0) * C3 + h * D + d) = pack_bf16(o[i][0] * scale, o[i][1] * scale);
      if (ok1) *(uint32_t*)(dq + (int64_t)(n0 + r1) * C3 + h * D + d) = pack_bf16(o[i][2] * scale, o[i][3] * scale);
    }
  }
}

template <int D>
__global__ void __launch_bounds__(kBmWarps * 32, 2) attention_bwd_mma_dkv_kernel(const ghn3_attention_bwd_args a) {
  using X = BmDims<D>;
  extern __shared__ __align__(16) uint8_t bm_smem[];
  __nv_bfloat16* sQ = (__nv_bfloat16*)bm_smem;                  // [QT][DS]  q * d^-1/2 * log2(e), rounded as forward
  __nv_bfloat16* sdO = sQ + kBmKT * X::DS;                      // [QT][DS]
  __nv_bfloat16* sQt = sdO + kBmKT * X::DS;                     // [DN][VS]
  __nv_bfloat16* sdOt = sQt + X::DN * X::VS;                    // [DN][VS]
  float* sLse = (float*)(sdOt + X::DN * X::VS);                 // [QT]
  float* sDel = sLse + kBmKT;                                   // [QT]
  float* sLutT = sDel + kBmKT;                                  // [lut_size]: lut with the two SPD digits swapped

  const int g = blockIdx.z, h = blockIdx.y;
  const int n0 = a.node_off[g];
  const int n = a.node_off[g + 1] - n0;
  const int j0 = blockIdx.x * kBmRows;
  if (j0 >= n) return;
  const int ld = (n + 15) & ~15;
  const int C = a.hid, C3 = 3 * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const __nv_bfloat16* qkv = (const __nv_bfloat16*)a.qkv + (int64_t)n0 * C3;
  const __nv_bfloat16* dout = (const __nv_bfloat16*)a.d_out + (int64_t)n0 * C;
  const uint16_t* pair = a.pair + a.mat_off[g];
  const float scale = rsqrtf((float)D);
  int V = 1;
  while (V * V < a.lut_size) ++V;

  pdl_launch_dependents();
  // bias(i, j) seen from row j: pair[j][i] = spd_ji * V + spd_ij -> lut[spd_ij * V + spd_ji]
  for (int i = threadIdx.x; i < a.lut_size; i += blockDim.x) {
    const int hi = i / V, lo = i - hi * V;
    sLutT[i] = a.lut[(int64_t)h * a.lut_size + lo * V + hi] * kLog2e;
  }
  pdl_wait();

  const int r0 = j0 + warp * 16 + gq, r1 = r0 + 8;
  const bool ok0 = r0 < n, ok1 = r1 < n;
  const bool warp_active = (j0 + warp * 16) < n;
  uint32_t ak[X::KK][4], av[X::KK][4];
  bm_row_frags<D>(ak, qkv + C + h * D, C3, r0, r1, ok0, ok1, tq, 1.f);
  bm_row_frags<D>(av, qkv + 2 * C + h * D, C3, r0, r1, ok0, ok1, tq, 1.f);
  const uint16_t* prow0 = pair + (int64_t)(ok0 ? r0 : j0) * ld;
  const uint16_t* prow1 = pair + (int64_t)(ok1 ? r1 : j0) * ld;
  uint32_t nw0[kBmNT], nw1[kBmNT];
  auto load_pairs = [&](int cbase) {
#pragma unroll
    for (int nt = 0; nt < kBmNT; ++nt) {
      const int col = cbase + nt * 8 + 2 * tq;
      nw0[nt] = 0; nw1[nt] = 0;
      if (warp_active && col < n) {
        nw0[nt] = __ldg((const uint32_t*)(prow0 + col));
        nw1[nt] = __ldg((const uint32_t*)(prow1 + col));
      }
    }
  };
  load_pairs(0);

  float ok_[X::NT2][4], ov[X::NT2][4];
#pragma unroll
  for (int i = 0; i < X::NT2; ++i) {
    ok_[i][0] = ok_[i][1] = ok_[i][2] = ok_[i][3] = 0.f;
    ov[i][0] = ov[i][1] = ov[i][2] = ov[i][3] = 0.f;
  }

  for (int i0 = 0; i0 < n; i0 += kBmKT) {
    const int it = min(kBmKT, n - i0);
    __syncthreads();
    bm_stage<D>(sQ, sQt, qkv + (int64_t)i0 * C3 + h * D, C3, it, scale * kLog2e);
    bm_stage<D>(sdO, sdOt, dout + (int64_t)i0 * C + h * D, C, it, 1.f);
    for (int i = threadIdx.x; i < kBmKT; i += blockDim.x) {
      const bool v = i < it;
      sLse[i] = v ? a.fwd_lse2[(int64_t)h * a.total_nodes + n0 + i0 + i] : 0.f;
      sDel[i] = v ? a.delta[(int64_t)h * a.total_nodes + n0 + i0 + i] : 0.f;
    }
    __syncthreads();
    if (!warp_active) continue;
    for (int c0 = 0; c0 < it; c0 += kBmChunk) {
      uint32_t pw0[kBmNT], pw1[kBmNT];
#pragma unroll
      for (int nt = 0; nt < kBmNT; ++nt) { pw0[nt] = nw0[nt]; pw1[nt] = nw1[nt]; }
      if (i0 + c0 + kBmChunk < n) load_pairs(i0 + c0 + kBmChunk);
      float s[kBmNT][4], dp[kBmNT][4];
#pragma unroll
      for (int nt = 0; nt < kBmNT; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
      }
#pragma unroll
      for (int kk = 0; kk < X::KK; ++kk) {
#pragma unroll
        for (int nt = 0; nt < kBmNT; ++nt) {
          const __nv_bfloat16* qp = sQ + (c0 + nt * 8 + gq) * X::DS + kk * 16 + 2 * tq;
          mma_bf16_16816(s[nt], ak[kk], *(const uint32_t*)qp, *(const uint32_t*)(qp + 8));
          const __nv_bfloat16* gp = sdO + (c0 + nt * 8 + gq) * X::DS + kk * 16 + 2 * tq;
          mma_bf16_16816(dp[nt], av[kk], *(const uint32_t*)gp, *(const uint32_t*)(gp + 8));
        }
      }
      float p[kBmNT][4];
#pragma unroll
      for (int nt = 0; nt < kBmNT; ++nt) {
        const int lc = c0 + nt * 8 + 2 * tq;             // column inside the staged tile
        const bool v0 = i0 + lc < n, v1 = i0 + lc + 1 < n;
        const float le0 = sLse[lc], le1 = sLse[lc + 1], de0 = sDel[lc], de1 = sDel[lc + 1];
        p[nt][0] = v0 ? exp2f(s[nt][0] + sLutT[pw0[nt] & 0xFFFFu] - le0) : 0.f;
        p[nt][1] = v1 ? exp2f(s[nt][1] + sLutT[pw0[nt] >> 16] - le1) : 0.f;
        p[nt][2] = v0 ? exp2f(s[nt][2] + sLutT[pw1[nt] & 0xFFFFu] - le0) : 0.f;
        p[nt][3] = v1 ? exp2f(s[nt][3] + sLutT[pw1[nt] >> 16] - le1) : 0.f;
        s[nt][0] = p[nt][0] * (dp[nt][0] - de0);
        s[nt][1] = p[nt][1] * (dp[nt][1] - de1);
        s[nt][2] = p[nt][2] * (dp[nt][2] - de0);
        s[nt][3] = p[nt][3] * (dp[nt][3] - de1);
      }
#pragma unroll
      for (int k2 = 0; k2 < kBmK2; ++k2) {
        uint32_t ap[4], as_[4];
        ap[0] = pack_bf16(p[2 * k2][0], p[2 * k2][1]);
        ap[1] = pack_bf16(p[2 * k2][2], p[2 * k2][3]);
        ap[2] = pack_bf16(p[2 * k2 + 1][0], p[2 * k2 + 1][1]);
        ap[3] = pack_bf16(p[2 * k2 + 1][2], p[2 * k2 + 1][3]);
        as_[0] = pack_bf16(s[2 * k2][0], s[2 * k2][1]);
        as_[1] = pack_bf16(s[2 * k2][2], s[2 * k2][3]);
        as_[2] = pack_bf16(s[2 * k2 + 1][0], s[2 * k2 + 1][1]);
        as_[3] = pack_bf16(s[2 * k2 + 1][2], s[2 * k2 + 1][3]);
#pragma unroll
        for (int i = 0; i < X::NT2; ++i) {
          const __nv_bfloat16* gp = sdOt + (i * 8 + gq) * X::VS + c0 + k2 * 16 + 2 * tq;
          mma_bf16_16816(ov[i], ap, *(const uint32_t*)gp, *(const uint32_t*)(gp + 8));
          const __nv_bfloat16* qp = sQt + (i * 8 + gq) * X::VS + c0 + k2 * 16 + 2 * tq;
          mma_bf16_16816(ok_[i], as_, *(const uint32_t*)qp, *(const uint32_t*)(qp + 8));
        }
      }
    }
  }
  if (!warp_active) return;
  __nv_bfloat16* dst = (__nv_bfloat16*)a.d_qkv;
#pragma unroll
  for (int i = 0; i < X::NT2; ++i) {
    const int d = i * 8 + 2 * tq;
    if (d < D) {
      // sQt carries d^-1/2 * log2(e): undo the log2(e)
      if (ok0) {
        *(uint32_t*)(dst + (int64_t)(n0 + r0) * C3 + C + h * D + d) = pack_bf16(ok_[i][0] * kLn2, ok_[i][1] * kLn2);
        *(uint32_t*)(dst + (int64_t)(n0 + r0) * C3 + 2 * C + h * D + d) = pack_bf16(ov[i][0], ov[i][1]);
      }
      if (ok1) {
        *(uint32_t*)(dst + (int64_t)(n0 + r1) * C3 + C + h * D + d) = pack_bf16(ok_[i][2] * kLn2, ok_[i][3] * kLn2);
        *(uint32_t*)(dst + (int64_t)(n0 + r1) * C3 + 2 * C + h * D + d) = pack_bf16(ov[i][2], ov[i][3]);
      }
    }
  }
}

// d_lut[h][pair[i][j]] += ds_total[g][h][i][j]: one CTA per (row chunk, head, graph), shared-memory histogram
__global__ void __launch_bounds__(256) lut_bin_kernel(const ghn3_lut_bin_args a) {
  extern __shared__ float lb_hist[];
  pdl_launch_dependents();
  pdl_wait();
  const int g = blockIdx.z, h = blockIdx.y;
  const int n0 = a.node_off[g];
  const int n = a.node_off[g + 1] - n0;
  const int i0 = blockIdx.x * 32;
  if (i0 >= n) return;
  const int ld = (n + 15) & ~15;
  for (int i = threadIdx.x; i < a.lut_size; i += blockDim.x) lb_hist[i] = 0.f;
  __syncthreads();
  const uint16_t* pair = a.pair + a.mat_off[g];
  const float* ds = a.ds_total + a.mat_off[g] * a.heads + (int64_t)h * n * ld;
  const int rows = min(32, n - i0);
  for (int idx = threadIdx.x; idx < rows * ld; idx += blockDim.x) {
    const int r = idx / ld, j = idx - r * ld;
    if (j < n) {
      const float v = ds[(int64_t)(i0 + r) * ld + j];
      if (v != 0.f) atomicAdd(lb_hist + pair[(int64_t)(i0 + r) * ld + j], v);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < a.lut_size; i += blockDim.x)
    if (lb_hist[i] != 0.f) atomicAdd(a.d_lut + (int64_t)h * a.lut_size + i, lb_hist[i]);
}

template <int D>
static int launch_bm(const ghn3_attention_bwd_args* a, cudaStream_t stream) {
  using X = BmDims<D>;
  const size_t smem_a = sizeof(__nv_bfloat16) * (2 * kBmKT * X::DS + X::DN * X::VS) + sizeof(float) * a->lut_size;
  const size_t smem_b = sizeof(__nv_bfloat16) * (2 * kBmKT * X::DS + 2 * X::DN * X::VS) +
                        sizeof(float) * (2 * kBmKT + a->lut_size);
  GHN3_REQUIRE(smem_a <= 200 * 1024 && smem_b <= 200 * 1024, "ghn3_attention_bwd: look-up table too large");
  GHN3_CUDA(cudaFuncSetAttribute(attention_bwd_mma_dq_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
  GHN3_CUDA(cudaFuncSetAttribute(attention_bwd_mma_dkv_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
  const dim3 grid((unsigned)ceil_div(a->max_nodes, kBmRows), (unsigned)a->heads, (unsigned)a->n_graphs);
  GHN3_CUDA(launch_pdl(attention_bwd_mma_dq_kernel<D>, grid, dim3(kBmWarps * 32), smem_a, stream, *a));
  GHN3_LAUNCH_CHECK("attention_bwd_mma_dq_kernel");
  GHN3_CUDA(launch_pdl(attention_bwd_mma_dkv_kernel<D>, grid, dim3(kBmWarps * 32), smem_b, stream, *a));
  GHN3_LAUNCH_CHECK("attention_bwd_mma_dkv_kernel");
  return GHN3_OK;
}

// tensor-core path of ghn3_attention_bwd (bf16 storage, statistics from the forward, dS accumulated for ghn3_lut_bin)
int attention_bwd_mma_impl(const ghn3_attention_bwd_args* a, cudaStream_t stream) {
  const int D = a->hid / a->heads;
  GHN3_REQUIRE(D % 2 == 0, "ghn3_attention_bwd: odd head dimension");
#define GHN3_BM_CASE(DV) if (D == DV) return launch_bm<DV>(a, stream);
  GHN3_BM_CASE(4)
  GHN3_BM_CASE(8)
  GHN3_BM_CASE(16)
  GHN3_BM_CASE(24)
  GHN3_BM_CASE(32)
#undef GHN3_BM_CASE
  set_error("ghn3_attention_bwd: head dim %d is not supported (4, 8, 16, 24, 32)", D);
  return GHN3_ERR_UNSUPPORTED;
}

}  // namespace ghn3

using namespace ghn3;

extern "C" int ghn3_lut_bin(const ghn3_lut_bin_args* a, ghn3_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHN3_REQUIRE(a != nullptr && a->ds_total && a->d_lut && a->pair && a->node_off && a->mat_off, "ghn3_lut_bin: null args");
  if (a->n_graphs <= 0 || a->max_nodes <= 0) return GHN3_OK;
  const dim3 grid((unsigned)ceil_div(a->max_nodes, 32), (unsigned)a->heads, (unsigned)a->n_graphs);
  GHN3_CUDA(launch_pdl(lut_bin_kernel, grid, dim3(256), sizeof(float) * a->lut_size, stream, *a));
  GHN3_LAUNCH_CHECK("lut_bin_kernel");
  return GHN3_OK;
}
