// Attention forward for fp32 storage (the 'tf32' compute modes) on the tensor cores: every fp32 operand is split into
// two bf16 terms (hi = bf16(v), lo = bf16(v - hi), together ~16 mantissa bits) and each product is evaluated as
// hi*hi + hi*lo + lo*hi with fp32 accumulation -- the bf16 analogue of the 3-term tf32 GEMM mode. Same structure and
// register layout as attention_mma_kernel (dense_kernels.cu); replaces the CUDA-core fp32 kernel on the hot path
// (ghn3/graphormer.py:121-140). Error of a logit ~1e-5 relative, far inside the 1e-3 budget of the tf32 mode.
#include "common.cuh"

namespace ghn3 {

constexpr int kSpKT = 128;       // keys staged per outer iteration
constexpr int kSpWarps = 8;
constexpr int kSpQT = 16 * kSpWarps;
constexpr float kSpLog2e = 1.44269504088896340736f;

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
__device__ __forceinline__ void split_pack(float a, float b, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat16 ah, al, bh, bl;
  split_bf16(a, ah, al);
  split_bf16(b, bh, bl);
  __nv_bfloat162 h = __halves2bfloat162(ah, bh), l = __halves2bfloat162(al, bl);
  hi = *(uint32_t*)&h;
  lo = *(uint32_t*)&l;
}

template <int D>
__global__ void __launch_bounds__(kSpWarps * 32) attention_split_kernel(const ghn3_attention_args a) {
  constexpr int DK = (D + 15) / 16 * 16;
  constexpr int DS = DK + 8;
  constexpr int DN = (D + 7) / 8 * 8;
  constexpr int VS = kSpKT + 8;
  constexpr int NT2 = DN / 8;
  constexpr int KK = DK / 16;
  extern __shared__ __align__(16) uint8_t sp_smem[];
  __nv_bfloat16* sKh = (__nv_bfloat16*)sp_smem;                 // [KT][DS]
  __nv_bfloat16* sKl = sKh + kSpKT * DS;
  __nv_bfloat16* sVh = sKl + kSpKT * DS;                        // [DN][VS]  (V^T)
  __nv_bfloat16* sVl = sVh + DN * VS;
  float* sLut = (float*)(sVl + DN * VS);

  const int g = blockIdx.z, h = blockIdx.y;
  const int n0 = a.node_off[g];
  const int n = a.node_off[g + 1] - n0;
  const int q0 = blockIdx.x * kSpQT;
  if (q0 >= n) return;
  const int ld = (n + 15) & ~15;
  const int C = a.hid, C3 = 3 * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const float* qkv = (const float*)a.qkv + (int64_t)n0 * C3;
  const uint16_t* pair = a.pair + a.mat_off[g];
  const float scale_log2 = rsqrtf((float)D) * kSpLog2e;

  pdl_launch_dependents();
  for (int i = threadIdx.x; i < a.lut_size; i += blockDim.x) sLut[i] = __ldg(a.lut + (int64_t)h * a.lut_size + i) * kSpLog2e;
  pdl_wait();

  const int r0 = q0 + warp * 16 + gq, r1 = r0 + 8;
  const bool ok0 = r0 < n, ok1 = r1 < n;
  const bool warp_active = (q0 + warp * 16) < n;
  uint32_t qh[KK][4], ql[KK][4];
#pragma unroll
  for (int kk = 0; kk < KK; ++kk) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int d = kk * 16 + half * 8 + 2 * tq;
      float v00 = 0.f, v01 = 0.f, v10 = 0.f, v11 = 0.f;
      if (d < D) {
        if (ok0) {
          const float2 t = *(const float2*)(qkv + (int64_t)r0 * C3 + h * D + d);
          v00 = t.x * scale_log2; v01 = t.y * scale_log2;
        }
        if (ok1) {
          const float2 t = *(const float2*)(qkv + (int64_t)r1 * C3 + h * D + d);
          v10 = t.x * scale_log2; v11 = t.y * scale_log2;
        }
      }
      split_pack(v00, v01, qh[kk][half * 2 + 0], ql[kk][half * 2 + 0]);
      split_pack(v10, v11, qh[kk][half * 2 + 1], ql[kk][half * 2 + 1]);
    }
  }
  const uint16_t* prow0 = pair + (int64_t)(ok0 ? r0 : q0) * ld;
  const uint16_t* prow1 = pair + (int64_t)(ok1 ? r1 : q0) * ld;

  float o[NT2][4];
#pragma unroll
  for (int i = 0; i < NT2; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int k0 = 0; k0 < n; k0 += kSpKT) {
    const int kt = min(kSpKT, n - k0);
    const int kt32 = (kt + 31) & ~31;
    __syncthreads();
    // stage K (hi / lo, key-major) and V^T (hi / lo, dim-major); zero padding for rows up to the chunk boundary and
    // for the padded dims
    for (int idx = threadIdx.x; idx < kt32 * (DK / 2); idx += blockDim.x) {
      const int j = idx / (DK / 2), d = (idx - j * (DK / 2)) * 2;
      float k0v = 0.f, k1v = 0.f, v0v = 0.f, v1v = 0.f;
      if (j < kt && d < D) {
        const float* row = qkv + (int64_t)(k0 + j) * C3 + h * D + d;
        const float2 kk2 = *(const float2*)(row + C);
        const float2 vv2 = *(const float2*)(row + 2 * C);
        k0v = kk2.x; k1v = kk2.y; v0v = vv2.x; v1v = vv2.y;
      }
      uint32_t hi, lo;
      split_pack(k0v, k1v, hi, lo);
      *(uint32_t*)(sKh + j * DS + d) = hi;
      *(uint32_t*)(sKl + j * DS + d) = lo;
      if (d < DN) {
        __nv_bfloat16 a0, b0, a1, b1;
        split_bf16(v0v, a0, b0);
        split_bf16(v1v, a1, b1);
        sVh[d * VS + j] = a0; sVl[d * VS + j] = b0;
        sVh[(d + 1) * VS + j] = a1; sVl[(d + 1) * VS + j] = b1;
      }
    }
    __syncthreads();
    if (!warp_active) continue;
    for (int c0 = 0; c0 < kt; c0 += 32) {
      uint32_t pw0[4], pw1[4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int col = k0 + c0 + nt * 8 + 2 * tq;
        pw0[nt] = 0; pw1[nt] = 0;
        if (col < n) {
          pw0[nt] = __ldg((const uint32_t*)(prow0 + col));
          pw1[nt] = __ldg((const uint32_t*)(prow1 + col));
        }
      }
      float s[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < KK; ++kk) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int off = (c0 + nt * 8 + gq) * DS + kk * 16 + 2 * tq;
          const uint32_t kh0 = *(const uint32_t*)(sKh + off), kh1 = *(const uint32_t*)(sKh + off + 8);
          const uint32_t kl0 = *(const uint32_t*)(sKl + off), kl1 = *(const uint32_t*)(sKl + off + 8);
          mma_bf16_16816(s[nt], ql[kk], kh0, kh1);
          mma_bf16_16816(s[nt], qh[kk], kl0, kl1);
          mma_bf16_16816(s[nt], qh[kk], kh0, kh1);
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int col = k0 + c0 + nt * 8 + 2 * tq;
        const bool v0 = col < n, v1 = col + 1 < n;
        s[nt][0] = v0 ? s[nt][0] + sLut[pw0[nt] & 0xFFFFu] : -INFINITY;
        s[nt][1] = v1 ? s[nt][1] + sLut[pw0[nt] >> 16] : -INFINITY;
        s[nt][2] = v0 ? s[nt][2] + sLut[pw1[nt] & 0xFFFFu] : -INFINITY;
        s[nt][3] = v1 ? s[nt][3] + sLut[pw1[nt] >> 16] : -INFINITY;
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
      for (int nt = 0; nt < 4; ++nt) {
        s[nt][0] = exp2f(s[nt][0] - mn0); s[nt][1] = exp2f(s[nt][1] - mn0);
        s[nt][2] = exp2f(s[nt][2] - mn1); s[nt][3] = exp2f(s[nt][3] - mn1);
        l0 += s[nt][0] + s[nt][1];
        l1 += s[nt][2] + s[nt][3];
      }
#pragma unroll
      for (int k2 = 0; k2 < 2; ++k2) {
        uint32_t ph[4], pl[4];
        split_pack(s[2 * k2][0], s[2 * k2][1], ph[0], pl[0]);
        split_pack(s[2 * k2][2], s[2 * k2][3], ph[1], pl[1]);
        split_pack(s[2 * k2 + 1][0], s[2 * k2 + 1][1], ph[2], pl[2]);
        split_pack(s[2 * k2 + 1][2], s[2 * k2 + 1][3], ph[3], pl[3]);
#pragma unroll
        for (int i = 0; i < NT2; ++i) {
          const int off = (i * 8 + gq) * VS + c0 + k2 * 16 + 2 * tq;
          const uint32_t vh0 = *(const uint32_t*)(sVh + off), vh1 = *(const uint32_t*)(sVh + off + 8);
          const uint32_t vl0 = *(const uint32_t*)(sVl + off), vl1 = *(const uint32_t*)(sVl + off + 8);
          mma_bf16_16816(o[i], pl, vh0, vh1);
          mma_bf16_16816(o[i], ph, vl0, vl1);
          mma_bf16_16816(o[i], ph, vh0, vh1);
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
  float* out = (float*)a.out;
  const bool tf = a.dtype == GHN3_TF32;
#pragma unroll
  for (int i = 0; i < NT2; ++i) {
    const int d = i * 8 + 2 * tq;
    if (d < D) {
      float2 y0 = make_float2(o[i][0] * i0, o[i][1] * i0), y1 = make_float2(o[i][2] * i1, o[i][3] * i1);
      if (tf) { y0.x = round_tf32(y0.x); y0.y = round_tf32(y0.y); y1.x = round_tf32(y1.x); y1.y = round_tf32(y1.y); }
      if (ok0) *(float2*)(out + (int64_t)(n0 + r0) * C + h * D + d) = y0;
      if (ok1) *(float2*)(out + (int64_t)(n0 + r1) * C + h * D + d) = y1;
    }
  }
}

template <int D>
static int launch_split(const ghn3_attention_args* a, cudaStream_t stream) {
  constexpr int DK = (D + 15) / 16 * 16, DN = (D + 7) / 8 * 8;
  const size_t smem = sizeof(__nv_bfloat16) * (2 * kSpKT * (DK + 8) + 2 * DN * (kSpKT + 8)) + sizeof(float) * a->lut_size;
  GHN3_REQUIRE(smem <= 200 * 1024, "ghn3_attention: look-up table too large for shared memory");
  GHN3_CUDA(cudaFuncSetAttribute(attention_split_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const dim3 grid((unsigned)ceil_div(a->max_nodes, kSpQT), (unsigned)a->heads, (unsigned)a->n_graphs);
  GHN3_CUDA(launch_pdl(attention_split_kernel<D>, grid, dim3(kSpWarps * 32), smem, stream, *a));
  GHN3_LAUNCH_CHECK("attention_split_kernel");
  return GHN3_OK;
}

// fp32-storage attention on the tensor cores (split-bf16); returns GHN3_ERR_UNSUPPORTED for head dims it does not cover
int attention_split_impl(const ghn3_attention_args* a, cudaStream_t stream) {
  const int D = a->hid / a->heads;
#define GHN3_SP_CASE(DV) if (D == DV) return launch_split<DV>(a, stream);
  GHN3_SP_CASE(8)
  GHN3_SP_CASE(16)
  GHN3_SP_CASE(24)
  GHN3_SP_CASE(32)
#undef GHN3_SP_CASE
  return GHN3_ERR_UNSUPPORTED;
}

}  // namespace ghn3
