// Tile / slice / normalise / scatter of the decoders' max-shape predictions straight into the target model's
// parameter storage -- replaces the per-tensor Python loop of the reference:
//   _tile_params  ghn3/nn.py:422-506   (channel crop + repeat, centred spatial window, pos-enc transpose)
//   _normalize    ghn3/nn.py:554-592   (fan-in scale | 2*sigmoid(x/2) | tanh(x/5))
//   _set_params   ghn3/nn.py:508-552   (param.data = tensor.clone())
// One launch per model batch; one CTA per GHN3_SCATTER_CHUNK-element chunk of one target tensor; the CTA finds its tensor by a
// binary search over the chunk prefix in the descriptor table. HBM-write bound: the predictions (<= 150 MB) stay in
// L2 while every target byte is written exactly once with 16-byte coalesced stores.
#include "common.cuh"

namespace ghn3 {

__device__ __forceinline__ float finish(float v, int mode, float scale) {
  // fp32 op order of the reference: p * scale | 2 * sigmoid(0.5 * p) | tanh(0.2 * p)
  if (mode == 1) return 2.0f * (1.0f / (1.0f + expf(-(0.5f * v))));
  if (mode == 2) return tanhf(0.2f * v);
  return v * scale;
}

// exact n / d for n < 2^31 with a host-computed multiplier: d == 1 -> mul == 0; else
// s = ceil(log2 d), mul = ceil(2^(31+s) / d), n / d = umulhi(n, mul) >> (s - 1)
__device__ __forceinline__ uint32_t fdiv(uint32_t n, uint32_t mul, uint32_t sh) {
  return mul ? (__umulhi(n, mul) >> sh) : n;
}

__device__ __forceinline__ float fetch_bilinear(const ghn3_scatter_desc& d, int a, int64_t col, int y, int x) {
  // bilinear resize (align_corners=False) of the kh_src x kw_src window to (t2, t3): nn.py:751-753
  const float sy = fmaxf(((float)y + 0.5f) * ((float)d.kh_src / (float)d.t2) - 0.5f, 0.f);
  const float sx = fmaxf(((float)x + 0.5f) * ((float)d.kw_src / (float)d.t3) - 0.5f, 0.f);
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = min(y0 + 1, d.kh_src - 1), x1 = min(x0 + 1, d.kw_src - 1);
  const float ly = sy - (float)y0, lx = sx - (float)x0;
  const int64_t base = (int64_t)a * d.ra;
  const float v00 = d.src[(base + y0 * d.kw_src + x0) * d.ld + col];
  const float v01 = d.src[(base + y0 * d.kw_src + x1) * d.ld + col];
  const float v10 = d.src[(base + y1 * d.kw_src + x0) * d.ld + col];
  const float v11 = d.src[(base + y1 * d.kw_src + x1) * d.ld + col];
  return (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
}

// One CTA per GHN3_SCATTER_CHUNK-element chunk of one target tensor; a thread handles 4 consecutive target elements
// per step (one 16-byte store). 16384-element chunks: with 8192 the isolated kernel is as fast, but the twice as many
// CTAs interfere much more with the next call's Graphormer chain in the overlapped mode (1.65 -> 1.47 ms per step). All index arithmetic is 32-bit with multiply-high divisions; the descriptor lives in
// registers.
__device__ __forceinline__ void block_sumsq(float acc, double* out) {
  __shared__ float part[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i];
    atomicAdd(out, (double)t);
  }
}

__global__ void __launch_bounds__(256) scatter_kernel(const ghn3_scatter_desc* __restrict__ descs, int n_descs,
                                                      const int32_t* __restrict__ chunk_desc,
                                                      double* __restrict__ norm_out) {
  int di;
  if (chunk_desc != nullptr) {
    di = __ldg(chunk_desc + blockIdx.x);               // host-built chunk -> descriptor table: one load
  } else {
    __shared__ int s_desc;
    if (threadIdx.x == 0) {
      const int64_t chunk = blockIdx.x;
      int lo = 0, hi = n_descs;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (descs[mid].chunk0 <= chunk) lo = mid; else hi = mid;
      }
      s_desc = lo;
    }
    __syncthreads();
    di = s_desc;
  }
  const ghn3_scatter_desc d = descs[di];               // uniform, served by L1/L2; kept in registers
  const uint32_t base = (uint32_t)(((int64_t)blockIdx.x - d.chunk0) * GHN3_SCATTER_CHUNK);
  const uint32_t end = (uint32_t)min((int64_t)base + GHN3_SCATTER_CHUNK, d.numel);
  const bool dst_vec = ((((uintptr_t)d.dst) & 15) == 0);
  const int hw = d.t2 * d.t3;
  const int mode = d.mode;
  const float scale = d.scale;
  const bool want_norm = norm_out != nullptr && d.norm_slot >= 0;
  float sq = 0.f;

  if (hw == 1 && mode != 3) {
    // matrices and vectors: row `a` = e / t1, column b = e % t1
    const bool src_vec = (d.si % 4 == 0) && (d.ca % 4 == 0) && (d.ld % 4 == 0) && ((((uintptr_t)d.src) & 15) == 0);
    const int64_t row_off = ((int64_t)d.cy * d.kw_src + d.cx) * d.ld;
#pragma unroll 4
    for (int k = 0; k < GHN3_SCATTER_CHUNK / 1024; ++k) {
      const uint32_t e = base + threadIdx.x * 4 + k * 1024;
      if (e >= end) break;
      uint32_t a = fdiv(e, d.m_t1, d.s_t1);
      uint32_t b = e - a * d.t1;
      uint32_t am = a - fdiv(a, d.m_so, d.s_so) * d.so;
      uint32_t bm = b - fdiv(b, d.m_si, d.s_si) * d.si;
      const float* srow = d.src + row_off + (int64_t)a * d.ra * d.ld;
      if (e + 4 <= end && dst_vec && b + 4 <= (uint32_t)d.t1 && bm + 4 <= (uint32_t)d.si) {
        float4 v;
        const float* sp = srow + (int64_t)am * d.ca + bm;
        if (src_vec && (bm & 3) == 0) {
          v = __ldg((const float4*)sp);
        } else {
          v = make_float4(__ldg(sp), __ldg(sp + 1), __ldg(sp + 2), __ldg(sp + 3));
        }
        float4 o;
        o.x = finish(v.x, mode, scale); o.y = finish(v.y, mode, scale);
        o.z = finish(v.z, mode, scale); o.w = finish(v.w, mode, scale);
        sq += (o.x * o.x + o.y * o.y) + (o.z * o.z + o.w * o.w);
        __stcs((float4*)(d.dst + e), o);
      } else {
        float o[4];
        const int cnt = (int)min(4u, end - e);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i < cnt) {
            o[i] = finish(__ldg(srow + (int64_t)am * d.ca + bm), mode, scale);
            if (++bm == (uint32_t)d.si) bm = 0;
            if (++b == (uint32_t)d.t1) {
              b = 0; bm = 0; ++a;
              if (++am == (uint32_t)d.so) am = 0;
              srow += (int64_t)d.ra * d.ld;
            }
          }
        }
        for (int i = 0; i < cnt; ++i) sq += o[i] * o[i];
        if (cnt == 4 && dst_vec) __stcs((float4*)(d.dst + e), make_float4(o[0], o[1], o[2], o[3]));
        else for (int i = 0; i < cnt; ++i) d.dst[e + i] = o[i];
      }
    }
    if (want_norm) block_sumsq(sq, norm_out + d.norm_slot);
    return;
  }

#pragma unroll 4
  for (int k = 0; k < GHN3_SCATTER_CHUNK / 1024; ++k) {
    const uint32_t e = base + threadIdx.x * 4 + k * 1024;
    if (e >= end) break;
    // decompose the first element, then step with carries
    uint32_t r = fdiv(e, d.m_t3, d.s_t3);
    int x = (int)(e - r * d.t3);
    uint32_t r2 = fdiv(r, d.m_t2, d.s_t2);
    int y = (int)(r - r2 * d.t2);
    uint32_t a = fdiv(r2, d.m_t1, d.s_t1);
    uint32_t b = r2 - a * d.t1;
    uint32_t am = a - fdiv(a, d.m_so, d.s_so) * d.so;
    uint32_t bm = b - fdiv(b, d.m_si, d.s_si) * d.si;
    float o[4];
    const int cnt = (int)min(4u, end - e);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i < cnt) {
        const int64_t col = (int64_t)am * d.ca + bm;
        float v;
        if (mode == 3) {
          v = fetch_bilinear(d, (int)a, col, y, x);
        } else {
          const int64_t row = (int64_t)a * d.ra + (int64_t)(y + d.cy) * d.kw_src + (x + d.cx);
          v = __ldg(d.src + row * d.ld + col);
        }
        o[i] = finish(v, mode, scale);
        if (++x == d.t3) {
          x = 0;
          if (++y == d.t2) {
            y = 0;
            if (++bm == (uint32_t)d.si) bm = 0;
            if (++b == (uint32_t)d.t1) {
              b = 0; bm = 0; ++a;
              if (++am == (uint32_t)d.so) am = 0;
            }
          }
        }
      }
    }
    for (int i = 0; i < cnt; ++i) sq += o[i] * o[i];
    if (cnt == 4 && dst_vec) __stcs((float4*)(d.dst + e), make_float4(o[0], o[1], o[2], o[3]));
    else for (int i = 0; i < cnt; ++i) d.dst[e + i] = o[i];
  }
  if (want_norm) block_sumsq(sq, norm_out + d.norm_slot);
}

// sum of squares over a list of tensors: grid.y = tensor, grid.x strides over its elements
__global__ void __launch_bounds__(256) sumsq_kernel(const float* const* __restrict__ ptrs,
                                                    const int64_t* __restrict__ numels, double* __restrict__ out) {
  const float* p = ptrs[blockIdx.y];
  const int64_t n = numels[blockIdx.y];
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = p[i];
    acc = fmaf(v, v, acc);
  }
  double dacc = (double)acc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dacc += __shfl_xor_sync(0xffffffffu, dacc, o);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = dacc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < 8; ++i) t += part[i];
    if (t != 0) atomicAdd(out, t);
  }
}

}  // namespace ghn3

using namespace ghn3;

extern "C" int ghn3_scatter(const ghn3_scatter_args* a, ghn3_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHN3_REQUIRE(a != nullptr, "ghn3_scatter: null args");
  static_assert(sizeof(ghn3_scatter_desc) == 136, "descriptor layout is part of the ABI");
  if (a->n_descs <= 0 || a->n_chunks <= 0) return GHN3_OK;
  GHN3_REQUIRE(a->n_chunks < (int64_t)2147483647, "ghn3_scatter: too many chunks");
  if (a->norm_out != nullptr && a->n_norm_slots > 0)
    GHN3_CUDA(cudaMemsetAsync(a->norm_out, 0, sizeof(double) * a->n_norm_slots, stream));
  scatter_kernel<<<(unsigned)a->n_chunks, 256, 0, stream>>>(a->descs, a->n_descs, a->chunk_desc, a->norm_out);
  GHN3_LAUNCH_CHECK("scatter_kernel");
  return GHN3_OK;
}

extern "C" int ghn3_sumsq(const ghn3_sumsq_args* a, ghn3_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHN3_REQUIRE(a != nullptr && a->out != nullptr, "ghn3_sumsq: null args");
  GHN3_CUDA(cudaMemsetAsync(a->out, 0, sizeof(double), stream));
  if (a->n <= 0) return GHN3_OK;
  sumsq_kernel<<<dim3(64, (unsigned)a->n), 256, 0, stream>>>(a->ptrs, a->numels, a->out);
  GHN3_LAUNCH_CHECK("sumsq_kernel");
  return GHN3_OK;
}
