// Tile / slice / normalise / scatter of the decoders' max-shape predictions straight into the target model's
// parameter storage -- replaces the per-tensor Python loop of the reference:
//   _tile_params  ghn3/nn.py:422-506   (channel crop + repeat, centred spatial window, pos-enc transpose)
//   _normalize    ghn3/nn.py:554-592   (fan-in scale | 2*sigmoid(x/2) | tanh(x/5))
//   _set_params   ghn3/nn.py:508-552   (param.data = tensor.clone())
// One launch per model batch; one CTA per 4096-element chunk of one target tensor; the CTA finds its tensor by a
// binary search over the chunk prefix in the descriptor table. HBM-write bound: the predictions (<= 150 MB) stay in
// L2 while every target byte is written exactly once with 16-byte coalesced stores.
#include "common.cuh"

namespace ghn3 {

__device__ __forceinline__ float finish(float v, const ghn3_scatter_desc& d) {
  // fp32 op order of the reference: p * scale | 2 * sigmoid(0.5 * p) | tanh(0.2 * p)
  if (d.mode == 1) return 2.0f * (1.0f / (1.0f + expf(-(0.5f * v))));
  if (d.mode == 2) return tanhf(0.2f * v);
  return v * d.scale;
}

__device__ __forceinline__ float fetch(const ghn3_scatter_desc& d, int a, int b, int y, int x) {
  const int am = a % d.so, bm = b % d.si;
  const int64_t col = (int64_t)am * d.ca + bm;
  if (d.mode == 3) {
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
  const int64_t row = (int64_t)a * d.ra + (int64_t)(y + d.cy) * d.kw_src + (x + d.cx);
  return d.src[row * d.ld + col];
}

__global__ void __launch_bounds__(256) scatter_kernel(const ghn3_scatter_desc* __restrict__ descs, int n_descs) {
  __shared__ ghn3_scatter_desc sd;
  if (threadIdx.x == 0) {
    const int64_t chunk = blockIdx.x;
    int lo = 0, hi = n_descs;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (descs[mid].chunk0 <= chunk) lo = mid; else hi = mid;
    }
    sd = descs[lo];
  }
  __syncthreads();
  const ghn3_scatter_desc& d = sd;
  const int64_t base = ((int64_t)blockIdx.x - d.chunk0) * GHN3_SCATTER_CHUNK;
  const int64_t end = min(base + (int64_t)GHN3_SCATTER_CHUNK, d.numel);
  const int hw = d.t2 * d.t3;
  const bool dst_vec = ((((uintptr_t)d.dst) & 15) == 0);

  if (hw == 1 && d.mode != 3) {
    // matrices and vectors: 4 consecutive elements share the row `a` whenever t1 % 4 == 0
    const bool fast = dst_vec && (d.t1 % 4 == 0) && (d.si % 4 == 0) && (d.ca % 4 == 0) && (d.ld % 4 == 0) &&
                      ((((uintptr_t)d.src) & 15) == 0);
    for (int64_t e = base + threadIdx.x * 4; e < end; e += 1024) {
      if (fast && e + 4 <= end) {
        const int a = (int)(e / d.t1), b = (int)(e - (int64_t)a * d.t1);
        const int am = a % d.so, bm = b % d.si;
        const int64_t row = (int64_t)a * d.ra + (int64_t)d.cy * d.kw_src + d.cx;
        const float4 v = *(const float4*)(d.src + row * d.ld + (int64_t)am * d.ca + bm);
        float4 o;
        o.x = finish(v.x, d); o.y = finish(v.y, d); o.z = finish(v.z, d); o.w = finish(v.w, d);
        *(float4*)(d.dst + e) = o;
      } else {
        for (int64_t i = e; i < min(e + 4, end); ++i) {
          const int a = (int)(i / d.t1), b = (int)(i - (int64_t)a * d.t1);
          d.dst[i] = finish(fetch(d, a, b, 0, 0), d);
        }
      }
    }
    return;
  }

  for (int64_t e = base + threadIdx.x * 4; e < end; e += 1024) {
    // decompose the first element, then step with carries
    int64_t r = e;
    int x = (int)(r % d.t3); r /= d.t3;
    int y = (int)(r % d.t2); r /= d.t2;
    int b = (int)(r % d.t1);
    int a = (int)(r / d.t1);
    float o[4];
    const int cnt = (int)min((int64_t)4, end - e);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i < cnt) {
        o[i] = finish(fetch(d, a, b, y, x), d);
        if (++x == d.t3) { x = 0; if (++y == d.t2) { y = 0; if (++b == d.t1) { b = 0; ++a; } } }
      }
    }
    if (cnt == 4 && dst_vec) {
      *(float4*)(d.dst + e) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
      for (int i = 0; i < cnt; ++i) d.dst[e + i] = o[i];
    }
  }
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
  static_assert(sizeof(ghn3_scatter_desc) == 88, "descriptor layout is part of the ABI");
  if (a->n_descs <= 0 || a->n_chunks <= 0) return GHN3_OK;
  GHN3_REQUIRE(a->n_chunks < (int64_t)2147483647, "ghn3_scatter: too many chunks");
  scatter_kernel<<<(unsigned)a->n_chunks, 256, 0, stream>>>(a->descs, a->n_descs);
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
