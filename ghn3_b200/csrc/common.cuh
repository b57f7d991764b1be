// Shared helpers for the ghn3_b200 CUDA sources (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ghn3_b200.h"

namespace ghn3 {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define GHN3_REQUIRE(cond, ...)                \
  do {                                         \
    if (!(cond)) {                             \
      ghn3::set_error(__VA_ARGS__);            \
      return GHN3_ERR_BAD_ARG;                 \
    }                                          \
  } while (0)

#define GHN3_CUDA(expr)                                                                      \
  do {                                                                                       \
    cudaError_t err__ = (expr);                                                              \
    if (err__ != cudaSuccess) {                                                              \
      ghn3::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
      return GHN3_ERR_CUDA;                                                                  \
    }                                                                                        \
  } while (0)

#define GHN3_LAUNCH_CHECK(name)                                                              \
  do {                                                                                       \
    cudaError_t err__ = cudaGetLastError();                                                  \
    if (err__ != cudaSuccess) {                                                              \
      ghn3::set_error("launch of %s failed: %s", name, cudaGetErrorString(err__));           \
      return GHN3_ERR_CUDA;                                                                  \
    }                                                                                        \
    ghn3::count_launch();                                                                    \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// fp32 -> tf32 with round-to-nearest (ties away), result kept in an fp32 container.
__device__ __forceinline__ float round_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// Typed loads/stores used by kernels templated on the activation storage type.
__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// mma.sync m16n8k16 bf16 x bf16 -> fp32 (flash-attention register layout) and bf16 pair packing
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *(uint32_t*)&v;
}

int num_sms();

// Programmatic dependent launch (PDL): a kernel launched with launch_pdl may start while its predecessor in the
// stream is still running. Everything before pdl_wait() must only touch data that no earlier kernel of the step
// writes (weights, look-up tables, descriptors, shared memory, TMEM); pdl_wait() returns once the predecessor grid
// has completed and its writes are visible.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

int persistent_cta_cap();   // api.cu: ghn3_set_persistent_ctas
bool pdl_enabled();     // api.cu: GHN3_NO_PDL=1 launches every kernel fully serialised (experiments)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace ghn3
