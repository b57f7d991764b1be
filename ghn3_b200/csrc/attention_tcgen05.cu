// Attention with the SPD / edge bias on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), bf16.
// Replaces reference ghn3/graphormer.py:121-140 for LARGE graphs (config 4: thousands of nodes), where the mma.sync
// kernel of dense_kernels.cu is limited by its operand-fragment loads from shared memory.
//
// CTA = (graph, head, 128 queries), 8 warps; key tiles of 128:
//   thread 0          : MMA issuer.  S = Q . K^T  : UMMA M = 128 queries, N = 128 keys, K = head dim (padded to 16/32)
//                                    PV = P . V   : UMMA M = 128 queries, N = 32 (head dim padded), K = 128 keys
//                       S lives in TMEM columns [0, 128), the tile's P.V in columns [128, 160)  (fp32)
//   warps 0..7        : two threads per query row (warps w and w + 4 may read the same TMEM lanes): keys 0..63 / 64..127
//                       of the tile, output dims 0..15 / 16..31. Per tile a thread
//                         - stages one K row / one V row of the tile into shared memory in the UMMA operand layouts
//                           (K-major rows of 128 B, 16-byte chunks XOR-swizzled by row & 7 = SWIZZLE_128B; V transposed)
//                         - pass 1: tcgen05.ld S, s = S * d^-1/2 * log2e + lut[pair] (bias from the per-head LUT by the
//                           (A_ij, A_ji) pair index), row maximum; s is written back to TMEM (tcgen05.st)
//                         - pass 2: p = exp2(s - m), row sum, P (bf16) -> shared memory as the A operand of P . V
//                         - o = o * exp2(m_old - m_new) + PV   (online softmax; o in registers, 32 floats per row)
// No TMA: the operands are 48-byte slices of [N][3C] rows that have to be padded / transposed on the way in.
// Two CTAs fit on an SM (87 KB of shared memory, 256 TMEM columns each): one CTA's softmax overlaps the other's MMAs.
#include <cstdlib>

#include "common.cuh"
#include "tcgen05.cuh"

namespace ghn3 {

constexpr int kTcQ = 128;
constexpr int kTcK = 128;
constexpr int kTcThreads = 256;              // 8 warps: two threads per query row; thread 0 also issues the MMAs
constexpr int kTcDV = 32;                 // head dim padded to the UMMA N granularity
constexpr int kTcTmemCols = 256;
constexpr uint32_t kTcSCol = 0, kTcPvCol = 128;

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int attn_tc_smem_bytes(int lut_size) {
  return 16384 /*Q*/ + 16384 /*K*/ + 2 * 16384 /*P*/ + 2 * 4096 /*V^T*/ + 2048 /*row exchange*/ +
         ((lut_size * 4 + 127) / 128) * 128 + 128 + 1024;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void bar_rows() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// 32 TMEM lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// Pass 1 over this thread's two 32-key chunks of its query row: logit = S * d^-1/2 * log2e + lut[pair] (written back
// to TMEM), returns the maximum. kMask: the tile reaches past the graph -- columns >= n get -inf.
template <bool kMask>
__device__ __forceinline__ float softmax_pass1(uint32_t t_s, const uint16_t* __restrict__ pidx, const float* __restrict__ sLut,
                                               float scale_log2, int k0, int n, int ld, int c_begin) {
  float tmax = -INFINITY;
#pragma unroll 1
  for (int c = c_begin; c < c_begin + 2; ++c) {
    const int col0 = k0 + c * 32;
    if (kMask && col0 >= n) break;                       // uniform over the warp pair
    uint32_t w[16];                                      // 32 pair indices of this row, two per word
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 t4 = make_uint4(0, 0, 0, 0);
      if (!kMask || col0 + i * 8 < ld) t4 = __ldg((const uint4*)(pidx + c * 32) + i);   // rows are padded to ld only
      w[4 * i] = t4.x; w[4 * i + 1] = t4.y; w[4 * i + 2] = t4.z; w[4 * i + 3] = t4.w;
    }
    uint32_t v[32];
    tmem_ld_32x32(t_s + (uint32_t)(c * 32), v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      uint32_t idx = (j & 1) ? (w[j >> 1] >> 16) : (w[j >> 1] & 0xFFFFu);
      if (kMask) idx = (col0 + j < n) ? idx : 0u;
      float sc = fmaf(__uint_as_float(v[j]), scale_log2, sLut[idx]);
      if (kMask) sc = (col0 + j < n) ? sc : -INFINITY;
      tmax = fmaxf(tmax, sc);
      v[j] = __float_as_uint(sc);
    }
    tmem_st_32x32(t_s + (uint32_t)(c * 32), v);
  }
  tmem_st_wait();
  return tmax;
}

// Pass 2: p = 2^(logit - m), P (bf16) into the swizzled A-operand row (this thread's 64 keys = one 16 KB K block);
// returns the sum over its columns. A key beyond the graph must contribute exactly 0 (its V row is zero, but
// 0 * garbage could be NaN), so fully masked chunks store zeros.
template <bool kMask>
__device__ __forceinline__ float softmax_pass2(uint32_t t_s, int k0, int n, int c_begin, float m_new, uint8_t* prow_smem,
                                               int sw) {
  float lt = 0.f;
#pragma unroll 1
  for (int c = c_begin; c < c_begin + 2; ++c) {
    const bool live = !kMask || (k0 + c * 32 < n);       // uniform
    uint32_t v[32];
    if (live) {
      tmem_ld_32x32(t_s + (uint32_t)(c * 32), v);
      tmem_ld_wait();
    }
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) {
      uint4 pk = make_uint4(0, 0, 0, 0);
      if (live) {
        float p[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          p[e] = ex2_approx(__uint_as_float(v[g8 * 8 + e]) - m_new);   // 2^(-inf) = 0 for masked columns
          lt += p[e];
        }
        pk.x = pack_bf16x2(p[0], p[1]); pk.y = pack_bf16x2(p[2], p[3]);
        pk.z = pack_bf16x2(p[4], p[5]); pk.w = pack_bf16x2(p[6], p[7]);
      }
      const int kk = (c & 1) * 32 + g8 * 8;              // key inside this thread's 64-key block
      *(uint4*)(prow_smem + (((kk >> 3) ^ sw) << 4)) = pk;
    }
  }
  return lt;
}

template <int D>
__global__ void __launch_bounds__(kTcThreads, 2) attention_tc_kernel(const ghn3_attention_args a) {
  static_assert(D % 8 == 0 && D <= 32, "head dim: 8, 16, 24, 32");
  constexpr int KK = (D + 15) / 16;          // K steps of Q . K^T
  constexpr int VPR = D / 8;                 // 16-byte vectors per q / k / v head row
  const int g = blockIdx.z, h = blockIdx.y;
  const int n0 = __ldg(a.node_off + g);
  const int n = __ldg(a.node_off + g + 1) - n0;
  const int q0 = blockIdx.x * kTcQ;
  if (q0 >= n) return;                       // before any barrier / TMEM allocation
  const int ld = (n + 15) & ~15;
  const int C = a.hid, C3 = 3 * C;

  extern __shared__ uint8_t attn_tc_smem[];
  const uint32_t base = (smem_u32(attn_tc_smem) + 1023u) & ~1023u;
  uint8_t* smem = attn_tc_smem + (base - smem_u32(attn_tc_smem));
  const uint32_t sQ = base, sK = base + 16384, sP = base + 32768, sVt = base + 65536;
  uint8_t* pQ = smem;
  uint8_t* pK = smem + 16384;
  uint8_t* pP = smem + 32768;
  uint8_t* pVt = smem + 65536;
  float* sX = (float*)(smem + 73728);                 // [2 tile parities][2 halves][128 rows] partial maxima / sums
  float* sLut = (float*)(smem + 73728 + 2048);
  const uint32_t lut_bytes = ((uint32_t)a.lut_size * 4 + 127) / 128 * 128;
  const uint32_t bar = base + 73728 + 2048 + lut_bytes;
  const uint32_t in_full = bar, s_full = bar + 8, p_full = bar + 16, pv_full = bar + 24, tmem_slot = bar + 32;
  uint32_t* tmem_slot_ptr = (uint32_t*)(smem + 73728 + 2048 + lut_bytes + 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(in_full, 256);
    mbar_init(s_full, 1);
    mbar_init(p_full, 256);
    mbar_init(pv_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc<kTcTmemCols>(tmem_slot);
  // zero the operand buffers once: padded dims of Q / K, padded rows of V^T are never written again
  for (int i = threadIdx.x; i < (16384 + 16384) / 16; i += kTcThreads) ((uint4*)pQ)[i] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < 8192 / 16; i += kTcThreads) ((uint4*)pVt)[i] = make_uint4(0, 0, 0, 0);
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < a.lut_size; i += kTcThreads)
    sLut[i] = __ldg(a.lut + (int64_t)h * a.lut_size + i) * 1.44269504088896340736f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // the zero fill is read by the tensor cores
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int n_tiles = (n + kTcK - 1) / kTcK;

  {
    // Two threads per query row: warps 0..3 (half 0) and 4..7 (half 1) own the same TMEM lane quarters (warp % 4);
    // half 0 takes keys 0..63 of a tile, the K rows of the staging and output dims 0..15, half 1 keys 64..127, the V
    // rows and dims 16..31. Row maxima / sums are exchanged through shared memory.
    const int r = (warp & 3) * 32 + lane;            // accumulator row = TMEM lane this thread may read
    const int half = warp >> 2;
    const bool issuer = threadIdx.x == 0;            // issues the tile's MMAs once everybody's operands have landed
    const uint32_t idesc_s = make_idesc(false, 128, kTcK);
    const uint32_t idesc_pv = make_idesc(false, 128, kTcDV);
    const uint32_t t_lane = (uint32_t)((warp & 3) * 32) << 16;
    const int qi = q0 + r;
    const bool q_ok = qi < n;
    const __nv_bfloat16* qkv = (const __nv_bfloat16*)a.qkv + (int64_t)n0 * C3;
    const uint16_t* prow = a.pair + a.mat_off[g] + (int64_t)(q_ok ? qi : q0) * ld;
    const float scale_log2 = rsqrtf((float)D) * 1.44269504088896340736f;
    const int sw = r & 7;
    pdl_wait();                                       // qkv comes from the preceding kernel

    if (half == 0) {                                  // Q row (zeros beyond the graph)
      const uint4* src = (const uint4*)(qkv + (int64_t)(q_ok ? qi : 0) * C3 + h * D);
#pragma unroll
      for (int c = 0; c < VPR; ++c)
        *(uint4*)(pQ + r * 128 + ((c ^ sw) << 4)) = q_ok ? __ldg(src + c) : make_uint4(0, 0, 0, 0);
    }
    float o[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) o[d] = 0.f;
    float m = -INFINITY, l = 0.f;                     // l: partial sum over this thread's columns

    // Software pipeline over key tiles: the K (half 0) / V (half 1) row of tile t + 1 is fetched into registers and
    // this row's pair indices of tile t + 1 are pulled into L1 while tile t is being processed.
    uint4 kvv[VPR];
    auto fetch_row = [&](int tile) {
      const int kj = tile * kTcK + r;
      const bool k_ok = kj < n;
      const uint4* src = (const uint4*)(qkv + (int64_t)(k_ok ? kj : 0) * C3 + (1 + half) * C + h * D);
#pragma unroll
      for (int c = 0; c < VPR; ++c) kvv[c] = k_ok ? __ldg(src + c) : make_uint4(0, 0, 0, 0);
    };
    auto prefetch_pairs = [&](int tile) {
      const int c0 = tile * kTcK + half * 64;
      if (c0 < ld) asm volatile("prefetch.global.L1 [%0];" ::"l"(prow + c0));
    };
    fetch_row(0);
    prefetch_pairs(0);

    for (int t = 0; t < n_tiles; ++t) {
      const uint32_t ph = (uint32_t)(t & 1);
      const int k0 = t * kTcK;
      // ---- stage key row k0 + r: K (K-major, swizzled) by half 0, V (transposed) by half 1 ----
      if (half == 0) {
#pragma unroll
        for (int c = 0; c < VPR; ++c) *(uint4*)(pK + r * 128 + ((c ^ sw) << 4)) = kvv[c];
      } else {
        const int kb = r >> 6, kc = (r & 63) >> 3, kbyte = (r & 7) * 2;
#pragma unroll
        for (int c = 0; c < VPR; ++c) {
          const __nv_bfloat16* ve = (const __nv_bfloat16*)&kvv[c];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int d = c * 8 + e;
            *(__nv_bfloat16*)(pVt + kb * 4096 + d * 128 + (((kc ^ (d & 7)) << 4) | kbyte)) = ve[e];
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic writes -> tensor-core reads
      tcgen05_fence_before();                                           // this thread's TMEM reads of the last tile
      mbar_arrive(in_full);
      if (t + 1 < n_tiles) {
        fetch_row(t + 1);
        prefetch_pairs(t + 1);
      }
      if (issuer) {                                   // S = Q . K^T
        mbar_wait(in_full, ph);
        tcgen05_fence_after();
        const uint64_t da = make_smem_desc(sQ), db = make_smem_desc(sK);
#pragma unroll
        for (int k = 0; k < KK; ++k) umma<false>(tmem_base + kTcSCol, da + 2 * k, db + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        tcgen05_commit(s_full);
      }
      __syncwarp();

      // ---- softmax of the tile ----
      mbar_wait(s_full, ph);
      tcgen05_fence_after();
      const bool tail = k0 + kTcK > n;
      const uint32_t t_s = tmem_base + t_lane + kTcSCol;
      float* xch = sX + (t & 1) * 256;
      const float tmax = tail ? softmax_pass1<true>(t_s, prow + k0, sLut, scale_log2, k0, n, ld, 2 * half)
                              : softmax_pass1<false>(t_s, prow + k0, sLut, scale_log2, k0, n, ld, 2 * half);
      xch[half * 128 + r] = tmax;
      bar_rows();
      const float m_new = fmaxf(m, fmaxf(tmax, xch[(half ^ 1) * 128 + r]));
      const float corr = ex2_approx(m - m_new);
      m = m_new;
      uint8_t* prs = pP + half * 16384 + r * 128;
      const float lt = tail ? softmax_pass2<true>(t_s, k0, n, 2 * half, m_new, prs, sw)
                            : softmax_pass2<false>(t_s, k0, n, 2 * half, m_new, prs, sw);
      l = l * corr + lt;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tcgen05_fence_before();
      mbar_arrive(p_full);
      if (issuer) {                                   // PV = P . V
        mbar_wait(p_full, ph);
        tcgen05_fence_after();
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t da = make_smem_desc(sP + kb * 16384), db = make_smem_desc(sVt + kb * 4096);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma<false>(tmem_base + kTcPvCol, da + 2 * k, db + 2 * k, idesc_pv, (kb | k) != 0 ? 1u : 0u);
        }
        tcgen05_commit(pv_full);
      }
      __syncwarp();

      // ---- o = o * corr + P . V  (this thread's 16 output dims) ----
      mbar_wait(pv_full, ph);
      tcgen05_fence_after();
      uint32_t pv[16];
      tmem_ld_32x16(tmem_base + t_lane + kTcPvCol + (uint32_t)(16 * half), pv);
      tmem_ld_wait();
#pragma unroll
      for (int d = 0; d < 16; ++d) o[d] = fmaf(o[d], corr, __uint_as_float(pv[d]));
    }
    // total row sum = the two halves' partial sums (same running maximum)
    float* xl = sX + ((n_tiles & 1) ? 256 : 0);        // the parity block the last tile did not use
    xl[half * 128 + r] = l;
    bar_rows();
    l += xl[(half ^ 1) * 128 + r];
    if (q_ok) {
      const float inv = 1.f / l;
      if (a.lse2 != nullptr && half == 0) a.lse2[(int64_t)h * a.total_nodes + n0 + qi] = m + log2f(l);
      __nv_bfloat16* out = (__nv_bfloat16*)a.out + (int64_t)(n0 + qi) * C + h * D + 16 * half;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (16 * half + c * 8 < D) {
          uint4 pk;
          pk.x = pack_bf16x2(o[c * 8 + 0] * inv, o[c * 8 + 1] * inv); pk.y = pack_bf16x2(o[c * 8 + 2] * inv, o[c * 8 + 3] * inv);
          pk.z = pack_bf16x2(o[c * 8 + 4] * inv, o[c * 8 + 5] * inv); pk.w = pack_bf16x2(o[c * 8 + 6] * inv, o[c * 8 + 7] * inv);
          *(uint4*)(out + c * 8) = pk;
        }
      }
    }
  }

  __syncwarp();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tcgen05_fence_after();
    tmem_dealloc<kTcTmemCols>(tmem_base);
  }
}

template <int D>
static int launch_attention_tc(const ghn3_attention_args* a, cudaStream_t stream) {
  const int smem = attn_tc_smem_bytes(a->lut_size);
  static bool configured = false;
  if (!configured) {
    GHN3_CUDA(cudaFuncSetAttribute(attention_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    configured = true;
  }
  if (smem > 110 * 1024) return GHN3_ERR_UNSUPPORTED;
  const dim3 grid((unsigned)ceil_div(a->max_nodes, kTcQ), (unsigned)a->heads, (unsigned)a->n_graphs);
  GHN3_CUDA(launch_pdl(attention_tc_kernel<D>, grid, dim3(kTcThreads), (size_t)smem, stream, *a));
  GHN3_LAUNCH_CHECK("attention_tc_kernel");
  return GHN3_OK;
}

// bf16 attention on tcgen05; GHN3_ERR_UNSUPPORTED if the head dim / alignment does not fit (the caller falls back to
// the mma.sync kernel).
int attention_tc_impl(const ghn3_attention_args* a, cudaStream_t stream) {
  const int D = a->hid / a->heads;
  if (a->dtype != GHN3_BF16 || (a->hid * 2) % 16 != 0 || (((uintptr_t)a->qkv) & 15) != 0 || (((uintptr_t)a->out) & 15) != 0)
    return GHN3_ERR_UNSUPPORTED;
  if (D == 8) return launch_attention_tc<8>(a, stream);
  if (D == 16) return launch_attention_tc<16>(a, stream);
  if (D == 24) return launch_attention_tc<24>(a, stream);
  if (D == 32) return launch_attention_tc<32>(a, stream);
  return GHN3_ERR_UNSUPPORTED;
}

}  // namespace ghn3
