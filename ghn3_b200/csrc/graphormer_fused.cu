// The Graphormer layer stack as ONE persistent kernel (sm_100a) -- see include/ghn3_b200.h (3c).
//
// Replaces, for small packed batches, the 7 launches per layer of ghn3_graphormer_stack: reference
// ghn3/graphormer.py:208-248 (pre-LN block), :119-142 (attention with edge bias), :22-47 (FFN), ghn3/nn.py:258-261.
//
// One CTA per SM, 10 warps:
//   warp 0 / one lane : weight producer -- TMA boxes of 128 weight rows x 64 K (SWIZZLE_128B) into an 8-slot ring.
//                       Weights never depend on activations, so the ring runs ahead of every dependency: the next
//                       tile's weights land while this tile's epilogue / the arrival counters are still in flight.
//   warp 1 / one lane : MMA issuer      -- tcgen05.mma kind::f16, A = weight block (UMMA M = 128 output features),
//                                          B = activation rows (UMMA N = rb), fp32 accumulator in TMEM.
//   warps 2..9        : compute         -- wait for the tile's inputs (arrival counters), build the activation operand
//                                          (LayerNorm prologue written straight into the swizzled UMMA layout, or a
//                                          TMA load), run the epilogue (TMEM -> registers -> global), signal; and the
//                                          attention units (mma.sync m16n8k16, flash-style, edge bias from a LUT).
// Per layer, five stages of tiles; a tile of stage s only waits for the tiles of stage s-1 that produce ITS rows
// (attention: the rows of its graph), through counters per 16-row granule that only ever grow:
//   0 QKV   qkv[l&1] = LN1(x) . Wqkv^T                         waits x     (FFN2 of the previous layer)
//   1 ATTN  ao = softmax(q k^T d^-1/2 + bias) v                waits qkv   (all rows of the graph)
//   2 PROJ  x += ao . Wout^T + b                               waits ao
//   3 FFN1  ff = gelu(LN2(x) . W1^T + b1)                      waits x (proj)
//   4 FFN2  x += ff . W2^T + b2                                waits ff
// Every CTA walks the stages in the same order and all CTAs are resident, so a waiting tile's producers are always
// ahead of it in some CTA's program order: no deadlock. qkv is double-buffered by layer parity because attention
// units of OTHER query blocks still read a graph's k/v rows when the next layer's QKV tile for these rows is ready.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tcgen05.cuh"

namespace ghn3 {

constexpr int kFRing = 8;                   // weight ring slots
constexpr int kFWBytes = 128 * 128;         // one slot: 128 weight rows x 64 bf16
constexpr int kFHalf = 4;                   // ring slots per barrier: the ring is handed over in halves
constexpr int kFActBytes = 64 * 1024;       // activation operand of ONE tile: all its K blocks, rb rows x 128 B each
constexpr int kFAcc1 = 64;                  // TMEM column of the second accumulator (odd K steps)
constexpr int kFMaxRb = 64;
constexpr int kFThreads = 320;
constexpr int kFCompute = 256;
constexpr int kFGran = 16;                  // rows per arrival counter
constexpr int kFQB = 64;                    // queries per attention unit
constexpr int kFKT = 448;                   // keys staged per attention tile (multiple of 64; K + V^T fit the activation buffer)
constexpr int kFMaxGraphs = 256;
constexpr int kFMaxGran = 512;              // total_nodes <= 8192
constexpr int kFMergeBytes = 128 * 16 * 4;  // (m0, m1, l0, l1, o[3][4]) of the second key half
constexpr int kFLutBytes = 10496;           // 51 * 51 floats, rounded up
constexpr int kFCntStride = 32;             // int32 per arrival counter: every counter has its own 128-byte line

enum { ST_QKV = 0, ST_ATTN = 1, ST_PROJ = 2, ST_FF1 = 3, ST_FF2 = 4 };
enum { CNT_X = 0, CNT_QKV = 1, CNT_AO = 2, CNT_X2 = 3, CNT_FF = 4, CNT_GRAPH = 5 };   // CNT_GRAPH: per graph, not granule

struct FusedStage {
  int32_t rb, n_ft, splits, kbps, nkb, n_tiles, n_feat, pad;
};

struct FusedKernelArgs {
  int32_t C, H, L, M, n_graphs, lut_size, n_gran, total_stages;
  const int32_t* node_off;
  const int64_t* mat_off;
  const uint16_t* pair;
  const float* lut;
  const ghn3_layer_weights* layers;
  float* x;
  __nv_bfloat16* ao;
  __nv_bfloat16* qkv;
  __nv_bfloat16* ff;
  int32_t* cnt;
  long long* trace;
  FusedStage st[5];
};

constexpr int fused_smem_bytes() {
  return kFRing * kFWBytes + kFActBytes + kFMergeBytes + kFLutBytes + (kFMaxGraphs + 8) * 4 + kFMaxGran * 4 + kFMaxGraphs * 4 + 512 + 1024;
}

// Arrival counters are polled with RELAXED gpu-scope loads: ld.acquire.gpu is LDG.STRONG + CCTL.IVALL (a full L1
// invalidation per poll). No L1 invalidation is needed here -- everything a consumer reads after the poll was written
// by other SMs and is read through L2 (ld.global.cg / TMA); the control dependency on the polled value plus bar.sync
// order those reads after it, and the producer's red.release.gpu (MEMBAR.GPU) ordered its data before the counter.
__device__ __forceinline__ int ld_relaxed_gpu(const int32_t* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// One MEMBAR.GPU, then any number of relaxed arrivals (red.release would repeat the fence for every counter).
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void red_relaxed_gpu(int32_t* p, int v) {
  asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Bounded spin on an arrival counter: a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void wait_counter(const int32_t* p, int target) {
  if (ld_relaxed_gpu(p) >= target) return;
  const long long t0 = clock64();
  while (ld_relaxed_gpu(p) < target) {
    __nanosleep(40);                          // one poller per counter and CTA; keep the counter's L2 slice free
    if (clock64() - t0 > 4000000000LL) {
      printf("ghn3 fused graphormer: counter timeout (block %d, thread %d, have %d, want %d)\n", (int)blockIdx.x,
             (int)threadIdx.x, ld_relaxed_gpu(p), target);
      __trap();
    }
  }
}

// Bring-up / profiling aid: thread 64 of every CTA appends (tag, clock64, globaltimer) records.
constexpr int kFTraceMax = 1024;
struct Tracer {
  long long* buf;
  int n;
  __device__ __forceinline__ void operator()(int tag) {
    if (buf != nullptr && n < kFTraceMax) {
      long long g;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g));
      buf[3 * n] = tag;
      buf[3 * n + 1] = clock64();
      buf[3 * n + 2] = g;
      ++n;
    }
  }
};

struct TileCoord {
  int rblk, ft, sp, kb0, kb1, r0, r1;
};
__device__ __forceinline__ TileCoord decode_tile(const FusedStage& S, int t, int M) {
  TileCoord c;
  const int per = S.n_ft * S.splits;
  c.rblk = t / per;
  const int rem = t - c.rblk * per;
  c.ft = rem / S.splits;
  c.sp = rem - c.ft * S.splits;
  c.kb0 = c.sp * S.kbps;
  c.kb1 = min(S.nkb, c.kb0 + S.kbps);
  c.r0 = c.rblk * S.rb;
  c.r1 = min(M, c.r0 + S.rb);
  return c;
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm prologue: rows [r0, r1) of x (fp32, read from L2 -- other SMs wrote them) -> bf16 B operand in shared
// memory, K-major rows of 128 bytes per 64-wide K block (block kb at act + kb * rb * 128), 16-byte chunks
// XOR-swizzled by (row & 7) exactly as a SWIZZLE_128B TMA box would have laid them out. gamma / beta come in
// registers (static data, fetched by the caller before it waits for the tile's inputs). One warp per row, two-pass mean / variance in registers (same
// arithmetic as layernorm_kernel). Rows >= r1 are left untouched: an accumulator column only depends on its own row.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ln_prologue(const float* __restrict__ x, int C, int r0, int r1, int rb,
                                            const float4 (&g)[3], const float4 (&b)[3], uint8_t* act, int cw, int lane) {
  const int C4 = C >> 2;
  const float invC = 1.f / (float)C;
  const int rpw = rb >> 3;                    // rows per warp: tile rows cw, cw + 8, ...
#pragma unroll 1
  for (int base = 0; base < rpw; base += 4) {
    // four rows per pass, no branches between them: their loads, shuffles and arithmetic interleave (the warp has
    // nobody to hide latency behind -- two warps per scheduler)
    float4 v[4][3];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int row = r0 + cw + 8 * (base + u);
      ok[u] = (base + u < rpw) && row < r1;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int f = lane + 32 * i;
        v[u][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok[u] && f < C4) v[u][i] = __ldcg((const float4*)(x + (int64_t)row * C) + f);
      }
    }
    float sum[4], sq[4], mean[4], rstd[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      sum[u] = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i) sum[u] += (v[u][i].x + v[u][i].y) + (v[u][i].z + v[u][i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < 4; ++u) sum[u] += __shfl_xor_sync(0xffffffffu, sum[u], o);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      mean[u] = sum[u] * invC;
      sq[u] = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (lane + 32 * i < C4) {
          const float dx = v[u][i].x - mean[u], dy = v[u][i].y - mean[u], dz = v[u][i].z - mean[u], dw = v[u][i].w - mean[u];
          sq[u] += (dx * dx + dy * dy) + (dz * dz + dw * dw);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < 4; ++u) sq[u] += __shfl_xor_sync(0xffffffffu, sq[u], o);
#pragma unroll
    for (int u = 0; u < 4; ++u) rstd[u] = 1.0f / sqrtf(sq[u] * invC + 1e-5f);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int rr = cw + 8 * (base + u);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int f = lane + 32 * i;
        if (ok[u] && f < C4) {
          uint2 pk;
          pk.x = pack_bf16x2((v[u][i].x - mean[u]) * rstd[u] * g[i].x + b[i].x, (v[u][i].y - mean[u]) * rstd[u] * g[i].y + b[i].y);
          pk.y = pack_bf16x2((v[u][i].z - mean[u]) * rstd[u] * g[i].z + b[i].z, (v[u][i].w - mean[u]) * rstd[u] * g[i].w + b[i].w);
          const int k = f * 4;
          const int kb = k >> 6, chunk = (k & 63) >> 3, byte = (k & 7) * 2;
          *(uint2*)(act + kb * (rb * 128) + rr * 128 + (((chunk ^ (rr & 7)) << 4) | byte)) = pk;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// One attention unit: graph g, head h, queries [q0, q0 + 64) -- adapted from attention_mma_kernel (dense_kernels.cu).
// Warp cw: query group qg = cw & 3 (16 queries), key half kh = cw >> 2 (64-key chunks of parity kh); the two halves
// are merged through shared memory at the end. q / k / v were written by other SMs during this kernel: ld.global.cg.
// ---------------------------------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ void attention_unit(const FusedKernelArgs& a, const __nv_bfloat16* __restrict__ qkvbuf, int n0,
                                               int n, int64_t moff, int q0, int h, uint8_t* smem_kv, float* s_merge,
                                               const float* sLut, int ct, Tracer& tr, int tagbase) {
  constexpr int DK = (D + 15) / 16 * 16;
  constexpr int DS = DK + 8;
  constexpr int DN = (D + 7) / 8 * 8;
  constexpr int VS = kFKT + 8;
  constexpr int NT2 = DN / 8;
  constexpr int KK = DK / 16;
  constexpr int ROW_BYTES = D * 2;
  static_assert(ROW_BYTES % 16 == 0, "head dim must be a multiple of 8");
  constexpr int VPR = ROW_BYTES / 16;
  static_assert((kFKT * DS + DN * VS) * 2 <= kFActBytes, "K / V tile must fit in the activation buffer");
  static_assert(4 + NT2 * 4 <= 16, "merge record");
  __nv_bfloat16* sK = (__nv_bfloat16*)smem_kv;          // [KT][DS]
  __nv_bfloat16* sVt = sK + kFKT * DS;                  // [DN][VS]

  const int C = a.C, C3 = 3 * C;
  const int ld = (n + 15) & ~15;
  const int cw = ct >> 5, lane = ct & 31;
  const int qg = cw & 3, kh = cw >> 2;
  const int gq = lane >> 2, tq = lane & 3;
  const __nv_bfloat16* qkv = qkvbuf + (int64_t)n0 * C3;
  const uint16_t* pair = a.pair + moff;
  const float scale_log2 = rsqrtf((float)D) * 1.44269504088896340736f;

  const int r0 = q0 + qg * 16 + gq, r1 = r0 + 8;
  const bool ok0 = r0 < n, ok1 = r1 < n;
  const bool warp_active = (q0 + qg * 16) < n;          // warp-uniform
  // raw Q words: requested now, converted after the K / V staging (their L2 latency hides behind it)
  uint32_t qraw[KK][2][2];
#pragma unroll
  for (int kk = 0; kk < KK; ++kk) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int d = kk * 16 + half * 8 + 2 * tq;
      qraw[kk][half][0] = 0; qraw[kk][half][1] = 0;
      if (d < D) {
        if (ok0) qraw[kk][half][0] = __ldcg((const uint32_t*)(qkv + (int64_t)r0 * C3 + h * D + d));
        if (ok1) qraw[kk][half][1] = __ldcg((const uint32_t*)(qkv + (int64_t)r1 * C3 + h * D + d));
      }
    }
  }
  const uint16_t* prow0 = pair + (int64_t)(ok0 ? r0 : q0) * ld;
  const uint16_t* prow1 = pair + (int64_t)(ok1 ? r1 : q0) * ld;
  // edge-bias indices of a 16 x 64 block (two keys per 32-bit load), fetched one block ahead of their use
  uint32_t pw0[8], pw1[8];
  auto load_pairs = [&](int colbase) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int col = colbase + nt * 8 + 2 * tq;
      pw0[nt] = 0; pw1[nt] = 0;
      if (warp_active && col < n) {
        pw0[nt] = __ldg((const uint32_t*)(prow0 + col));
        pw1[nt] = __ldg((const uint32_t*)(prow1 + col));
      }
    }
  };
  load_pairs(kh * 64);

  uint32_t aq[KK][4];
  float o[NT2][4];
#pragma unroll
  for (int i = 0; i < NT2; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int k0 = 0; k0 < n; k0 += kFKT) {
    const int kt = min(kFKT, n - k0);
    const int kt64 = (kt + 63) & ~63;
    bar_compute();                          // the staging buffer is free (previous tile / previous user consumed)
    {
      // all loads of a pass (4 x 256 vectors of K and of V) are in flight together: one L2 round trip per pass
      const char* kbase = (const char*)(qkv + (int64_t)k0 * C3 + C + h * D);
      const size_t row_stride = (size_t)C3 * 2, v_off = (size_t)C * 2;
      const int nvec = kt64 * VPR;
#pragma unroll 1
      for (int base = 0; base < nvec; base += 4 * kFCompute) {
        uint4 kv[4], vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = base + u * kFCompute + ct;
          const int j = idx / VPR, c = idx - j * VPR;
          kv[u] = make_uint4(0, 0, 0, 0);
          vv[u] = make_uint4(0, 0, 0, 0);
          if (idx < nvec && j < kt) {
            const char* src = kbase + (size_t)j * row_stride + c * 16;
            kv[u] = __ldcg((const uint4*)src);
            vv[u] = __ldcg((const uint4*)(src + v_off));
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = base + u * kFCompute + ct;
          const int j = idx / VPR, c = idx - j * VPR;
          if (idx < nvec) {
            *(uint4*)(sK + j * DS + c * 8) = kv[u];
            if (DK > D && c == 0) *(uint4*)(sK + j * DS + D) = make_uint4(0, 0, 0, 0);   // zero the padded dims
            const __nv_bfloat16* ve = (const __nv_bfloat16*)&vv[u];
#pragma unroll
            for (int e = 0; e < 8; ++e) sVt[(c * 8 + e) * VS + j] = ve[e];
          }
        }
      }
    }
    bar_compute();
    if (ct == 0) tr(tagbase | 3);
    if (!warp_active) continue;
    if (k0 == 0) {
      // Q fragments of this warp's 16 queries, pre-scaled by d^-1/2 * log2(e)
#pragma unroll
      for (int kk = 0; kk < KK; ++kk) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const __nv_bfloat162 t0 = *(const __nv_bfloat162*)&qraw[kk][half][0];
          const __nv_bfloat162 t1 = *(const __nv_bfloat162*)&qraw[kk][half][1];
          aq[kk][half * 2 + 0] = pack_bf16(__low2float(t0) * scale_log2, __high2float(t0) * scale_log2);
          aq[kk][half * 2 + 1] = pack_bf16(__low2float(t1) * scale_log2, __high2float(t1) * scale_log2);
        }
      }
    }
    for (int c0 = kh * 64; c0 < kt; c0 += 128) {
      uint32_t cw0[8], cw1[8];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) { cw0[nt] = pw0[nt]; cw1[nt] = pw1[nt]; }
      // this warp's next block: 128 keys further in this tile, or its first block of the next tile
      load_pairs(c0 + 128 < kt ? k0 + c0 + 128 : k0 + kFKT + kh * 64);
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
  // merge the two key halves: warps 4..7 publish their running (m, l, o), warps 0..3 combine and store
  float* rec = s_merge + ((qg * 32 + lane) * 16);
  if (kh == 1) {
    rec[0] = m0; rec[1] = m1; rec[2] = l0; rec[3] = l1;
#pragma unroll
    for (int i = 0; i < NT2; ++i) { rec[4 + 4 * i] = o[i][0]; rec[5 + 4 * i] = o[i][1]; rec[6 + 4 * i] = o[i][2]; rec[7 + 4 * i] = o[i][3]; }
  }
  bar_compute();
  if (ct == 0) tr(tagbase | 4);
  if (kh == 0 && warp_active) {
    const float mb0 = rec[0], mb1 = rec[1];
    const float mn0 = fmaxf(m0, mb0), mn1 = fmaxf(m1, mb1);
    const float ca0 = exp2f(m0 - mn0), ca1 = exp2f(m1 - mn1);      // this half always saw >= 1 key: m0, m1 finite
    const float cb0 = exp2f(mb0 - mn0), cb1 = exp2f(mb1 - mn1);    // exp2(-inf) = 0 when the other half saw none
    l0 = l0 * ca0 + rec[2] * cb0;
    l1 = l1 * ca1 + rec[3] * cb1;
#pragma unroll
    for (int i = 0; i < NT2; ++i) {
      o[i][0] = o[i][0] * ca0 + rec[4 + 4 * i] * cb0; o[i][1] = o[i][1] * ca0 + rec[5 + 4 * i] * cb0;
      o[i][2] = o[i][2] * ca1 + rec[6 + 4 * i] * cb1; o[i][3] = o[i][3] * ca1 + rec[7 + 4 * i] * cb1;
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
#pragma unroll
    for (int i = 0; i < NT2; ++i) {
      const int d = i * 8 + 2 * tq;
      if (d < D) {
        if (ok0) *(uint32_t*)(a.ao + (int64_t)(n0 + r0) * C + h * D + d) = pack_bf16(o[i][0] * i0, o[i][1] * i0);
        if (ok1) *(uint32_t*)(a.ao + (int64_t)(n0 + r1) * C + h * D + d) = pack_bf16(o[i][2] * i1, o[i][3] * i1);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kFThreads, 1)
graphormer_fused_kernel(const __grid_constant__ CUtensorMap m_wqkv, const __grid_constant__ CUtensorMap m_wout,
                        const __grid_constant__ CUtensorMap m_wff1, const __grid_constant__ CUtensorMap m_wff2,
                        const __grid_constant__ CUtensorMap m_ao, const __grid_constant__ CUtensorMap m_ff,
                        const FusedKernelArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sRing = smem_base;
  const uint32_t sAct = sRing + kFRing * kFWBytes;
  uint8_t* act = smem + kFRing * kFWBytes;
  float* s_merge = (float*)(act + kFActBytes);
  float* sLut = (float*)((uint8_t*)s_merge + kFMergeBytes);
  int32_t* s_noff = (int32_t*)((uint8_t*)sLut + kFLutBytes);
  int32_t* s_expect = s_noff + kFMaxGraphs + 8;
  int32_t* s_nunits = s_noff + kFMaxGraphs + 4;         // attention units = H x (64-query blocks over all graphs)
  int32_t* s_gexpect = s_expect + kFMaxGran;            // QKV tiles overlapping each graph (arrivals per layer)
  const uint32_t bar_base = sAct + kFActBytes + kFMergeBytes + kFLutBytes + (kFMaxGraphs + 8) * 4 + kFMaxGran * 4 +
                            kFMaxGraphs * 4;
  const uint32_t ring_full = bar_base;                    // one per ring half (kFHalf slots)
  const uint32_t ring_empty = bar_base + 16;
  const uint32_t act_full = bar_base + 32;                // one: the activation operand of the current tile
  const uint32_t tmem_full = bar_base + 40;
  const uint32_t tmem_slot = tmem_full + 8;
  uint32_t* tmem_slot_ptr = (uint32_t*)(smem + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = a.M;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&m_wqkv); tma_prefetch_desc(&m_wout); tma_prefetch_desc(&m_wff1);
    tma_prefetch_desc(&m_wff2); tma_prefetch_desc(&m_ao); tma_prefetch_desc(&m_ff);
    for (int s = 0; s < kFRing / kFHalf; ++s) {
      mbar_init(ring_full + 8 * s, 1);
      mbar_init(ring_empty + 8 * s, 1);
    }
    mbar_init(act_full, 1);
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<2 * kFAcc1>(tmem_slot);
  // graph offsets and the expected attention arrivals per granule (static inputs: legal before pdl_wait)
  for (int i = threadIdx.x; i <= a.n_graphs; i += kFThreads) s_noff[i] = __ldg(a.node_off + i);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  for (int gr = threadIdx.x; gr < a.n_gran; gr += kFThreads) {
    // units (graph g, 64-query block) overlapping rows [16 gr, 16 gr + 16), times the number of heads
    const int lo = gr * kFGran, hi = min(M, lo + kFGran);
    int cntu = 0;
    for (int g = 0; g < a.n_graphs; ++g) {
      const int g0 = s_noff[g], g1 = s_noff[g + 1];
      const int b0 = max(lo, g0), b1 = min(hi, g1);
      if (b0 < b1) cntu += (b1 - 1 - g0) / kFQB - (b0 - g0) / kFQB + 1;
    }
    s_expect[gr] = cntu * a.H;
  }
  for (int g = threadIdx.x; g < a.n_graphs; g += kFThreads) {
    const int rbq = a.st[ST_QKV].rb;
    const int g0 = s_noff[g], g1 = s_noff[g + 1];
    s_gexpect[g] = g1 > g0 ? ((g1 - 1) / rbq - g0 / rbq + 1) * a.st[ST_QKV].n_ft : 0;
  }
  if (threadIdx.x == 0) {
    int nqb = 0;
    for (int g = 0; g < a.n_graphs; ++g) nqb += (s_noff[g + 1] - s_noff[g] + kFQB - 1) / kFQB;
    *s_nunits = nqb * a.H;
  }
  __syncthreads();
  pdl_launch_dependents();

  if (warp == 0) {
    // ===== weight producer =====
    // The ring is handed over in halves of kFHalf slots: one `full` barrier (expect_tx = the blocks that will land in
    // the half) and one `empty` barrier (tcgen05.commit after its last block) per half -- a wait / commit per K block
    // costs more than the block's four MMAs at these tile sizes. Halves ignore tile boundaries; weights of later tiles
    // never depend on anything, so a half that straddles two tiles still always fills.
    if (lane == 0) {
      uint32_t total = 0;                    // K blocks this CTA will consume over the whole kernel
      for (int gs = 0; gs < a.total_stages; ++gs) {
        const int s = gs % 5;
        if (s == ST_ATTN) continue;
        const FusedStage S = a.st[s];
        for (int t = blockIdx.x; t < S.n_tiles; t += gridDim.x) total += (uint32_t)S.nkb;
      }
      uint32_t it = 0;
      for (int gs = 0; gs < a.total_stages; ++gs) {
        const int l = gs / 5, s = gs - l * 5;
        if (s == ST_ATTN) continue;
        const FusedStage S = a.st[s];
        const CUtensorMap* wm = s == ST_QKV ? &m_wqkv : s == ST_PROJ ? &m_wout : s == ST_FF1 ? &m_wff1 : &m_wff2;
        for (int t = blockIdx.x; t < S.n_tiles; t += gridDim.x) {
          const TileCoord tc = decode_tile(S, t, M);
          const int wrow = l * S.n_feat + tc.ft * 128;
          for (int kb = tc.kb0; kb < tc.kb1; ++kb, ++it) {
            const int slot = it % kFRing, hf = slot / kFHalf;
            if (it % kFHalf == 0) {
              mbar_wait(ring_empty + 8 * hf, ((it / kFRing) & 1) ^ 1);
              mbar_arrive_expect_tx(ring_full + 8 * hf, min((uint32_t)kFHalf, total - it) * kFWBytes);
            }
            tma_load_2d(sRing + slot * kFWBytes, wm, ring_full + 8 * hf, kb * 64, wrow);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // Two accumulators (TMEM columns [0, rb) and [64, 64 + rb)) take the even / odd K steps: consecutive MMAs into ONE
    // accumulator are a dependent chain whose turnaround exceeds the 16..32 cycles of work an N = rb instruction holds.
    if (lane == 0) {
      uint32_t it = 0, tcount = 0;
      Tracer tr = {nullptr, 0};
      if (a.trace != nullptr) tr.buf = a.trace + ((int64_t)blockIdx.x * kFTraceMax + kFTraceMax / 2) * 3;
      for (int gs = 0; gs < a.total_stages; ++gs) {
        const int s = gs % 5;
        if (s == ST_ATTN) continue;
        const FusedStage S = a.st[s];
        const uint32_t idesc = make_idesc(false, 128, S.rb);
        for (int t = blockIdx.x; t < S.n_tiles; t += gridDim.x, ++tcount) {
          const TileCoord tc = decode_tile(S, t, M);
          mbar_wait(act_full, tcount & 1);
          tr((gs << 8) | 16);
          const int nk = tc.kb1 - tc.kb0;
          for (int i = 0; i < nk; ++i, ++it) {
            const int slot = it % kFRing, hf = slot / kFHalf;
            if (it % kFHalf == 0 || i == 0) mbar_wait(ring_full + 8 * hf, (it / kFRing) & 1);
            if (i == 0) tr((gs << 8) | 17);
            tcgen05_fence_after();
            const uint64_t da = make_smem_desc(sRing + slot * kFWBytes);
            const uint64_t db = make_smem_desc(sAct + i * (S.rb * 128));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma<false>(tmem_base + (uint32_t)((k & 1) * kFAcc1), da + 2 * k, db + 2 * k, idesc, (i == 0 && k < 2) ? 0u : 1u);
            if (it % kFHalf == kFHalf - 1) tcgen05_commit(ring_empty + 8 * hf);
          }
          tcgen05_commit(tmem_full);
          tr((gs << 8) | 18);
        }
      }
    }
  } else {
    // ===== compute warps =====
    const int ct = threadIdx.x - 64;
    const int cw = ct >> 5;
    const int wq = warp & 3;                 // TMEM lane quarter this warp may read
    const int half = cw >> 2;                // which half of the accumulator columns this warp drains
    const int C = a.C;
    int cur_head = -1;
    uint32_t tcount = 0;
    Tracer tr = {nullptr, 0};
    if (ct == 0 && a.trace != nullptr) tr.buf = a.trace + (int64_t)blockIdx.x * (3 * kFTraceMax);
    if (ct == 0) tr(-1);
    pdl_wait();                              // x comes from the preceding kernel (node features)
    if (ct == 0) tr(-2);
    for (int gs = 0; gs < a.total_stages; ++gs) {
      const int l = gs / 5, s = gs - l * 5;
      const FusedStage S = a.st[s];
      const ghn3_layer_weights* lw = a.layers + l;
      __nv_bfloat16* qkvbuf = a.qkv + (int64_t)(l & 1) * M * 3 * C;
      if (s == ST_ATTN) {
        // units: u = (query block index over all graphs) * H + head
        const int n_units = *s_nunits;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
          const int h = u % a.H;
          int qb = u / a.H, g = 0;
          for (; g < a.n_graphs; ++g) {
            const int nq = (s_noff[g + 1] - s_noff[g] + kFQB - 1) / kFQB;
            if (qb < nq) break;
            qb -= nq;
          }
          const int n0 = s_noff[g], n = s_noff[g + 1] - n0, q0 = qb * kFQB;
          // inputs: q, k, v of every row of the graph
          const int tagbase = gs << 8;
          if (ct == 0) tr(tagbase | 0);
          if (ct == 0) wait_counter(a.cnt + (CNT_GRAPH * a.n_gran + g) * kFCntStride, s_gexpect[g] * (l + 1));
          bar_compute();                     // every granule of the graph has arrived; nobody reads the old LUT
          if (ct == 0) tr(tagbase | 1);
          if (h != cur_head) {
            for (int i = ct; i < a.lut_size; i += kFCompute)
              sLut[i] = __ldg(a.lut + (int64_t)h * a.lut_size + i) * 1.44269504088896340736f;
            cur_head = h;
          }
          attention_unit<D>(a, qkvbuf, n0, n, __ldg(a.mat_off + g), q0, h, act, s_merge, sLut, ct, tr, tagbase);
          bar_compute();
          if (ct == 0) tr(tagbase | 5);
          if (ct == 0) {
            fence_gpu();                     // after bar.sync: the CTA's stores are ordered before the arrivals
            const int a0 = (n0 + q0) / kFGran, a1 = (n0 + min(q0 + kFQB, n) + kFGran - 1) / kFGran;
            for (int gr = a0; gr < a1; ++gr) red_relaxed_gpu(a.cnt + (CNT_AO * a.n_gran + gr) * kFCntStride, 1);
            tr(tagbase | 6);
          }
        }
        continue;
      }
      for (int t = blockIdx.x; t < S.n_tiles; t += gridDim.x, ++tcount) {
        const TileCoord tc = decode_tile(S, t, M);
        const int nk = tc.kb1 - tc.kb0;
        const int g0 = tc.r0 / kFGran, g1 = (tc.r1 + kFGran - 1) / kFGran;
        // static per-tile inputs, requested before waiting for the producers: LayerNorm gamma / beta, the bias
        float4 lg[3], lb[3];
        if (s == ST_QKV || s == ST_FF1) {
          const float* gp = s == ST_QKV ? lw->ln1_w : lw->ln2_w;
          const float* bp = s == ST_QKV ? lw->ln1_b : lw->ln2_b;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            lg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            lb[i] = lg[i];
            if (lane + 32 * i < (C >> 2)) {
              lg[i] = __ldg((const float4*)gp + lane + 32 * i);
              lb[i] = __ldg((const float4*)bp + lane + 32 * i);
            }
          }
        }
        const int f = tc.ft * 128 + wq * 32 + lane;
        const bool f_ok = f < S.n_feat;
        float bias = 0.f;
        if (f_ok) {
          if (s == ST_PROJ) bias = __ldg(lw->b_out + f);
          else if (s == ST_FF1) bias = __ldg(lw->b_ff1 + f);
          else if (s == ST_FF2) bias = __ldg(lw->b_ff2 + f);
        }
        const int tagbase = gs << 8;
        if (ct == 0) tr(tagbase | 0);
        // ---- inputs ----
        if (ct < g1 - g0) {
          const int gr = g0 + ct;
          // arrivals per granule and layer = (feature tiles x K splits) of the PRODUCING stage
          if (s == ST_QKV) { if (l > 0) wait_counter(a.cnt + (CNT_X * a.n_gran + gr) * kFCntStride, a.st[ST_FF2].pad * l); }
          else if (s == ST_PROJ) wait_counter(a.cnt + (CNT_AO * a.n_gran + gr) * kFCntStride, s_expect[gr] * (l + 1));
          else if (s == ST_FF1) wait_counter(a.cnt + (CNT_X2 * a.n_gran + gr) * kFCntStride, a.st[ST_PROJ].pad * (l + 1));
          else wait_counter(a.cnt + (CNT_FF * a.n_gran + gr) * kFCntStride, a.st[ST_FF1].pad * (l + 1));
        }
        bar_compute();
        if (ct == 0) tr(tagbase | 1);
        // ---- activation operand: K block i of the tile at act + i * rb * 128 (free: the previous tile's MMAs are done) ----
        if (s == ST_QKV || s == ST_FF1) {
          ln_prologue(a.x, C, tc.r0, tc.r1, S.rb, lg, lb, act, cw, lane);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic writes -> tensor-core reads
          bar_compute();
          if (ct == 0) mbar_arrive(act_full);
        } else if (ct == 0) {
          asm volatile("fence.proxy.async.global;" ::: "memory");          // other SMs' stores -> our TMA reads
          const CUtensorMap* am = s == ST_PROJ ? &m_ao : &m_ff;
          mbar_arrive_expect_tx(act_full, (uint32_t)(nk * S.rb * 128));
          for (int i = 0; i < nk; ++i) tma_load_2d(sAct + i * (S.rb * 128), am, act_full, (tc.kb0 + i) * 64, tc.r0);
        }
        // ---- epilogue: thread = output feature (TMEM lane), registers = activation rows ----
        if (ct == 0) tr(tagbase | 2);
        mbar_wait(tmem_full, tcount & 1);
        tcgen05_fence_after();
        if (ct == 0) tr(tagbase | 3);
        const int nch = S.rb >> 3;
        const int c_begin = half ? (nch >> 1) : 0, c_end = half ? nch : (nch >> 1);
#pragma unroll 1
        for (int ch = c_begin; ch < c_end; ++ch) {
          uint32_t r[8], r2[8];
          const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(ch * 8);
          tmem_ld_32x8(taddr, r);
          tmem_ld_32x8(taddr + kFAcc1, r2);
          tmem_ld_wait();
          float acc[8];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) acc[jj] = __uint_as_float(r[jj]) + __uint_as_float(r2[jj]);
          const int rowb = tc.r0 + ch * 8;
          if (!f_ok) {
            // feature tile reaches past the matrix (hid not a multiple of 128): nothing to store
          } else if (s == ST_QKV) {
            __nv_bfloat16* dst = qkvbuf + (int64_t)rowb * (3 * C) + f;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj)
              if (rowb + jj < tc.r1) dst[(int64_t)jj * (3 * C)] = __float2bfloat16_rn(acc[jj]);
          } else if (s == ST_FF1) {
            __nv_bfloat16* dst = a.ff + (int64_t)rowb * (4 * C) + f;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const float v = acc[jj] + bias;
              if (rowb + jj < tc.r1)
                dst[(int64_t)jj * (4 * C)] = __float2bfloat16_rn(0.5f * v * (1.f + erff(v * 0.70710678118654752440f)));
            }
          } else {                                         // PROJ / FFN2: x += D + bias (each element has one owner)
            float* dst = a.x + (int64_t)rowb * C + f;
            float old[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) old[jj] = (rowb + jj < tc.r1) ? __ldcg(dst + (int64_t)jj * C) : 0.f;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj)
              if (rowb + jj < tc.r1) dst[(int64_t)jj * C] = old[jj] + acc[jj] + bias;
          }
        }
        tcgen05_fence_before();
        bar_compute();
        if (ct == 0) tr(tagbase | 4);
        if (ct == 0) {                                     // bar.sync + fence.gpu: the CTA's stores are visible
          fence_gpu();
          if (s == ST_QKV) {                               // consumers are whole graphs: one counter per graph
            for (int g = 0; g < a.n_graphs; ++g)
              if (s_noff[g] < tc.r1 && s_noff[g + 1] > tc.r0)
                red_relaxed_gpu(a.cnt + (CNT_GRAPH * a.n_gran + g) * kFCntStride, 1);
          } else {
            const int which = s == ST_PROJ ? CNT_X2 : s == ST_FF1 ? CNT_FF : CNT_X;
            for (int gr = g0; gr < g1; ++gr) red_relaxed_gpu(a.cnt + (which * a.n_gran + gr) * kFCntStride, 1);
          }
          tr(tagbase | 5);
        }
      }
    }
  }

  __syncwarp();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tcgen05_fence_after();
    tmem_dealloc<2 * kFAcc1>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
int make_bf16_map(CUtensorMap* map, const void* base, int64_t rows, int64_t k, int box_rows);   // gemm_tcgen05.cu

template <int D>
static int launch_fused(const CUtensorMap* maps, const FusedKernelArgs& ka, int grid, cudaStream_t stream) {
  constexpr int smem = fused_smem_bytes();
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static bool configured = false;
  if (!configured) {
    GHN3_CUDA(cudaFuncSetAttribute(graphormer_fused_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  GHN3_CUDA(launch_pdl(graphormer_fused_kernel<D>, dim3((unsigned)grid), dim3(kFThreads), (size_t)smem, stream, maps[0],
                       maps[1], maps[2], maps[3], maps[4], maps[5], ka));
  GHN3_LAUNCH_CHECK("graphormer_fused_kernel");
  return GHN3_OK;
}

static long long* g_fused_trace = nullptr;

static int64_t fused_sync_ints(int total_nodes) {
  return (5 * ceil_div(std::max(total_nodes, 1), kFGran) + kFMaxGraphs) * kFCntStride;
}

int graphormer_fused_impl(const ghn3_graphormer_fused_args* a, cudaStream_t stream) {
  GHN3_REQUIRE(a != nullptr, "ghn3_graphormer_fused: null args");
  const int C = a->hid, H = a->heads, M = a->total_nodes;
  GHN3_REQUIRE(C > 0 && C % 64 == 0 && C <= 384, "ghn3_graphormer_fused: hid must be a multiple of 64, <= 384 (got %d)", C);
  GHN3_REQUIRE(H > 0 && C % H == 0, "ghn3_graphormer_fused: hid must be divisible by heads");
  const int D = C / H;
  if (M <= 0 || a->layers <= 0) return GHN3_OK;
  if (!(D == 8 || D == 16 || D == 24)) {
    set_error("ghn3_graphormer_fused: head dim %d is not supported (8, 16, 24)", D);
    return GHN3_ERR_UNSUPPORTED;
  }
  if (a->n_graphs > kFMaxGraphs || M > kFMaxGran * kFGran || a->lut_size * 4 > kFLutBytes) {
    set_error("ghn3_graphormer_fused: batch too large for the fused kernel (graphs %d, nodes %d, lut %d)", a->n_graphs, M,
              a->lut_size);
    return GHN3_ERR_UNSUPPORTED;
  }
  GHN3_REQUIRE(a->sync != nullptr && a->layers_dev != nullptr && a->x && a->ao && a->qkv && a->ff,
               "ghn3_graphormer_fused: null buffer");
  int grid = num_sms();
  if (a->max_ctas > 0) grid = std::min(grid, (int)a->max_ctas);

  FusedKernelArgs ka = {};
  ka.C = C; ka.H = H; ka.L = a->layers; ka.M = M; ka.n_graphs = a->n_graphs; ka.lut_size = a->lut_size;
  ka.n_gran = (int)ceil_div(M, kFGran);
  ka.total_stages = a->layers * 5;
  if (a->stop_after > 0) ka.total_stages = std::min(ka.total_stages, (int)a->stop_after);
  ka.node_off = a->node_off; ka.mat_off = a->mat_off; ka.pair = a->pair; ka.lut = a->lut;
  ka.layers = a->layers_dev;
  ka.x = a->x; ka.ao = (__nv_bfloat16*)a->ao; ka.qkv = (__nv_bfloat16*)a->qkv; ka.ff = (__nv_bfloat16*)a->ff;
  ka.cnt = a->sync;
  ka.trace = g_fused_trace;

  // rows per tile: the smallest multiple of 16 that keeps a stage within one wave of CTAs (capped at 64)
  auto make_stage = [&](int n_feat, int K, bool split) {
    FusedStage S = {};
    S.n_feat = n_feat;
    S.n_ft = (int)ceil_div(n_feat, 128);
    S.nkb = K / 64;
    S.splits = 1;                            // K blocks stream through the ring: no split-K, no atomics
    (void)split;
    S.kbps = (int)ceil_div(S.nkb, S.splits);
    // rows per tile: the whole activation operand of a tile (nkb blocks of rb x 128 B) lives in shared memory
    const int rb_cap = std::max(16, std::min(kFMaxRb, (kFActBytes / (S.nkb * 128)) / 16 * 16));
    int rb = 16;
    while (rb < rb_cap && ceil_div(M, rb) * S.n_ft * S.splits > grid) rb += 16;
    S.rb = rb;
    S.n_tiles = (int)ceil_div(M, rb) * S.n_ft * S.splits;
    S.pad = S.n_ft * S.splits;               // arrivals per granule and layer
    return S;
  };
  ka.st[ST_QKV] = make_stage(3 * C, C, false);
  ka.st[ST_PROJ] = make_stage(C, C, false);
  ka.st[ST_FF1] = make_stage(4 * C, C, false);
  ka.st[ST_FF2] = make_stage(C, 4 * C, true);
  GHN3_REQUIRE(4 * C / 64 * 16 * 128 <= kFActBytes, "ghn3_graphormer_fused: hid too large for the activation buffer");
  // attention: the number of units (H x 64-query blocks over all graphs) is derived from node_off inside the kernel;
  // what a unit waits for is the QKV stage's arrivals

  CUtensorMap maps[6];
  int rc;
  const int64_t L = a->layers;
  if ((rc = make_bf16_map(&maps[0], a->w_qkv, L * 3 * C, C, 128)) != GHN3_OK) return rc;
  if ((rc = make_bf16_map(&maps[1], a->w_out, L * C, C, 128)) != GHN3_OK) return rc;
  if ((rc = make_bf16_map(&maps[2], a->w_ff1, L * 4 * C, C, 128)) != GHN3_OK) return rc;
  if ((rc = make_bf16_map(&maps[3], a->w_ff2, L * C, 4 * C, 128)) != GHN3_OK) return rc;
  if ((rc = make_bf16_map(&maps[4], a->ao, M, C, ka.st[ST_PROJ].rb)) != GHN3_OK) return rc;
  if ((rc = make_bf16_map(&maps[5], a->ff, M, 4 * C, ka.st[ST_FF2].rb)) != GHN3_OK) return rc;

  GHN3_CUDA(cudaMemsetAsync(a->sync, 0, sizeof(int32_t) * (size_t)fused_sync_ints(M), stream));
  if (D == 8) return launch_fused<8>(maps, ka, grid, stream);
  if (D == 16) return launch_fused<16>(maps, ka, grid, stream);
  return launch_fused<24>(maps, ka, grid, stream);
}

}  // namespace ghn3

extern "C" int64_t ghn3_graphormer_fused_sync_ints(int32_t total_nodes) { return ghn3::fused_sync_ints(total_nodes); }

extern "C" int ghn3_graphormer_fused(const ghn3_graphormer_fused_args* args, ghn3_stream_t stream) {
  return ghn3::graphormer_fused_impl(args, (cudaStream_t)stream);
}

// Bring-up aid: device buffer of [n_ctas][1024][3] int64 (tag, clock64, globaltimer) records written by thread 64 of
// every CTA of subsequent ghn3_graphormer_fused launches (NULL = off).
extern "C" int ghn3_debug_fused_trace(void* device_buffer) {
  ghn3::g_fused_trace = (long long*)device_buffer;
  return GHN3_OK;
}
