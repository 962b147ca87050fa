// Fused tail of the 3-D discriminator (transfer_em/models/discriminator.py:72-99): the layers behind the last strided
// block,  d5 (3x3x3, C4 -> 32) -> d6 (4x4x4 stride 2, 32 -> 32, LeakyReLU twice: slope 0.09) -> d7 (1x1x1, 32 -> C4) ->
// d8 (1x1x1, C4 -> 1, bias)  on a 6^3 (or 7^3) x C4 activation, i.e. 64 -> 1 -> 1 -> 1 voxels per sample.
//
// As separate launches these four layers are 4 forward + 8 backward kernels per discriminator pass, each 10-60 us of
// launch, TMEM / pipeline set-up and atomics for < 4 MFLOP of work (profiles/tags_r1.txt: d5.wgrad 62 us at 4 GB/s,
// d6.dgrad 27 us for 0.04 GFLOP): ~1 ms of a train step.  Here FOUR CTAs per sample run the whole tail out of shared
// memory on CUDA cores (the work is ~2 MMAC per sample; no tensor cores, no TMA): one launch forward, one backward.
// Forward: the four CTAs of a sample form a thread-block cluster; CTA q computes 8 of d5's 32 output channels, the
// quarters meet in global memory behind a cluster barrier, CTA q then reduces one z-slab of d6's 4x4x4 window and sends
// its 32 partial sums to CTA 0 through distributed shared memory, which finishes d6, d7 and d8.
// Backward: no exchange is needed - every CTA recomputes the cheap chain g8 -> g7 -> g6 -> g5 and owns a quarter of the
// expensive parts (dW6 by z-slab, dW5 by output-channel group, the data gradient by input-channel group).
// A first version with one CTA per sample and plain loops took 157 / 51 us (backward / forward): every loop was a chain of
// dependent global-load round trips.  All weight reads are now issued in batches of independent 16 B loads.
// Numerics follow the layer-by-layer path: bf16-rounded weights, fp32 accumulation, activations and data gradients
// rounded to bf16 where that path stores them, LeakyReLU' taken from the sign of the stored activation.
#include <string.h>
#include "tem_kernels.cuh"

extern unsigned long long g_tem_launches;

namespace {

constexpr int kThreads = 256;
constexpr int NQ = 4;                            // CTAs per sample
constexpr int C5 = 32, C6 = 32;                  // literal widths of block "3" (discriminator.py:72)

__device__ __forceinline__ float bfr(float v) { return bf2f(__float2bfloat16_rn(v)); }
__device__ __forceinline__ float lrelu_f(float v, float s) { return v > 0.f ? v : v * s; }
__device__ __forceinline__ void red4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ int win_voxel(int tap, int e5) {      // voxel of a5 under tap (kz,ky,kx) of d6's 4x4x4 window at the origin
  return ((tap >> 4) * e5 + ((tap >> 2) & 3)) * e5 + (tap & 3);
}
__device__ __forceinline__ void stage_a4(const DiscTailArgs& a, int b, int nv4, bf16* a4s, int tid) {
  const uint4* src = reinterpret_cast<const uint4*>(a.a4 + (size_t)b * nv4 * a.C4);
  uint4* dst = reinterpret_cast<uint4*>(a4s);
  for (int i = tid; i < nv4 * a.C4 / 8; i += kThreads) dst[i] = __ldg(src + i);
}

// ---------------------------------------------------------------------------------------------------------------------
// forward (cluster of NQ CTAs per sample)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) disc_tail_fwd_kernel(const DiscTailArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int C4 = a.C4, e4 = a.e4, e5 = a.e5;
  const int nv4 = e4 * e4 * e4, nv5 = e5 * e5 * e5;
  bf16* a4s = reinterpret_cast<bf16*>(smem);                                  // [nv4][C4]
  bf16* w5q = a4s + ((nv4 * C4 + 7) & ~7);                                    // [27 * C4][8]: this CTA's 8 output channels of d5
  float* slab = reinterpret_cast<float*>(w5q + 27 * C4 * 8);                  // [16][32]: the z-slab of a5 under taps kz = q
  float* part = slab + 16 * C5;                                               // [8][32] K-slice partial sums of d6
  float* xsum = part + 8 * C6;                                                // [NQ][32] (CTA 0): the four slab sums
  float* a6s = xsum + NQ * C6;                                                // [32]
  float* a7s = a6s + C6;                                                      // [C4]
  const int tid = threadIdx.x, b = blockIdx.x, q = blockIdx.y;
  stage_a4(a, b, nv4, a4s, tid);
  {   // w5[k][32] -> w5q[k][8] for the channels 8q .. 8q+7: two 16 B loads per row, all issued before the first store
    const int n = 27 * C4 * 2;
    constexpr int R = 14;                                                     // C4 <= 64: 27 * 64 * 2 / 256 = 13.5
    float4 tmp[R];
#pragma unroll
    for (int j = 0; j < R; ++j) { const int i = tid + j * kThreads; if (i < n) tmp[j] = __ldg(reinterpret_cast<const float4*>(a.w5 + (size_t)(i >> 1) * C5 + 8 * q) + (i & 1)); }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int i = tid + j * kThreads;
      if (i < n) { uint2 pk; pk.x = pack2(tmp[j].x, tmp[j].y); pk.y = pack2(tmp[j].z, tmp[j].w); *reinterpret_cast<uint2*>(w5q + (size_t)(i >> 1) * 8 + 4 * (i & 1)) = pk; }
    }
  }
  __syncthreads();
  // d5, this CTA's 8 output channels: four lanes share a voxel and split the input channels in chunks of 8
  for (int it = 0; it < (nv5 * 4 + kThreads - 1) / kThreads; ++it) {
    const int item = tid + it * kThreads, v = item >> 2, sl = item & 3;
    const bool valid = v < nv5;
    float acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = 0.f;
    if (valid) {
      const int vz = v / (e5 * e5), vy = (v / e5) % e5, vx = v % e5;
      for (int tap = 0; tap < 27; ++tap) {
        const int u4 = ((vz + tap / 9) * e4 + vy + (tap / 3) % 3) * e4 + vx + tap % 3;
        for (int c0 = 8 * sl; c0 < C4; c0 += 32) {
          float xv[8]; unpack8(*reinterpret_cast<const uint4*>(a4s + u4 * C4 + c0), xv);
          const bf16* wp = w5q + (size_t)(tap * C4 + c0) * 8;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float wv[8]; unpack8(*reinterpret_cast<const uint4*>(wp + c * 8), wv);
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[u] = fmaf(xv[c], wv[u], acc[u]);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) { acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], 1); acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], 2); }
    if (valid && sl == 0) {
      float o[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) o[u] = lrelu_f(acc[u], a.slope5);
      uint4 pk; pk.x = pack2(o[0], o[1]); pk.y = pack2(o[2], o[3]); pk.z = pack2(o[4], o[5]); pk.w = pack2(o[6], o[7]);
      *reinterpret_cast<uint4*>(a.a5 + ((size_t)b * nv5 + v) * C5 + 8 * q) = pk;
    }
  }
  __threadfence();
  cluster_sync();                                   // the four channel quarters of a5 are in global memory
  // d6: one output voxel, K = 64 taps x 32 channels; this CTA owns the taps kz = q (16 voxels of a5, all 32 channels)
  for (int i = tid; i < 16 * C5 / 8; i += kThreads) {
    const int vi = i >> 2, c8 = i & 3;
    float f[8]; unpack8(*reinterpret_cast<const uint4*>(a.a5 + ((size_t)b * nv5 + win_voxel(16 * q + vi, e5)) * C5 + c8 * 8), f);
#pragma unroll
    for (int u = 0; u < 8; ++u) slab[vi * C5 + c8 * 8 + u] = f[u];
  }
  __syncthreads();
  {
    const int co = tid & 31, ks = tid >> 5;         // 8 K slices of 64 (tap, ci) pairs, eight weight loads in flight
    float acc = 0.f;
    const float* wq = a.w6 + (size_t)(q * 512 + ks * 64) * C6 + co;
    for (int k0 = 0; k0 < 64; k0 += 8) {
      float wv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) wv[j] = __ldg(wq + (size_t)(k0 + j) * C6);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(slab[ks * 64 + k0 + j], bfr(wv[j]), acc);      // slab index = (tap - 16q) * 32 + ci = k - 512 q
    }
    part[ks * C6 + co] = acc;
  }
  __syncthreads();
  if (tid < C6) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += part[i * C6 + tid];
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"((uint32_t)__cvta_generic_to_shared(xsum + q * C6 + tid)), "r"(0));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(s) : "memory");
  }
  cluster_sync();                                   // the four slab sums are in CTA 0's shared memory
  if (q != 0) return;
  if (tid < C6) {
    const float s = (xsum[tid] + xsum[C6 + tid]) + (xsum[2 * C6 + tid] + xsum[3 * C6 + tid]);
    const bf16 qv = __float2bfloat16_rn(lrelu_f(s, a.slope6));
    a.a6[(size_t)b * C6 + tid] = qv; a6s[tid] = bf2f(qv);
  }
  __syncthreads();
  if (tid < C4) {                                   // d7: 1x1x1, 32 -> C4
    float wv[C6];
#pragma unroll
    for (int ci = 0; ci < C6; ++ci) wv[ci] = __ldg(a.w7 + ci * C4 + tid);
    float s = 0.f;
#pragma unroll
    for (int ci = 0; ci < C6; ++ci) s = fmaf(a6s[ci], bfr(wv[ci]), s);
    const bf16 qv = __float2bfloat16_rn(lrelu_f(s, a.slope7));
    a.a7[(size_t)b * C4 + tid] = qv; a7s[tid] = bf2f(qv);
  }
  __syncthreads();
  if (tid < 32) {                                   // d8: 1x1x1, C4 -> 1, bias, linear, fp32 logit
    float s = 0.f;
    for (int ci = tid; ci < C4; ci += 32) s = fmaf(a7s[ci], bfr(__ldg(a.w8 + ci)), s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (tid == 0) a.logits[b] = s + (a.b8 ? a.b8[0] : 0.f);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward: data gradient w.r.t. a4 (LeakyReLU' of a4 applied) and, optionally, the weight gradients of d5..d8
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) disc_tail_bwd_kernel(const DiscTailArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int C4 = a.C4, e4 = a.e4, e5 = a.e5, C4q = a.C4 / NQ;
  const int nv4 = e4 * e4 * e4, nv5 = e5 * e5 * e5;
  bf16* a4s = reinterpret_cast<bf16*>(smem);                                  // [nv4][C4]
  bf16* w5t = a4s + ((nv4 * C4 + 7) & ~7);                                    // [27][32 co][C4q ci]: this CTA's input channels, transposed
  float* a5s = reinterpret_cast<float*>(w5t + 27 * C5 * C4q);                 // [nv5][32]
  float* g5s = a5s + nv5 * C5;                                                // [nv5][32] (bf16-rounded)
  float* a6s = g5s + nv5 * C5;                                                // [32]
  float* g6s = a6s + C6;                                                      // [32]
  float* a7s = g6s + C6;                                                      // [C4]
  float* g7s = a7s + C4;                                                      // [C4]
  const int tid = threadIdx.x, b = blockIdx.x, q = blockIdx.y;
  const bool wg = a.dw5 != nullptr;
  stage_a4(a, b, nv4, a4s, tid);
  {   // w5[tap][q C4q + ci][co] -> w5t[tap][co][ci]: rows of 32 output channels are read as 8 x 16 B, all loads first
    const int n = 27 * C4q * 8;
    constexpr int R = 14;                                                     // C4 <= 64: 27 * 16 * 8 / 256 = 13.5
    float4 tmp[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int i = tid + j * kThreads;
      if (i < n) { const int row = i >> 3, tap = row / C4q, ci = row % C4q; tmp[j] = __ldg(reinterpret_cast<const float4*>(a.w5 + ((size_t)tap * C4 + q * C4q + ci) * C5) + (i & 7)); }
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int i = tid + j * kThreads;
      if (i < n) {
        const int row = i >> 3, tap = row / C4q, ci = row % C4q, co = 4 * (i & 7);
        bf16* d = w5t + ((size_t)tap * C5 + co) * C4q + ci;
        d[0] = __float2bfloat16_rn(tmp[j].x); d[C4q] = __float2bfloat16_rn(tmp[j].y); d[2 * C4q] = __float2bfloat16_rn(tmp[j].z); d[3 * C4q] = __float2bfloat16_rn(tmp[j].w);
      }
    }
  }
  for (int i = tid; i < nv5 * C5 / 8; i += kThreads) {
    float f[8]; unpack8(__ldg(reinterpret_cast<const uint4*>(a.a5 + (size_t)b * nv5 * C5) + i), f);
#pragma unroll
    for (int u = 0; u < 8; ++u) { a5s[i * 8 + u] = f[u]; g5s[i * 8 + u] = 0.f; }
  }
  if (tid < C6) a6s[tid] = bf2f(a.a6[(size_t)b * C6 + tid]);
  if (tid < C4) a7s[tid] = bf2f(a.a7[(size_t)b * C4 + tid]);
  __syncthreads();
  const float g8 = a.dlogits[b];
  // d8: dW8[ci] += a7[ci] g8, db8 += g8;   g7[ci] = g8 w8[ci] lrelu'(a7[ci])
  if (tid < C4) {
    if (wg && q == 0) atomicAdd(a.dw8 + tid, a7s[tid] * g8);
    g7s[tid] = bfr(g8 * bfr(__ldg(a.w8 + tid)) * (a7s[tid] > 0.f ? 1.f : a.slope7));
  }
  if (wg && q == 0 && tid == 0 && a.db8) atomicAdd(a.db8, g8);
  __syncthreads();
  // d7: dW7[ci][co] += a6[ci] g7[co];   g6[ci] = sum_co g7[co] w7[ci][co] lrelu'(a6[ci])
  if (wg && q == 0) for (int i = tid; i < C6 * C4; i += kThreads) atomicAdd(a.dw7 + i, a6s[i / C4] * g7s[i % C4]);
  if (tid < C6) {
    float s = 0.f;
    for (int c0 = 0; c0 < C4; c0 += 32) {
      float4 wv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) wv[j] = __ldg(reinterpret_cast<const float4*>(a.w7 + tid * C4 + c0) + j);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s = fmaf(g7s[c0 + 4 * j], bfr(wv[j].x), s); s = fmaf(g7s[c0 + 4 * j + 1], bfr(wv[j].y), s);
        s = fmaf(g7s[c0 + 4 * j + 2], bfr(wv[j].z), s); s = fmaf(g7s[c0 + 4 * j + 3], bfr(wv[j].w), s);
      }
    }
    g6s[tid] = bfr(s * (a6s[tid] > 0.f ? 1.f : a.slope6));
  }
  __syncthreads();
  // d6 data gradient (every CTA, all rows): g5[v(tap)][ci] = sum_co g6[co] w6[tap][ci][co] lrelu'(a5); one row per thread step
  for (int k = tid; k < 64 * C5; k += kThreads) {
    const float4* wp = reinterpret_cast<const float4*>(a.w6 + (size_t)k * C6);
    float4 wv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) wv[j] = __ldg(wp + j);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s = fmaf(g6s[4 * j], bfr(wv[j].x), s); s = fmaf(g6s[4 * j + 1], bfr(wv[j].y), s);
      s = fmaf(g6s[4 * j + 2], bfr(wv[j].z), s); s = fmaf(g6s[4 * j + 3], bfr(wv[j].w), s);
    }
    const int v = win_voxel(k >> 5, e5), ci = k & 31;
    g5s[v * C5 + ci] = bfr(s * (a5s[v * C5 + ci] > 0.f ? 1.f : a.slope5));
  }
  // d6 weight gradient, this CTA's z-slab of taps: dW6[tap][ci][co] += a5[v(tap)][ci] g6[co]
  if (wg) {
    for (int i = tid; i < 512 * (C6 / 4); i += kThreads) {
      const int c4 = i & 7, k = q * 512 + (i >> 3);
      const float x = a5s[win_voxel(k >> 5, e5) * C5 + (k & 31)];
      red4(a.dw6 + (size_t)k * C6 + c4 * 4, x * g6s[c4 * 4], x * g6s[c4 * 4 + 1], x * g6s[c4 * 4 + 2], x * g6s[c4 * 4 + 3]);
    }
  }
  __syncthreads();
  // d5 weight gradient, output channels 8q .. 8q+7: dW5[tap][ci][co] += sum_v a4[v + tap][ci] g5[v][co];   item = (tap, ci)
  if (wg) {
    for (int item = tid; item < 27 * C4; item += kThreads) {
      const int ci = item % C4, tap = item / C4;
      const int tz = tap / 9, ty = (tap / 3) % 3, tx = tap % 3;
      float acc[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] = 0.f;
      for (int v = 0; v < nv5; ++v) {
        const int vz = v / (e5 * e5), vy = (v / e5) % e5, vx = v % e5;
        const float x = bf2f(a4s[(((vz + tz) * e4 + vy + ty) * e4 + vx + tx) * C4 + ci]);
        const float4 g0 = *reinterpret_cast<const float4*>(g5s + v * C5 + q * 8), g1 = *reinterpret_cast<const float4*>(g5s + v * C5 + q * 8 + 4);
        acc[0] = fmaf(x, g0.x, acc[0]); acc[1] = fmaf(x, g0.y, acc[1]); acc[2] = fmaf(x, g0.z, acc[2]); acc[3] = fmaf(x, g0.w, acc[3]);
        acc[4] = fmaf(x, g1.x, acc[4]); acc[5] = fmaf(x, g1.y, acc[5]); acc[6] = fmaf(x, g1.z, acc[6]); acc[7] = fmaf(x, g1.w, acc[7]);
      }
      float* dst = a.dw5 + ((size_t)tap * C4 + ci) * C5 + q * 8;
      red4(dst, acc[0], acc[1], acc[2], acc[3]); red4(dst + 4, acc[4], acc[5], acc[6], acc[7]);
    }
  }
  // d5 data gradient, input channels q C4q ..: g4[u][ci] = lrelu'(a4[u][ci]) sum_{tap, co} g5[u - tap][co] w5[tap][ci][co];  item = (u, 8 channels)
  for (int item = tid; item < nv4 * (C4q / 8); item += kThreads) {
    const int cg = item % (C4q / 8), u = item / (C4q / 8);
    const int uz = u / (e4 * e4), uy = (u / e4) % e4, ux = u % e4;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
    for (int tap = 0; tap < 27; ++tap) {
      const int vz = uz - tap / 9, vy = uy - (tap / 3) % 3, vx = ux - tap % 3;
      if (vz < 0 || vz >= e5 || vy < 0 || vy >= e5 || vx < 0 || vx >= e5) continue;
      const float* gp = g5s + ((vz * e5 + vy) * e5 + vx) * C5;
      const bf16* wp = w5t + (size_t)(tap * C5) * C4q + cg * 8;
#pragma unroll 4
      for (int co = 0; co < C5; ++co) {
        const float gv = gp[co];
        float wv[8]; unpack8(*reinterpret_cast<const uint4*>(wp + co * C4q), wv);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = fmaf(gv, wv[c], acc[c]);
      }
    }
    const int c0 = q * C4q + cg * 8;
    float xv[8]; unpack8(*reinterpret_cast<const uint4*>(a4s + u * C4 + c0), xv);
    float o[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) o[c] = acc[c] * (xv[c] > 0.f ? 1.f : a.slope4);
    uint4 pk;
    pk.x = pack2(o[0], o[1]); pk.y = pack2(o[2], o[3]); pk.z = pack2(o[4], o[5]); pk.w = pack2(o[6], o[7]);
    *reinterpret_cast<uint4*>(a.d_a4 + ((size_t)b * nv4 + u) * C4 + c0) = pk;
  }
}

size_t tail_smem(const DiscTailArgs& a, bool bwd) {
  const size_t nv4 = (size_t)a.e4 * a.e4 * a.e4, nv5 = (size_t)a.e5 * a.e5 * a.e5;
  size_t s = ((nv4 * a.C4 + 7) & ~(size_t)7) * 2;
  if (!bwd) s += (size_t)27 * a.C4 * 8 * 2 + (16 * C5 + 8 * C6 + NQ * C6 + C6 + a.C4) * 4;
  else s += (size_t)27 * C5 * (a.C4 / NQ) * 2 + (2 * nv5 * C5 + 2 * C6 + 2 * a.C4) * 4;
  return s + 64;
}

}  // namespace

bool disc_tail_supported(const DiscTailArgs& a) {
  if (a.C4 % 32 || a.C4 < 32 || a.C4 > 64) return false;                  // wf = 8 / 4 (a CTA owns C4 / 4 input channels in chunks of 8)
  if (a.e5 != a.e4 - 2 || a.e5 < 4 || a.e5 > 5) return false;             // d6 (4x4x4 stride 2) must reduce to ONE voxel
  const uintptr_t al = reinterpret_cast<uintptr_t>(a.a4) | reinterpret_cast<uintptr_t>(a.a5) | reinterpret_cast<uintptr_t>(a.d_a4) |
                       reinterpret_cast<uintptr_t>(a.dw5) | reinterpret_cast<uintptr_t>(a.dw6) | reinterpret_cast<uintptr_t>(a.w5) |
                       reinterpret_cast<uintptr_t>(a.w6) | reinterpret_cast<uintptr_t>(a.w7);
  if (al & 15) return false;                                               // 16 B vector loads / stores / reductions
  return tail_smem(a, true) <= 200 * 1024 && tail_smem(a, false) <= 200 * 1024;
}

cudaError_t launch_disc_tail_fwd(const DiscTailArgs& a, cudaStream_t st) {
  static bool attr = false;
  if (!attr) { cudaError_t e = cudaFuncSetAttribute(disc_tail_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); if (e) return e; attr = true; }
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)a.B, NQ, 1); cfg.blockDim = dim3(kThreads, 1, 1); cfg.dynamicSmemBytes = tail_smem(a, false); cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = NQ; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, disc_tail_fwd_kernel, a); ++g_tem_launches;
  return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t launch_disc_tail_bwd(const DiscTailArgs& a, cudaStream_t st) {
  static bool attr = false;
  if (!attr) { cudaError_t e = cudaFuncSetAttribute(disc_tail_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); if (e) return e; attr = true; }
  disc_tail_bwd_kernel<<<dim3((unsigned)a.B, NQ), kThreads, tail_smem(a, true), st>>>(a); ++g_tem_launches;
  return cudaGetLastError();
}
