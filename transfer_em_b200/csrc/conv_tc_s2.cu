// tcgen05 implicit-GEMM kernels for the stride-2 layers of the reference graph (4x4x4 kernels):
//   DOWN (form 0): out[o] = sum_k in[2o + k - pad] w[k]          strided downsample conv forward (models/utils.py:80)
//                                                                and the data gradient of Conv3DTranspose (:129-130)
//   UP   (form 1): out[j] = sum_{k = (j+pad) mod 2} in[(j+pad-k)/2] w[k]
//                                                                Conv3DTranspose(4, 2, 'same') forward and the data
//                                                                gradient of the strided conv
//
// Both are sums of 2x2x2-tap stride-1 correlations once the stride is folded into a parity split:
//   DOWN: k - pad = 2m + r  ->  out[o] = sum_r sum_m in_r[o + m] w[2m + r + pad],  in_r = every second voxel of `in`.
//         The de-interleaved halo tiles in_r are produced by TMA itself (tensor map with elementStrides = 2 along x and
//         y; z is the slice coordinate), so a tap is again only a shifted UMMA descriptor start address.
//         GEMM: M = 128 output voxels (16 y x 8 x of one output z-slice), N = Cout, K = 64 taps x Cin.
//   UP:   j + pad = 2q + r  ->  out[2q + r - pad] = sum_{m'} in[q - 1 + m'] w[r + 2(1 - m')]: every output parity class
//         r = (rz,ry,rx) is a 2x2x2 correlation over the SAME input window, so the eight classes become the N dimension:
//         GEMM: M = 128 q-voxels, N = 8 classes x Cout, K = 8 taps x Cin; the epilogue scatters class r to voxel 2q+r-pad.
// Everything else follows conv_tc.cu: one CTA marches along z over an (x,y) tile column, input slices are staged once
// by TMA into an mbarrier ring (OOB zero-fill = padding), bf16 UMMA weight images are loaded once per CTA, accumulators
// are double-buffered in TMEM, warp roles = TMA producer / MMA issuer / 4 epilogue warps with the fused
// LeakyReLU / LeakyReLU' x dropout / accumulate epilogue.
#include <cuda.h>
#include <string.h>
#include <stdlib.h>
#include "tem_kernels.cuh"
#include "ptx_sm100.cuh"

extern unsigned long long g_tem_launches;

namespace {

constexpr int TX = 8, TY = 16;
constexpr int SXV = TX + 1, SYR = TY + 1;          // sub-tile: 17 rows x 9 voxels of one 8-channel plane
constexpr int SUB_BYTES = SYR * SXV * 16;          // 2448
constexpr int SUB_STRIDE = 2560;                   // padded to a multiple of 128 B (TMA destination alignment)
constexpr int ROW_B = SXV * 16;                    // 144: y-row pitch = stride between 8-row core-matrix groups
constexpr int RING_MAX = 8;
constexpr int kThreads = 192;
constexpr int kThreadsUp = 320;              // UP: two epilogue warps per TMEM lane quadrant (one per output z parity)

struct S2Args {
  int B, L[3];                 // logical output extent (z,y,x)
  int Q[3];                    // UP: q extents; DOWN: = L
  int pad;                     // 0 or 1 (all axes)
  int planes;                  // Cin / 8
  int merged;                  // UP: map folds (channel, x) when Cin == 8
  int shift[3];                // tensor coordinate = conv-input coordinate + shift
  int cin8;
  const bf16* wpacked; int wbytes;
  int ntx, nty, nzc, zc, ring;      // ring: input slots (resident-weight kernels) / K-steps per weight stage (wide DOWN)
  int np;                            // wide DOWN: MMA N (columns per output slice)
  int dbg;                           // experiment bits (TEM_S2_DBG): 1 no epilogue memory traffic, 2 no weight loads, 4 no input loads
  int swz;                           // wide kernels: input tiles are [voxel][64 ch] rows in the 128B-swizzle layout (one TMA request per voxel)
  bf16* out; int OZ, OY, OX, out_C, out_coff, out_off[3];
  int Cout;
  float slope;
  const bf16* ref; int RZ, RY, RX, ref_C, ref_coff, ref_off[3]; float ref_slope;
  uint32_t drop_key;
  int accumulate;
};

struct Epi {   // fused epilogue on 8 channels of one output voxel; refq / accq: operands fetched ahead by the caller
  __device__ static __forceinline__ long long ref_off(const S2Args& a, int b, int oz, int oy, int ox) {
    return ((((long long)b * a.RZ + oz + a.ref_off[0]) * a.RY + oy + a.ref_off[1]) * a.RX + ox + a.ref_off[2]) * a.ref_C + a.ref_coff;
  }
  __device__ static __forceinline__ long long out_off(const S2Args& a, int b, int oz, int oy, int ox) {
    return ((((long long)b * a.OZ + oz + a.out_off[0]) * a.OY + oy + a.out_off[1]) * a.OX + ox + a.out_off[2]) * a.out_C + a.out_coff;
  }
  // MODE (resident-weight kernels): 0 forward, 1 forward with dropout, 2 data gradient (LeakyReLU' mask, optional accumulate),
  // 3 everything decided at run time.  Carrying all variants made conv_down_tc 110 KB of code for eight epilogue warps.
  template <int MODE>
  __device__ static __forceinline__ void run(const S2Args& a, float* v, int c0, int b, int oz, int oy, int ox, const uint4& refq, const uint4& accq) {
    if (MODE >= 2 && a.ref) {
      float f[8];
      unpack8(refq, f);
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] *= (f[u] > 0.f) ? 1.f : a.ref_slope;
    }
    if ((MODE == 1 || MODE == 3) && a.drop_key) {
      const uint32_t di = (uint32_t)(((((long long)b * a.L[0] + oz) * a.L[1] + oy) * a.L[2] + ox) * a.Cout) + (uint32_t)c0;
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] *= 2.f * tem_keep(a.drop_key, di + u);
    }
    float o[8];
    if (MODE >= 2 && a.accumulate) unpack8(accq, o);
    else {
#pragma unroll
      for (int u = 0; u < 8; ++u) o[u] = 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      o[u] += v[u];
      if (MODE != 2 && a.slope != 1.f) o[u] = o[u] > 0.f ? o[u] : o[u] * a.slope;
    }
    uint4 pk;
    pk.x = pack2(o[0], o[1]); pk.y = pack2(o[2], o[3]); pk.z = pack2(o[4], o[5]); pk.w = pack2(o[6], o[7]);
    if (!(a.dbg & 32)) *reinterpret_cast<uint4*>(a.out + out_off(a, b, oz, oy, ox) + c0) = pk;
  }
};

// the same epilogue on a precomputed output pointer / dropout index (index arithmetic hoisted by the caller); LeakyReLU' of
// the reference is taken from the raw bf16 halves: ref > 0 <=> sign bit clear and magnitude non-zero
template <int MODE>
__device__ __forceinline__ void epi_finish(const S2Args& a, float* v, const uint4& rq, const uint4& aq, bf16* op, uint32_t di) {
  if (MODE >= 2 && a.ref) {
    const uint32_t w[4] = {rq.x, rq.y, rq.z, rq.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (!((w[i] & 0x7fffu) != 0u && (w[i] & 0x8000u) == 0u)) v[2 * i] *= a.ref_slope;
      if (!((w[i] & 0x7fff0000u) != 0u && (w[i] & 0x80000000u) == 0u)) v[2 * i + 1] *= a.ref_slope;
    }
  }
  if ((MODE == 1 || MODE == 3) && a.drop_key) {
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] *= 2.f * tem_keep(a.drop_key, di + (uint32_t)u);
  }
  if (MODE >= 2 && a.accumulate) {
    float o[8]; unpack8(aq, o);
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] += o[u];
  }
  if (MODE != 2 && a.slope != 1.f) {
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = v[u] > 0.f ? v[u] : v[u] * a.slope;
  }
  uint4 pk;
  pk.x = pack2(v[0], v[1]); pk.y = pack2(v[2], v[3]); pk.z = pack2(v[4], v[5]); pk.w = pack2(v[6], v[7]);
  if (!(a.dbg & 32)) *reinterpret_cast<uint4*>(op) = pk;
}

__device__ __forceinline__ void decode_work(const S2Args& a, int& b, int& x0, int& y0, int& z0, int& nz) {
  int w = blockIdx.x;
  const int zc_i = w % a.nzc; w /= a.nzc;
  const int tx_i = w % a.ntx; w /= a.ntx;
  const int ty_i = w % a.nty; w /= a.nty;
  b = w;
  x0 = tx_i * TX; y0 = ty_i * TY; z0 = zc_i * a.zc;
  nz = min(a.zc, a.Q[0] - z0);
}

// ------------------------------------------------------------------------------------------------
// UP: NP = 8 * CP accumulator columns per q-slice (class-major), two TMEM stages
// ------------------------------------------------------------------------------------------------
template <int CP, int MODE>
__global__ void __launch_bounds__(kThreadsUp)
conv_up_tc_kernel(const __grid_constant__ CUtensorMap map0, const S2Args a) {
  constexpr int NP = 8 * CP;
  constexpr uint32_t kTmemCols = 2 * NP;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[RING_MAX], empty_bar[RING_MAX], w_bar, tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_s;
  const int RING = a.ring;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int planes = a.planes;
  const uint32_t wbytes_pad = (uint32_t)((a.wbytes + 1023) & ~1023);
  uint8_t* wsm = smem;
  uint8_t* ring = smem + wbytes_pad;
  const uint32_t slot_bytes = (uint32_t)planes * SUB_STRIDE;

  if (threadIdx.x == 0) {
    for (int i = 0; i < RING; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&w_bar, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  int b, x0, y0, z0, nz; decode_work(a, b, x0, y0, z0, nz);
  const int nslices = nz + 1;                      // input slices z0-1 .. z0+nz-1

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&w_bar, (uint32_t)a.wbytes);
      bulk_load(wsm, a.wpacked, (uint32_t)a.wbytes, &w_bar);
      int slot = 0; uint32_t ph = 0;
      for (int s = 0; s < nslices; ++s) {
        mbar_wait(&empty_bar[slot], ph ^ 1u);
        if (a.dbg & 4) { mbar_arrive(&full_bar[slot]); if (++slot == RING) { slot = 0; ph ^= 1u; } continue; }
        mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)planes * SUB_BYTES);
        uint8_t* dst = ring + (size_t)slot * slot_bytes;
        for (int p = 0; p < planes; ++p)
          tma_load_plane(dst + p * SUB_STRIDE, &map0, &full_bar[slot], a.merged, p, x0 - 1 + a.shift[2], y0 - 1 + a.shift[1], z0 - 1 + s + a.shift[0], b);
        if (++slot == RING) { slot = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    {   // convergent issue loop: the whole warp waits, one elected lane issues the MMAs / commits of a q-slice
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      mbar_wait(&w_bar, 0);
      const uint32_t wb16 = smem_u32(wsm) >> 4;
      const uint32_t rbase = smem_u32(ring);
      const uint32_t a_hi = ((uint32_t)ROW_B >> 4) | (1u << 14), b_hi = (128u >> 4) | (1u << 14);
      const uint32_t b_lbo = ((uint32_t)(NP * 16) >> 4) << 16;
      const uint32_t a_lbo = a.cin8 ? (1u << 16) : (((uint32_t)SUB_STRIDE >> 4) << 16);   // K halves: next voxel / next plane
      int waited = 0, wslot = 0, zslot = 0; uint32_t wph = 0;
      const int kcs = planes >> 1;
      for (int zo = 0; zo < nz; ++zo) {
        while (waited < zo + 2) { mbar_wait(&full_bar[wslot], wph); ++waited; if (++wslot == RING) { wslot = 0; wph ^= 1u; } }
        mbar_wait(&tempty_bar[zo & 1], (((uint32_t)(zo >> 1)) & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(zo & 1) * NP;
        int sl1 = zslot + 1; if (sl1 >= RING) sl1 -= RING;
        if (elect_one()) {
          uint32_t acc = 0;
          uint32_t blo = wb16 | b_lbo;
#pragma unroll
          for (int mz = 0; mz < ((a.dbg & 16) ? 0 : 2); ++mz) {
            const uint32_t sb16 = (rbase + (uint32_t)(mz ? sl1 : zslot) * slot_bytes) >> 4;
#pragma unroll
            for (int my = 0; my < 2; ++my) {
              if (a.cin8) {          // one plane: the K=16 step covers the taps mx = 0, 1 (adjacent voxels)
                umma_bf16(d_tmem, ((uint64_t)a_hi << 32) | ((sb16 + (uint32_t)(my * SXV)) | a_lbo), ((uint64_t)b_hi << 32) | blo, idesc, acc);
                acc = 1; blo += (uint32_t)(NP * 32) >> 4;
              } else {
#pragma unroll
                for (int mx = 0; mx < 2; ++mx) {
                  uint32_t alo = (sb16 + (uint32_t)(my * SXV + mx)) | a_lbo;
                  for (int kc = 0; kc < kcs; ++kc) {
                    umma_bf16(d_tmem, ((uint64_t)a_hi << 32) | alo, ((uint64_t)b_hi << 32) | blo, idesc, acc);
                    acc = 1; alo += (uint32_t)(2 * SUB_STRIDE) >> 4; blo += (uint32_t)(NP * 32) >> 4;
                  }
                }
              }
            }
          }
          umma_commit(&tfull_bar[zo & 1]);
          umma_commit(&empty_bar[zslot]);      // input slice zo-1 is not used by later outputs
        }
        __syncwarp();
        if (++zslot == RING) zslot = 0;
      }
    }
  } else {
    const int q = warp & 3;                        // TMEM lane quadrant of this warp
    const int rz = (warp - 2) >> 2;                // output z parity handled by this warp
    const int row = q * 32 + lane;
    const int yl = row >> 3, xl = row & 7;
    constexpr int NCH = CP / 8;
    constexpr bool kPrefetch = CP <= 16;           // operands of all four (ry,rx) classes are fetched before the accumulator wait
    // measured (TEM_S2_DBG=1): this epilogue, not the MMAs or the loads, bounds the kernel (g2.dgrad 49.9 us, 20.2 us without
    // it), and its cost is index arithmetic: a thread owns ONE q-voxel, so class offsets, validity and z strides are hoisted
    const int oy0 = 2 * (y0 + yl) - a.pad, ox0 = 2 * (x0 + xl) - a.pad, oz0 = 2 * z0 + rz - a.pad;
    bool okc[4]; int ooff[4], roff[4]; uint32_t doff[4];
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
      const int oy = oy0 + (c4 >> 1), ox = ox0 + (c4 & 1);
      okc[c4] = !(a.dbg & 1) && oy >= 0 && oy < a.L[1] && ox >= 0 && ox < a.L[2];
      ooff[c4] = ((c4 >> 1) * a.OX + (c4 & 1)) * a.out_C;
      roff[c4] = ((c4 >> 1) * a.RX + (c4 & 1)) * a.ref_C;
      doff[c4] = (uint32_t)(((c4 >> 1) * a.L[2] + (c4 & 1)) * a.Cout);
    }
    const long long o_zs = 2LL * a.OY * a.OX * a.out_C, r_zs = 2LL * a.RY * a.RX * a.ref_C;      // two output slices per q-slice
    bf16* const out0 = a.out + Epi::out_off(a, b, oz0, oy0, ox0);
    const bf16* const ref0 = a.ref ? a.ref + Epi::ref_off(a, b, oz0, oy0, ox0) : a.out;        // never read when a.ref == nullptr
    const uint32_t di0 = (uint32_t)(((((long long)b * a.L[0] + oz0) * a.L[1] + oy0) * a.L[2] + ox0) * a.Cout);
    const uint32_t di_zs = (uint32_t)(2 * a.L[1] * a.L[2] * a.Cout);
    for (int zo = 0; zo < nz; ++zo) {
      const int oz = oz0 + 2 * zo;
      const bool zv = oz >= 0 && oz < a.L[0];
      bf16* const outz = out0 + zo * o_zs;
      const bf16* const refz = ref0 + zo * r_zs;
      uint4 refq[kPrefetch ? 4 : 1][NCH], accq[kPrefetch ? 4 : 1][NCH];
      if (kPrefetch) {
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          if (zv && okc[c4]) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
              if (c * 8 < a.Cout) {
                if (MODE >= 2 && a.ref) refq[c4][c] = __ldg(reinterpret_cast<const uint4*>(refz + roff[c4] + c * 8));
                if (MODE >= 2 && a.accumulate) accq[c4][c] = *reinterpret_cast<const uint4*>(outz + ooff[c4] + c * 8);
              }
            }
          }
        }
      }
      mbar_wait(&tfull_bar[zo & 1], ((uint32_t)(zo >> 1)) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(zo & 1) * NP + (uint32_t)(rz * 4 * CP);
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        __syncwarp();
        const bool ok = zv && okc[c4];
        if (!kPrefetch && ok) {
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            if (c * 8 < a.Cout) {
              if (MODE >= 2 && a.ref) refq[0][c] = __ldg(reinterpret_cast<const uint4*>(refz + roff[c4] + c * 8));
              if (MODE >= 2 && a.accumulate) accq[0][c] = *reinterpret_cast<const uint4*>(outz + ooff[c4] + c * 8);
            }
          }
        }
        uint32_t r[CP];
#pragma unroll
        for (int c = 0; c < CP; c += 8) tmem_ld8(taddr + (uint32_t)(c4 * CP + c), r + c);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c4 == 3) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          mbar_arrive(&tempty_bar[zo & 1]);
        }
        if (ok) {
          const uint32_t di = di0 + (uint32_t)zo * di_zs + doff[c4];
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            if (c * 8 < a.Cout) {
              float v[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) v[u] = __uint_as_float(r[c * 8 + u]);
              epi_finish<MODE>(a, v, refq[kPrefetch ? c4 : 0][c], accq[kPrefetch ? c4 : 0][c], outz + ooff[c4] + c * 8, di + (uint32_t)(c * 8));
            }
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// DOWN: one ring slot = one input z-slice = 4 (ry,rx) de-interleaved sub-tiles x planes
// ------------------------------------------------------------------------------------------------
template <int NPAD, int MODE>
__global__ void __launch_bounds__(kThreads)
conv_down_tc_kernel(const __grid_constant__ CUtensorMap map0, const S2Args a) {
  constexpr uint32_t kTmemCols = (2 * NPAD < 32) ? 32 : 2 * NPAD;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[RING_MAX], empty_bar[RING_MAX], w_bar, tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_s;
  const int RING = a.ring;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int planes = a.planes;
  const uint32_t wbytes_pad = (uint32_t)((a.wbytes + 1023) & ~1023);
  uint8_t* wsm = smem;
  uint8_t* ring = smem + wbytes_pad;
  const uint32_t slot_bytes = (uint32_t)(4 * planes) * SUB_STRIDE;

  if (threadIdx.x == 0) {
    for (int i = 0; i < RING; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&w_bar, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  int b, x0, y0, z0, nz; decode_work(a, b, x0, y0, z0, nz);
  const int nslices = 2 * nz + 2;                  // input slices 2*z0 - pad .. 2*(z0+nz-1) + 3 - pad

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&w_bar, (uint32_t)a.wbytes);
      bulk_load(wsm, a.wpacked, (uint32_t)a.wbytes, &w_bar);
      int slot = 0; uint32_t ph = 0;
      for (int s = 0; s < nslices; ++s) {
        mbar_wait(&empty_bar[slot], ph ^ 1u);
        if (a.dbg & 4) { mbar_arrive(&full_bar[slot]); if (++slot == RING) { slot = 0; ph ^= 1u; } continue; }
        mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)(4 * planes) * SUB_BYTES);
        uint8_t* dst = ring + (size_t)slot * slot_bytes;
        const int zin = 2 * z0 - a.pad + s + a.shift[0];
        for (int rr = 0; rr < 4; ++rr) {
          const int ry = rr >> 1, rx = rr & 1;
          // parity r of (k - pad): k = kb, kb + 2 with kb = (r + pad) & 1; first tap offset m_lo = (kb - pad - r) / 2
          const int mly = (((ry + a.pad) & 1) - a.pad - ry) / 2, mlx = (((rx + a.pad) & 1) - a.pad - rx) / 2;
          const int cy = 2 * (y0 + mly) + ry + a.shift[1], cx = 2 * (x0 + mlx) + rx + a.shift[2];
          for (int p = 0; p < planes; ++p)
            tma_load_5d(dst + (rr * planes + p) * SUB_STRIDE, &map0, &full_bar[slot], p * 8, cx, cy, zin, b);
        }
        if (++slot == RING) { slot = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    {   // convergent issue loop (see the UP kernel)
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NPAD >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      mbar_wait(&w_bar, 0);
      const uint32_t wb16 = smem_u32(wsm) >> 4;
      const uint32_t rbase = smem_u32(ring);
      const uint32_t a_hi = ((uint32_t)ROW_B >> 4) | (1u << 14), b_hi = (128u >> 4) | (1u << 14);
      const uint32_t b_lbo = ((uint32_t)(NPAD * 16) >> 4) << 16;
      const uint32_t a_lbo = a.cin8 ? (1u << 16) : (((uint32_t)SUB_STRIDE >> 4) << 16);
      int waited = 0, wslot = 0, zslot = 0; uint32_t wph = 0;
      const int kcs = planes >> 1;
      for (int zo = 0; zo < nz; ++zo) {
        while (waited < 2 * zo + 4) { mbar_wait(&full_bar[wslot], wph); ++waited; if (++wslot == RING) { wslot = 0; wph ^= 1u; } }
        mbar_wait(&tempty_bar[zo & 1], (((uint32_t)(zo >> 1)) & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(zo & 1) * NPAD;
        int s1 = zslot + 1; if (s1 >= RING) s1 -= RING;
        if (elect_one()) {
          uint32_t acc = 0;
          uint32_t blo = wb16 | b_lbo;
          // kz and the (ry, rx) class stay rolled: fully unrolled (32 bodies with both the Cin = 8 and the generic path) this loop
          // alone was ~6000 instructions, i.e. the single issuing warp ran out of a 90 KB instruction footprint
#pragma unroll 1
          for (int kz = 0; kz < ((a.dbg & 16) ? 0 : 4); ++kz) {
            int sl = zslot + kz; if (sl >= RING) sl -= RING;
            const uint32_t sb16 = (rbase + (uint32_t)sl * slot_bytes) >> 4;
#pragma unroll 1
            for (int rr = 0; rr < 4; ++rr) {
              const uint32_t tb16 = sb16 + (((uint32_t)(rr * planes) * SUB_STRIDE) >> 4);
#pragma unroll
              for (int my = 0; my < 2; ++my) {
                if (a.cin8) {
                  umma_bf16(d_tmem, ((uint64_t)a_hi << 32) | ((tb16 + (uint32_t)(my * SXV)) | a_lbo), ((uint64_t)b_hi << 32) | blo, idesc, acc);
                  acc = 1; blo += (uint32_t)(NPAD * 32) >> 4;
                } else {
#pragma unroll
                  for (int mx = 0; mx < 2; ++mx) {
                    uint32_t alo = (tb16 + (uint32_t)(my * SXV + mx)) | a_lbo;
                    for (int kc = 0; kc < kcs; ++kc) {
                      umma_bf16(d_tmem, ((uint64_t)a_hi << 32) | alo, ((uint64_t)b_hi << 32) | blo, idesc, acc);
                      acc = 1; alo += (uint32_t)(2 * SUB_STRIDE) >> 4; blo += (uint32_t)(NPAD * 32) >> 4;
                    }
                  }
                }
              }
            }
          }
          umma_commit(&tfull_bar[zo & 1]);
          umma_commit(&empty_bar[zslot]);                          // input slices 2zo and 2zo+1 are done
          umma_commit(&empty_bar[s1]);
        }
        __syncwarp();
        zslot += 2; if (zslot >= RING) zslot -= RING;
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int yl = row >> 3, xl = row & 7;
    const int oy = y0 + yl, ox = x0 + xl;
    const bool inside = oy < a.L[1] && ox < a.L[2] && !(a.dbg & 1);
    for (int zo = 0; zo < nz; ++zo) {
      const int oz = z0 + zo;
      uint4 refq[NPAD / 8], accq[NPAD / 8];
      if (inside) {
#pragma unroll
        for (int c = 0; c < NPAD / 8; ++c) {
          if (c * 8 < a.Cout) {
            if (MODE >= 2 && a.ref) refq[c] = __ldg(reinterpret_cast<const uint4*>(a.ref + Epi::ref_off(a, b, oz, oy, ox) + c * 8));
            if (MODE >= 2 && a.accumulate) accq[c] = *reinterpret_cast<const uint4*>(a.out + Epi::out_off(a, b, oz, oy, ox) + c * 8);
          }
        }
      }
      mbar_wait(&tfull_bar[zo & 1], ((uint32_t)(zo >> 1)) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r[NPAD];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(zo & 1) * NPAD;
#pragma unroll
      for (int c = 0; c < NPAD; c += 8) tmem_ld8(taddr + c, r + c);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&tempty_bar[zo & 1]);
      if (!inside) continue;
#pragma unroll
      for (int c = 0; c < NPAD / 8; ++c) {
        if (c * 8 < a.Cout) {
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = __uint_as_float(r[c * 8 + u]);
          Epi::run<MODE>(a, v, c * 8, b, oz, oy, ox, refq[c], accq[c]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// DOWN, input-slice-major (second half of round 2).  conv_down_tc_kernel issues, per OUTPUT slice, one MMA per (kz, ry/rx class,
// m'y[, m'x, channel pair]) with N = Cout padded to 16: 32 small MMAs per slice at Cin = 8, half of N padding at Cout = 8, and
// it was MMA-issue bound (profiles/ablation_s2_r2.txt).  An input slice s feeds exactly two output slices -- zo = s/2 - 1
// through kz = (s & 1) + 2 and zo = s/2 through kz = s & 1 -- so here the MMAs are issued per INPUT slice with
// N = [W(kz = p + 2) | W(kz = p)] (p = s & 1) into two ADJACENT column groups of a TMEM-resident strip (group g = zo + 1):
// half the MMAs, no padding, every ring slot is consumed by exactly one MMA group.  The strip is zeroed once (one work item
// per CTA); output slice zo is complete when input slice 2 zo + 3 has been multiplied.
// ------------------------------------------------------------------------------------------------
constexpr int kDown2MaxZ = 32;
template <int CO, int MODE>
__global__ void __launch_bounds__(kThreads)
conv_down2_tc_kernel(const __grid_constant__ CUtensorMap map0, const S2Args a) {
  constexpr int N2 = 2 * CO;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[RING_MAX], empty_bar[RING_MAX], w_bar, strip_bar, tfull_bar[kDown2MaxZ];
  __shared__ uint32_t tmem_base_s;
  const int RING = a.ring;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int planes = a.planes;
  const uint32_t wbytes_pad = (uint32_t)((a.wbytes + 1023) & ~1023);
  uint8_t* wsm = smem;
  uint8_t* ring = smem + wbytes_pad;
  const uint32_t slot_bytes = (uint32_t)(4 * planes) * SUB_STRIDE;
  const uint32_t tmem_cols = (uint32_t)a.np;          // strip: (zc + 2) groups of CO columns, rounded up to a power of two

  int b, x0, y0, z0, nz; decode_work(a, b, x0, y0, z0, nz);
  const int nslices = 2 * nz + 2;                  // input slices 2*z0 - pad .. 2*(z0+nz-1) + 3 - pad

  if (threadIdx.x == 0) {
    for (int i = 0; i < RING; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&w_bar, 1); mbar_init(&strip_bar, 4);
    for (int i = 0; i < nz; ++i) mbar_init(&tfull_bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&w_bar, (uint32_t)a.wbytes);
      bulk_load(wsm, a.wpacked, (uint32_t)a.wbytes, &w_bar);
      int slot = 0; uint32_t ph = 0;
      for (int s = 0; s < nslices; ++s) {
        mbar_wait(&empty_bar[slot], ph ^ 1u);
        if (a.dbg & 4) { mbar_arrive(&full_bar[slot]); if (++slot == RING) { slot = 0; ph ^= 1u; } continue; }
        mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)(4 * planes) * SUB_BYTES);
        uint8_t* dst = ring + (size_t)slot * slot_bytes;
        const int zin = 2 * z0 - a.pad + s + a.shift[0];
        for (int rr = 0; rr < 4; ++rr) {
          const int ry = rr >> 1, rx = rr & 1;
          // parity r of (k - pad): k = kb, kb + 2 with kb = (r + pad) & 1; first tap offset m_lo = (kb - pad - r) / 2
          const int mly = (((ry + a.pad) & 1) - a.pad - ry) / 2, mlx = (((rx + a.pad) & 1) - a.pad - rx) / 2;
          const int cy = 2 * (y0 + mly) + ry + a.shift[1], cx = 2 * (x0 + mlx) + rx + a.shift[2];
          for (int p = 0; p < planes; ++p)
            tma_load_5d(dst + (rr * planes + p) * SUB_STRIDE, &map0, &full_bar[slot], p * 8, cx, cy, zin, b);
        }
        if (++slot == RING) { slot = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N2 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    mbar_wait(&w_bar, 0);
    const uint32_t wb16 = smem_u32(wsm) >> 4;
    const uint32_t rbase = smem_u32(ring);
    const uint32_t a_hi = ((uint32_t)ROW_B >> 4) | (1u << 14), b_hi = (128u >> 4) | (1u << 14);
    const uint32_t b_lbo = ((uint32_t)(N2 * 16) >> 4) << 16;
    const uint32_t a_lbo = a.cin8 ? (1u << 16) : (((uint32_t)SUB_STRIDE >> 4) << 16);
    const int kcs = planes >> 1;
    const uint32_t par16 = (uint32_t)(a.wbytes >> 1) >> 4;       // the image of input-slice parity 1 follows that of parity 0
    mbar_wait(&strip_bar, 0);                                     // strip zeroed by the epilogue warps
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    int slot = 0; uint32_t ph = 0;
    for (int s = 0; s < nslices; ++s) {
      mbar_wait(&full_bar[slot], ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t d_tmem = tmem_base + (uint32_t)((s >> 1) * CO);
        const uint32_t sb16 = (rbase + (uint32_t)slot * slot_bytes) >> 4;
        uint32_t blo = (wb16 + (uint32_t)(s & 1) * par16) | b_lbo;
        if (!(a.dbg & 16)) {
#pragma unroll 1
          for (int rr = 0; rr < 4; ++rr) {
            const uint32_t tb16 = sb16 + (((uint32_t)(rr * planes) * SUB_STRIDE) >> 4);
#pragma unroll
            for (int my = 0; my < 2; ++my) {
              if (a.cin8) {
                umma_bf16(d_tmem, ((uint64_t)a_hi << 32) | ((tb16 + (uint32_t)(my * SXV)) | a_lbo), ((uint64_t)b_hi << 32) | blo, idesc, 1u);
                blo += (uint32_t)(N2 * 32) >> 4;
              } else {
#pragma unroll
                for (int mx = 0; mx < 2; ++mx) {
                  uint32_t alo = (tb16 + (uint32_t)(my * SXV + mx)) | a_lbo;
                  for (int kc = 0; kc < kcs; ++kc) {
                    umma_bf16(d_tmem, ((uint64_t)a_hi << 32) | alo, ((uint64_t)b_hi << 32) | blo, idesc, 1u);
                    alo += (uint32_t)(2 * SUB_STRIDE) >> 4; blo += (uint32_t)(N2 * 32) >> 4;
                  }
                }
              }
            }
          }
        }
        umma_commit(&empty_bar[slot]);                           // this input slice is used by this MMA group only
        if ((s & 1) && s >= 3) umma_commit(&tfull_bar[(s - 3) >> 1]);   // output slice (s - 3) / 2 has its four kz parts
      }
      __syncwarp();
      if (++slot == RING) { slot = 0; ph ^= 1u; }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int yl = row >> 3, xl = row & 7;
    const int oy = y0 + yl, ox = x0 + xl;
    const bool inside = oy < a.L[1] && ox < a.L[2] && !(a.dbg & 1);
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    // zero the strip: groups 0 .. nz + 1 (group 0 and nz + 1 only collect the parts of the output slices outside this chunk)
    for (int c = 0; c < (nz + 2) * CO; c += 8)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(lane_base + (uint32_t)c), "r"(0u) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (lane == 0) mbar_arrive(&strip_bar);
    for (int zo = 0; zo < nz; ++zo) {
      const int oz = z0 + zo;
      uint4 refq[CO / 8], accq[CO / 8];
      if (inside) {
#pragma unroll
        for (int c = 0; c < CO / 8; ++c) {
          if (c * 8 < a.Cout) {
            if (MODE >= 2 && a.ref) refq[c] = __ldg(reinterpret_cast<const uint4*>(a.ref + Epi::ref_off(a, b, oz, oy, ox) + c * 8));
            if (MODE >= 2 && a.accumulate) accq[c] = *reinterpret_cast<const uint4*>(a.out + Epi::out_off(a, b, oz, oy, ox) + c * 8);
          }
        }
      }
      mbar_wait(&tfull_bar[zo], 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r[CO];
      const uint32_t taddr = lane_base + (uint32_t)((zo + 1) * CO);
#pragma unroll
      for (int c = 0; c < CO; c += 8) tmem_ld8(taddr + c, r + c);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (!inside) continue;
#pragma unroll
      for (int c = 0; c < CO / 8; ++c) {
        if (c * 8 < a.Cout) {
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = __uint_as_float(r[c * 8 + u]);
          Epi::run<MODE>(a, v, c * 8, b, oz, oy, ox, refq[c], accq[c]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// UP, wide layers (Cin multiple of 64, any Cout: wf <= 2).  Same GEMM as conv_up_tc_kernel with CP = 32 (N = 8 classes x
// 32 output channels = 256 columns, blockIdx.y = group of 32 output channels), but K = 8 taps x Cin no longer fits in
// shared memory as weights (Cin = 64 already needs 256 KB), so BOTH operands are streamed:
//   * inputs: per q-slice and 64-channel chunk, the two source slices (qz-1, qz) of that chunk (40 KB stage, 2 stages);
//   * weights: two K-steps (2 x 16 channels x 256 columns = 16 KB) per stage through a 6-stage ring, in MMA order.
// Two producer warps (inputs / weights), one MMA issuer, eight epilogue warps.
// ------------------------------------------------------------------------------------------------
constexpr int UW_PL = 8;                                 // planes (64 channels) per input chunk
constexpr int UW_IN_STAGE = 2 * UW_PL * SUB_STRIDE;      // 40960
constexpr int UW_NIN = 2;
constexpr int UW_KSTEP = 256 * 32;                       // 8192 B: one K-step [k-half][32 n-groups][8][8] bf16
constexpr int UW_WRING = 96 * 1024;                      // weight ring: 6 slots of two K-steps (plane layout) / 3 slots of four (swizzled rows)
constexpr int UW_NW = 6;
constexpr int kThreadsUpw = 352;                         // warps: 0 input TMA, 1 MMA, 2-9 epilogue, 10 weight loads

__global__ void __launch_bounds__(kThreadsUpw, 1)
conv_upw_tc_kernel(const __grid_constant__ CUtensorMap map0, const S2Args a) {
  constexpr int CP = 32, NP = 256;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t in_full[UW_NIN], in_empty[UW_NIN], w_full[UW_NW], w_empty[UW_NW], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* inring = smem;
  uint8_t* wring = smem + UW_NIN * UW_IN_STAGE;
  const int cg = blockIdx.y, co0 = cg * CP;
  const int nchunks = a.planes / UW_PL;
  const int kps = a.swz ? 4 : 2;                         // K-steps per weight stage
  const int wstage = kps * UW_KSTEP, nw = UW_WRING / wstage;

  if (threadIdx.x == 0) {
    for (int i = 0; i < UW_NIN; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_empty[i], 1); }
    for (int i = 0; i < UW_NW; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  int b, x0, y0, z0, nz; decode_work(a, b, x0, y0, z0, nz);

  if (warp == 0) {
    int slot = 0; uint32_t ph = 0;
    for (int zo = 0; zo < nz; ++zo)
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&in_empty[slot], ph ^ 1u);
        if (elect_one()) {
          if (a.dbg & 4) { mbar_arrive(&in_full[slot]); } else {
          mbar_arrive_expect_tx(&in_full[slot], (uint32_t)(2 * UW_PL) * SUB_BYTES);
          uint8_t* dst = inring + (size_t)slot * UW_IN_STAGE;
          for (int sl = 0; sl < 2; ++sl) {
            if (a.swz) tma_load_5d(dst + sl * (UW_PL * SUB_STRIDE), &map0, &in_full[slot], c * 64, x0 - 1 + a.shift[2], y0 - 1 + a.shift[1], z0 + zo - 1 + sl + a.shift[0], b);
            else
              for (int p = 0; p < UW_PL; ++p)
                tma_load_5d(dst + (sl * UW_PL + p) * SUB_STRIDE, &map0, &in_full[slot], (c * UW_PL + p) * 8, x0 - 1 + a.shift[2], y0 - 1 + a.shift[1], z0 + zo - 1 + sl + a.shift[0], b);
          }
          }
        }
        __syncwarp();
        if (++slot == UW_NIN) { slot = 0; ph ^= 1u; }
      }
  } else if (warp == 10) {
    int slot = 0; uint32_t ph = 0;
    const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.wpacked) + (size_t)cg * nchunks * 32 * UW_KSTEP;
    for (int zo = 0; zo < nz; ++zo)
      for (int st = 0; st < nchunks * (32 / kps); ++st) {   // stage order = MMA order: chunk, tap (mz,my,mx), group of 16-channel steps
        mbar_wait(&w_empty[slot], ph ^ 1u);
        if (elect_one()) {
          if (a.dbg & 2) mbar_arrive(&w_full[slot]);
          else {
            mbar_arrive_expect_tx(&w_full[slot], (uint32_t)wstage);
            bulk_load(wring + (size_t)slot * wstage, wsrc + (size_t)st * wstage, (uint32_t)wstage, &w_full[slot]);
          }
        }
        __syncwarp();
        if (++slot == nw) { slot = 0; ph ^= 1u; }
      }
  } else if (warp == 1) {
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // A operand.  Plane layout: 8-channel planes, SBO = one tile row, LBO = next plane (K halves).  Swizzled layout: rows of
    // 64 channels (128 B, SWIZZLE_128B as written by TMA), SBO = one tile row of 9 voxels, a K-step is 32 B further in the row;
    // a tap is a start address shifted by whole rows (the swizzle XOR is a function of the shared-memory address).
    const uint32_t a_hi = a.swz ? (((uint32_t)(SXV * 128) >> 4) | (1u << 14) | (2u << 29)) : (((uint32_t)ROW_B >> 4) | (1u << 14));
    // B operand.  Plane layout: [k-half][n-group][8][8] per K-step.  Swizzled: 256 rows (n) of 64 input channels (128 B,
    // chunk index XOR row & 7 done by the pack kernel), SBO = 1024 B, a K-step is 32 B further in the row.
    const uint32_t b_hi = a.swz ? ((1024u >> 4) | (1u << 14) | (2u << 29)) : ((128u >> 4) | (1u << 14));
    const uint32_t a_lbo = a.swz ? (1u << 16) : (((uint32_t)SUB_STRIDE >> 4) << 16);
    const uint32_t b_lbo = a.swz ? (1u << 16) : (((uint32_t)(NP * 16) >> 4) << 16);             // K halves of a weight stage
    const uint32_t in16 = smem_u32(inring) >> 4, w16 = smem_u32(wring) >> 4;
    int islot = 0; uint32_t iph = 0; int wslot = 0; uint32_t wph = 0;
    for (int zo = 0; zo < nz; ++zo) {
      mbar_wait(&tempty_bar[zo & 1], (((uint32_t)(zo >> 1)) & 1u) ^ 1u);
      const uint32_t d_tmem = tmem_base + (uint32_t)(zo & 1) * NP;
      uint32_t acc = 0u;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&in_full[islot], iph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ib16 = in16 + (uint32_t)islot * (UW_IN_STAGE >> 4);
#pragma unroll 1
        for (int tap = 0; tap < 8; ++tap) {
          const int mz = tap >> 2, my = (tap >> 1) & 1, mx = tap & 1;
#pragma unroll 1
          for (int kp = 0; kp < 4 / kps; ++kp) {
            mbar_wait(&w_full[wslot], wph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
#pragma unroll 2
              for (int k2 = 0; k2 < kps; ++k2) {
                const int kc = kps * kp + k2;
                const uint32_t aoff = a.swz ? (uint32_t)((mz * UW_PL * SUB_STRIDE) >> 4) + (uint32_t)((my * SXV + mx) * 8 + kc * 2)
                                            : (uint32_t)(((mz * UW_PL + 2 * kc) * SUB_STRIDE) >> 4) + (uint32_t)(my * SXV + mx);
                const uint32_t alo = (ib16 + aoff) | a_lbo;
                const uint32_t blo = (w16 + (uint32_t)wslot * ((uint32_t)wstage >> 4) + (uint32_t)(a.swz ? k2 * 2 : k2 * (UW_KSTEP >> 4))) | b_lbo;
                umma_bf16(d_tmem, ((uint64_t)a_hi << 32) | alo, ((uint64_t)b_hi << 32) | blo, idesc, (k2 == 0) ? acc : 1u);
              }
              umma_commit(&w_empty[wslot]);
            }
            __syncwarp();
            acc = 1u;
            if (++wslot == nw) { wslot = 0; wph ^= 1u; }
          }
        }
        if (elect_one()) {
          umma_commit(&in_empty[islot]);
          if (c == nchunks - 1) umma_commit(&tfull_bar[zo & 1]);
        }
        __syncwarp();
        if (++islot == UW_NIN) { islot = 0; iph ^= 1u; }
      }
    }
  } else {
    const int q = warp & 3;                        // TMEM lane quadrant of this warp
    const int rz = (warp - 2) >> 2;                // output z parity handled by this warp
    const int row = q * 32 + lane;
    const int yl = row >> 3, xl = row & 7;
    constexpr int NCH = CP / 8;
    // The epilogue is instruction-latency bound (8 warps, a few hundred dependent instructions per class block), so all
    // index arithmetic is hoisted: a thread owns ONE q-voxel, its four (ry,rx) classes are fixed element offsets from the
    // (0,0) class, z advances by a constant stride, validity of a class is a per-thread constant.  The long-latency operands
    // (LeakyReLU' reference, accumulate target) of the next q-slice are pulled into L2 while this one is processed and the
    // registers of class block c4 + 1 are loaded before block c4 is converted.
    const int oy0 = 2 * (y0 + yl) - a.pad, ox0 = 2 * (x0 + xl) - a.pad;
    const int oz0 = 2 * z0 + rz - a.pad;
    bool okc[4]; int ooff[4], roff[4];
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
      const int oy = oy0 + (c4 >> 1), ox = ox0 + (c4 & 1);
      okc[c4] = !(a.dbg & 1) && oy >= 0 && oy < a.L[1] && ox >= 0 && ox < a.L[2] && co0 < a.Cout;
      ooff[c4] = ((c4 >> 1) * a.OX + (c4 & 1)) * a.out_C;
      roff[c4] = ((c4 >> 1) * a.RX + (c4 & 1)) * a.ref_C;
    }
    const int ncol = min(CP, a.Cout - co0);
    const long long o_zs = 2LL * a.OY * a.OX * a.out_C, r_zs = 2LL * a.RY * a.RX * a.ref_C;      // two output slices per q-slice
    bf16* const out0 = a.out + Epi::out_off(a, b, oz0, oy0, ox0) + co0;                        // class (0,0) voxel of q-slice 0
    const bf16* const ref0 = a.ref ? a.ref + Epi::ref_off(a, b, oz0, oy0, ox0) + co0 : nullptr;
    const uint32_t di0 = (uint32_t)(((((long long)b * a.L[0] + oz0) * a.L[1] + oy0) * a.L[2] + ox0) * a.Cout) + (uint32_t)co0;
    const uint32_t di_zs = (uint32_t)(2 * a.L[1] * a.L[2] * a.Cout);
    auto zok = [&](int zo) { const int oz = oz0 + 2 * zo; return zo < nz && oz >= 0 && oz < a.L[0]; };
    auto prefetch_l2 = [&](int zo) {
      if (!zok(zo)) return;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        if (okc[c4]) {
          if (a.ref) asm volatile("prefetch.global.L2 [%0];" ::"l"(ref0 + zo * r_zs + roff[c4]));
          if (a.accumulate) asm volatile("prefetch.global.L2 [%0];" ::"l"(out0 + zo * o_zs + ooff[c4]));
        }
      }
    };
    auto fetch = [&](int zo, int c4, uint4* rq, uint4* aq) {
      if (zok(zo) && okc[c4]) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          if (c * 8 < ncol) {
            if (a.ref) rq[c] = __ldg(reinterpret_cast<const uint4*>(ref0 + zo * r_zs + roff[c4] + c * 8));
            if (a.accumulate) aq[c] = *reinterpret_cast<const uint4*>(out0 + zo * o_zs + ooff[c4] + c * 8);
          }
        }
      }
    };
    prefetch_l2(0);
    uint4 refq[2][NCH], accq[2][NCH];
    for (int zo = 0; zo < nz; ++zo) {
      prefetch_l2(zo + 1);
      fetch(zo, 0, refq[0], accq[0]);
      const bool zv = zok(zo);
      bf16* const outz = out0 + zo * o_zs;
      mbar_wait(&tfull_bar[zo & 1], ((uint32_t)(zo >> 1)) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(zo & 1) * NP + (uint32_t)(rz * 4 * CP);
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        __syncwarp();
        if (c4 < 3) fetch(zo, c4 + 1, refq[(c4 + 1) & 1], accq[(c4 + 1) & 1]);
        uint32_t r[CP];
#pragma unroll
        for (int c = 0; c < CP; c += 8) tmem_ld8(taddr + (uint32_t)(c4 * CP + c), r + c);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c4 == 3) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          mbar_arrive(&tempty_bar[zo & 1]);
        }
        if (zv && okc[c4]) {
          bf16* const op = outz + ooff[c4];
          const uint32_t di = di0 + (uint32_t)zo * di_zs + (uint32_t)(((c4 >> 1) * a.L[2] + (c4 & 1)) * a.Cout);
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            if (c * 8 < ncol) {
              float v[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) v[u] = __uint_as_float(r[c * 8 + u]);
              if (a.ref) {
                const uint4 rq = refq[c4 & 1][c];
                const uint32_t w[4] = {rq.x, rq.y, rq.z, rq.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {      // LeakyReLU'(ref): ref > 0 <=> sign bit clear and magnitude non-zero, on the raw bf16 halves
                  if (!((w[i] & 0x7fffu) != 0u && (w[i] & 0x8000u) == 0u)) v[2 * i] *= a.ref_slope;
                  if (!((w[i] & 0x7fff0000u) != 0u && (w[i] & 0x80000000u) == 0u)) v[2 * i + 1] *= a.ref_slope;
                }
              }
              if (a.drop_key) {
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] *= 2.f * tem_keep(a.drop_key, di + (uint32_t)(c * 8 + u));
              }
              if (a.accumulate) {
                float o[8]; unpack8(accq[c4 & 1][c], o);
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] += o[u];
              }
              if (a.slope != 1.f) {
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = v[u] > 0.f ? v[u] : v[u] * a.slope;
              }
              uint4 pk;
              pk.x = pack2(v[0], v[1]); pk.y = pack2(v[2], v[3]); pk.z = pack2(v[4], v[5]); pk.w = pack2(v[6], v[7]);
              *reinterpret_cast<uint4*>(op + c * 8) = pk;
            }
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// weight stages of the wide UP kernel: [cout group][chunk][tap = (mz,my,mx)][kc][k-half][n-group (32)][8][8], n = class*32 + co
// swizzled form: [cout group][chunk][tap][n (256 rows)][64 input channels], 16 B chunk index XOR (n & 7) (SWIZZLE_128B rows)
struct PackUpwArgs { const float* w; long long ws_tap, ws_in, ws_out; int nchunks, cout, swz; bf16* dst; long long total; };
__global__ void pack_weights_upw_kernel(const PackUpwArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.total) return;
  long long t = i;
  int e, n, kc, j;
  if (a.swz) {
    e = (int)(t & 7); t >>= 3;
    const int ch = (int)(t & 7); t >>= 3;
    n = (int)(t & 255); t >>= 8;
    const int lc = ch ^ (n & 7);                         // logical 8-channel chunk stored at physical chunk ch
    kc = lc >> 1; j = lc & 1;
  } else {
    e = (int)(t & 7); t >>= 3;
    const int r = (int)(t & 7); t >>= 3;
    const int g = (int)(t & 31); t >>= 5;
    j = (int)(t & 1); t >>= 1;
    kc = (int)(t & 3); t >>= 2;
    n = g * 8 + r;
  }
  const int tap = (int)(t & 7); t >>= 3;
  const int c = (int)(t % a.nchunks); t /= a.nchunks;
  const int cg = (int)t;
  const int cls = n >> 5, co = cg * 32 + (n & 31);
  const int rz = cls >> 2, ry = (cls >> 1) & 1, rx = cls & 1;
  const int mz = tap >> 2, my = (tap >> 1) & 1, mx = tap & 1;
  const int kz = rz + 2 * (1 - mz), ky = ry + 2 * (1 - my), kx = rx + 2 * (1 - mx);
  const int ci = c * 64 + (2 * kc + j) * 8 + e;
  float v = 0.f;
  if (co < a.cout) v = a.w[(long long)((kz * 4 + ky) * 4 + kx) * a.ws_tap + (long long)ci * a.ws_in + (long long)co * a.ws_out];
  a.dst[i] = __float2bfloat16_rn(v);
}

// ------------------------------------------------------------------------------------------------
// DOWN, wide layers (Cin multiple of 16, Cout up to 256 per CTA: wf <= 4).  N = NP = Cout rounded to 64 columns; the
// K = 64 taps x Cin reduction is split into 16-channel chunks (one K-step = two planes) and the CTA owns a TMEM strip of
// ZS = 512 / NP output slices that stays resident over all chunks (the conv_tcw.cu recipe), so that every input slice
// pair (2j, 2j+1) of a chunk is staged ONCE and feeds both output slices that read it (zo = j with kz = 0,1 and
// zo = j - 1 with kz = 2,3).  Weights are streamed in MMA order, one K-step (16 x NP) per ring slot.
//   input stage : [slice 2][parity (ry,rx) 4][plane 2] de-interleaved sub-tiles = 40 KB, 3 stages
//   weight stage: KS = 8 / 4 / 2 K-steps of NP x 32 B (NP = 64 / 128 / 256), six 16 KB slots
// Warps: 0 input TMA, 1 MMA issuer, 2-9 epilogue (two per TMEM lane quadrant, half of the columns each), 10 weights.
// ------------------------------------------------------------------------------------------------
constexpr int DW_STAGE = 16 * SUB_STRIDE;                // 40960
constexpr int DW_NIN = 3;
constexpr int DW_WSLOT = 16 * 1024;                      // one ring slot = KS K-steps (KS x NP x 32 B <= 16 KB): 256 tensor cycles per barrier round trip
constexpr int DW_WMAX = 6;
constexpr int kThreadsDw = 352;
constexpr int DW_ZMAX = 8;                               // output slices per CTA at NP = 64

__global__ void __launch_bounds__(kThreadsDw, 1)
conv_downw_tc_kernel(const __grid_constant__ CUtensorMap map0, const S2Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t in_full[DW_NIN], in_empty[DW_NIN], w_full[DW_WMAX], w_empty[DW_WMAX], tfull_bar[DW_ZMAX];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* inring = smem;
  uint8_t* wring = smem + DW_NIN * DW_STAGE;
  const int NP = a.np;
  const int KS = a.ring;                                 // K-steps per weight stage
  const int wstage = KS * NP * 32;
  const int nst = 32 / KS;                               // stages per (chunk, kz pair)
  const int cg = blockIdx.y, co0 = cg * NP;
  const int nch = a.planes >> 1;                         // 16-channel chunks

  int b, x0, y0, z0, nz; decode_work(a, b, x0, y0, z0, nz);

  if (threadIdx.x == 0) {
    for (int i = 0; i < DW_NIN; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_empty[i], 1); }
    for (int i = 0; i < DW_WMAX; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < DW_ZMAX; ++i) mbar_init(&tfull_bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    int slot = 0; uint32_t ph = 0;
    for (int c = 0; c < nch; ++c)
      for (int j = 0; j <= nz; ++j) {
        mbar_wait(&in_empty[slot], ph ^ 1u);
        if (elect_one()) {
          if (a.dbg & 4) mbar_arrive(&in_full[slot]);
          else {
          mbar_arrive_expect_tx(&in_full[slot], 16u * SUB_BYTES);
          uint8_t* dst = inring + (size_t)slot * DW_STAGE;
          for (int sl = 0; sl < 2; ++sl) {
            const int zin = 2 * (z0 + j) + sl - a.pad + a.shift[0];
            for (int rr = 0; rr < 4; ++rr) {
              const int ry = rr >> 1, rx = rr & 1;
              const int mly = (((ry + a.pad) & 1) - a.pad - ry) / 2, mlx = (((rx + a.pad) & 1) - a.pad - rx) / 2;
              const int cy = 2 * (y0 + mly) + ry + a.shift[1], cx = 2 * (x0 + mlx) + rx + a.shift[2];
              for (int p = 0; p < 2; ++p)
                tma_load_5d(dst + ((sl * 4 + rr) * 2 + p) * SUB_STRIDE, &map0, &in_full[slot], (2 * c + p) * 8, cx, cy, zin, b);
            }
          }
          }
        }
        __syncwarp();
        if (++slot == DW_NIN) { slot = 0; ph ^= 1u; }
      }
  } else if (warp == 10) {
    int slot = 0; uint32_t ph = 0;
    const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.wpacked) + (size_t)cg * nch * 64 * (NP * 32);
    for (int c = 0; c < nch; ++c)
      for (int j = 0; j <= nz; ++j)
        for (int h = 0; h < 2; ++h) {
          const int zo = j - h;
          if (zo < 0 || zo >= nz) continue;
          const uint8_t* src = wsrc + (size_t)(c * 64 + h * 32) * (NP * 32);    // K-step order = MMA order: kz, parity, my, mx
          for (int i = 0; i < nst; ++i) {
            mbar_wait(&w_empty[slot], ph ^ 1u);
            if (elect_one()) {
              if (a.dbg & 2) mbar_arrive(&w_full[slot]);
              else {
                mbar_arrive_expect_tx(&w_full[slot], (uint32_t)wstage);
                bulk_load(wring + (size_t)slot * DW_WSLOT, src + (size_t)i * wstage, (uint32_t)wstage, &w_full[slot]);
              }
            }
            __syncwarp();
            if (++slot == DW_WMAX) { slot = 0; ph ^= 1u; }
          }
        }
  } else if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_hi = ((uint32_t)ROW_B >> 4) | (1u << 14), b_hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lbo = ((uint32_t)SUB_STRIDE >> 4) << 16;            // K halves = the two planes of the chunk
    const uint32_t b_lbo = ((uint32_t)(NP * 16) >> 4) << 16;
    const uint32_t in16 = smem_u32(inring) >> 4, w16 = smem_u32(wring) >> 4;
    const uint32_t kst16 = (uint32_t)(NP * 32) >> 4;
    int islot = 0; uint32_t iph = 0; int wslot = 0; uint32_t wph = 0;
    for (int c = 0; c < nch; ++c)
      for (int j = 0; j <= nz; ++j) {
        mbar_wait(&in_full[islot], iph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ib16 = in16 + (uint32_t)islot * (DW_STAGE >> 4);
        for (int h = 0; h < 2; ++h) {
          const int zo = j - h;
          if (zo < 0 || zo >= nz) continue;
          const uint32_t d_tmem = tmem_base + (uint32_t)(zo * NP);
          uint32_t acc = (c == 0 && h == 0) ? 0u : 1u;
#pragma unroll 1
          for (int st = 0; st < nst; ++st) {
            mbar_wait(&w_full[wslot], wph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
              const uint32_t wb16 = w16 + (uint32_t)wslot * (DW_WSLOT >> 4);
#pragma unroll 2
              for (int k = 0; k < KS; ++k) {
                const int i = st * KS + k;                // i = (sl, parity, my, mx)
                const uint32_t alo = (ib16 + (uint32_t)((((i >> 2) * 2) * SUB_STRIDE) >> 4) + (uint32_t)(((i >> 1) & 1) * SXV + (i & 1))) | a_lbo;
                const uint32_t blo = (wb16 + (uint32_t)k * kst16) | b_lbo;
                umma_bf16(d_tmem, ((uint64_t)a_hi << 32) | alo, ((uint64_t)b_hi << 32) | blo, idesc, (k == 0) ? acc : 1u);
              }
              umma_commit(&w_empty[wslot]);
            }
            __syncwarp();
            acc = 1u;
            if (++wslot == DW_WMAX) { wslot = 0; wph ^= 1u; }
          }
          if (c == nch - 1 && h == 1) {                   // output slice zo has received all four kz taps of the last chunk
            if (elect_one()) umma_commit(&tfull_bar[zo]);
            __syncwarp();
          }
        }
        if (elect_one()) umma_commit(&in_empty[islot]);
        __syncwarp();
        if (++islot == DW_NIN) { islot = 0; iph ^= 1u; }
      }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;                    // column half handled by this warp
    const int row = q * 32 + lane;
    const int yl = row >> 3, xl = row & 7;
    const int oy = y0 + yl, ox = x0 + xl;
    const bool inside = oy < a.L[1] && ox < a.L[2] && !(a.dbg & 1);
    const int cbeg = half * (NP >> 1), cend = cbeg + (NP >> 1);
    if (inside && (a.ref || a.accumulate)) {            // epilogue operands of the whole strip into L2 while the MMAs run
      for (int zo = 0; zo < nz; ++zo)
        for (int cb = cbeg; cb < cend; cb += 32)
          if (co0 + cb < a.Cout) {
            if (a.ref) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.ref + Epi::ref_off(a, b, z0 + zo, oy, ox) + co0 + cb));
            if (a.accumulate) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.out + Epi::out_off(a, b, z0 + zo, oy, ox) + co0 + cb));
          }
    }
    for (int zo = 0; zo < nz; ++zo) {
      const int oz = z0 + zo;
      mbar_wait(&tfull_bar[zo], 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(zo * NP);
      for (int cb = cbeg; cb < cend; cb += 32) {
        __syncwarp();
        uint4 refq[4], accq[4];
        if (inside) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (co0 + cb + c * 8 < a.Cout) {
              if (a.ref) refq[c] = __ldg(reinterpret_cast<const uint4*>(a.ref + Epi::ref_off(a, b, oz, oy, ox) + co0 + cb + c * 8));
              if (a.accumulate) accq[c] = *reinterpret_cast<const uint4*>(a.out + Epi::out_off(a, b, oz, oy, ox) + co0 + cb + c * 8);
            }
          }
        }
        uint32_t r[32];
#pragma unroll
        for (int c = 0; c < 32; c += 8) tmem_ld8(taddr + (uint32_t)(cb + c), r + c);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (inside) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (co0 + cb + c * 8 < a.Cout) {
              float v[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) v[u] = __uint_as_float(r[c * 8 + u]);
              Epi::run<3>(a, v, co0 + cb + c * 8, b, oz, oy, ox, refq[c], accq[c]);
            }
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// weight stages of the wide DOWN kernel: [cout group][chunk][kz][parity (ry,rx)][my][mx][k-half][n-group (NP/8)][8][8]
struct PackDwArgs { const float* w; long long ws_tap, ws_in, ws_out; int nch, cout, np, pad; bf16* dst; long long total; };
__global__ void pack_weights_dw_kernel(const PackDwArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.total) return;
  long long t = i;
  const int e = (int)(t & 7); t >>= 3;
  const int r = (int)(t & 7); t >>= 3;
  const int ng = a.np >> 3;
  const int g = (int)(t % ng); t /= ng;
  const int j = (int)(t & 1); t >>= 1;
  const int mx = (int)(t & 1); t >>= 1;
  const int my = (int)(t & 1); t >>= 1;
  const int rr = (int)(t & 3); t >>= 2;
  const int kz = (int)(t & 3); t >>= 2;
  const int c = (int)(t % a.nch); t /= a.nch;
  const int cg = (int)t;
  const int ry = rr >> 1, rx = rr & 1;
  const int ky = ((ry + a.pad) & 1) + 2 * my, kx = ((rx + a.pad) & 1) + 2 * mx;
  const int co = cg * a.np + g * 8 + r, ci = c * 16 + j * 8 + e;
  float v = 0.f;
  if (co < a.cout) v = a.w[(long long)((kz * 4 + ky) * 4 + kx) * a.ws_tap + (long long)ci * a.ws_in + (long long)co * a.ws_out];
  a.dst[i] = __float2bfloat16_rn(v);
}

int wide_swizzle() {       // debug knob: 8-channel plane tiles (SWIZZLE_NONE) in the wide UP kernel
  static const int v = getenv("TEM_S2_NO_SWIZZLE") ? 0 : 1;
  return v;
}
int dw_np(int cout) { const int n = (cout + 63) / 64 * 64; return n > 256 ? 256 : n; }

// bf16 UMMA B image [step][k-half][n-group][8 rows][8 elems]; the step order is the issue order of the kernels above
struct PackS2Args {
  const float* w; long long ws_tap, ws_in, ws_out;
  int up, pad, cin, cols, cp, np, cin8;
  bf16* dst; int total;
};
__global__ void pack_weights_s2_kernel(const PackS2Args a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.total) return;
  int t = i;
  const int e = t & 7; t >>= 3;
  const int r = t & 7; t >>= 3;
  const int ng = a.np >> 3;
  const int g = t % ng; t /= ng;
  const int j = t & 1; t >>= 1;
  int s = t;
  const int n = g * 8 + r;
  const int kcs = a.cin >> 4;
  int mx, ci;
  if (a.cin8) { mx = j; ci = e; }
  else { const int kc = s % kcs; s /= kcs; mx = s & 1; s >>= 1; ci = (2 * kc + j) * 8 + e; }
  const int my = s & 1; s >>= 1;
  int kz, ky, kx, co;
  if (a.up) {
    const int mz = s & 1;
    const int cls = n / a.cp; co = n % a.cp;
    const int rz = cls >> 2, ry = (cls >> 1) & 1, rx = cls & 1;
    kz = rz + 2 * (1 - mz); ky = ry + 2 * (1 - my); kx = rx + 2 * (1 - mx);
  } else {
    const int rr = s & 3; s >>= 2;
    kz = s;
    const int ry = rr >> 1, rx = rr & 1;
    const int kby = (ry + a.pad) & 1, kbx = (rx + a.pad) & 1;      // k = kb + 2 m'
    ky = kby + 2 * my; kx = kbx + 2 * mx;
    co = n;
  }
  float v = 0.f;
  if (co < a.cols) v = a.w[(long long)((kz * 4 + ky) * 4 + kx) * a.ws_tap + (long long)ci * a.ws_in + (long long)co * a.ws_out];
  a.dst[i] = __float2bfloat16_rn(v);
}

// B image of conv_down2_tc_kernel: [input-slice parity p][ry/rx class][m'y][(m'x, channel pair)] steps of [k half][n group][8][8] with
// N = 2 * CO columns: n < CO -> kz = p + 2, n >= CO -> kz = p
struct PackD2Args {
  const float* w; long long ws_tap, ws_in, ws_out;
  int pad, cin, cols, co, cin8;
  bf16* dst; int total;
};
__global__ void pack_weights_down2_kernel(const PackD2Args a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.total) return;
  int t = i;
  const int e = t & 7; t >>= 3;
  const int r = t & 7; t >>= 3;
  const int ng = (2 * a.co) >> 3;
  const int g = t % ng; t /= ng;
  const int j = t & 1; t >>= 1;
  int s = t;
  const int n = g * 8 + r;
  const int kcs = a.cin >> 4;
  int mx, ci;
  if (a.cin8) { mx = j; ci = e; }
  else { const int kc = s % kcs; s /= kcs; mx = s & 1; s >>= 1; ci = (2 * kc + j) * 8 + e; }
  const int my = s & 1; s >>= 1;
  const int rr = s & 3; s >>= 2;
  const int p = s;                                       // input-slice parity
  const int kz = (n < a.co) ? p + 2 : p;
  const int co = (n < a.co) ? n : n - a.co;
  const int ry = rr >> 1, rx = rr & 1;
  const int kby = (ry + a.pad) & 1, kbx = (rx + a.pad) & 1;      // k = kb + 2 m'
  const int ky = kby + 2 * my, kx = kbx + 2 * mx;
  float v = 0.f;
  if (co < a.cols) v = a.w[(long long)((kz * 4 + ky) * 4 + kx) * a.ws_tap + (long long)ci * a.ws_in + (long long)co * a.ws_out];
  a.dst[i] = __float2bfloat16_rn(v);
}

int cp_of(int cout) { return cout <= 8 ? 8 : (cout <= 16 ? 16 : 32); }
bool down2_on() { static const bool off = getenv("TEM_CONV_DOWN_V1") != nullptr; return !off; }      // debug knob: per-output-slice kernel
int npad_of(int cout) { return cout <= 16 ? 16 : 32; }
int steps_of(const ConvArgs& a) {
  const int cin = a.C0;
  const int per_yz = (cin == 8) ? 1 : 2 * (cin / 16);
  return (a.form == 1) ? 4 * per_yz : 32 * per_yz;
}
size_t resident_packed_bytes(const ConvArgs& a) {
  if (a.form != 1 && down2_on()) return (size_t)(steps_of(a) / 2) * (2 * cp_of(a.Cout)) * 32;     // conv_down2: half the steps, N = 2 * CO
  const int np = (a.form == 1) ? 8 * cp_of(a.Cout) : npad_of(a.Cout);
  return (size_t)steps_of(a) * np * 32;
}
int ring_of(const ConvArgs& a, size_t& smem_out) {
  const int cin = a.C0;
  const size_t wb = (resident_packed_bytes(a) + 1023) & ~(size_t)1023;
  const size_t slot = (size_t)(a.form == 1 ? 1 : 4) * (cin / 8) * SUB_STRIDE;
  const int want = (a.form == 1) ? 6 : 8, least = (a.form == 1) ? 3 : 5;
  int ring = want;
  while (ring > least && wb + ring * slot + 1024 > 200 * 1024) --ring;
  smem_out = wb + ring * slot + 1024;
  return ring;
}
// 0: resident-weight kernels (conv_up_tc / conv_down_tc), 1: wide UP, 2: wide DOWN, -1: shape not covered
int variant_of(const ConvArgs& a) {
  const int cin = a.C0;
  if ((cin == 8 || (cin % 16 == 0 && cin <= 64)) && a.Cout <= 32) {
    size_t smem; ring_of(a, smem);
    if (smem <= 200 * 1024) return 0;
  }
  if (a.form == 1 && cin >= 64 && cin % 64 == 0) return 1;
  // 32 -> 32 (weights do not fit beside the ring of the resident kernel): wide kernel unless the layer is tiny (d4 at wf = 8)
  const long long vox = (long long)a.B * a.L[0] * a.L[1] * a.L[2];
  if (a.form == 0 && cin >= 16 && cin % 16 == 0 && (a.Cout > 32 || cin > 64 || vox >= 4096)) return 2;
  return -1;
}

}  // namespace

size_t tc_s2_packed_bytes(const ConvArgs& a) {
  const int v = variant_of(a);
  if (v == 1) return (size_t)((a.Cout + 31) / 32) * (a.C0 / 64) * 32 * UW_KSTEP;
  if (v == 2) { const int np = dw_np(a.Cout); return (size_t)((a.Cout + np - 1) / np) * (a.C0 / 16) * 64 * np * 32; }
  return resident_packed_bytes(a);
}

const char* tc_s2_kernel_name(const ConvArgs& a) {
  const int v = variant_of(a);
  return v == 1 ? "conv_upw_tc_kernel" : v == 2 ? "conv_downw_tc_kernel" : (a.form == 1 ? "conv_up_tc_kernel" : "conv_down_tc_kernel");
}

bool tc_s2_supported(const ConvArgs& a) {
  for (int i = 0; i < 3; ++i) if (a.k[i] != 4 || a.stride[i] != 2 || a.conv_off[i]) return false;
  if (!(a.pad[0] == a.pad[1] && a.pad[1] == a.pad[2] && (a.pad[0] == 0 || a.pad[0] == 1))) return false;
  if (a.form != 0 && a.form != 1) return false;
  if (a.s0.dtype != DT_BF16 || a.out_dtype != DT_BF16) return false;
  if (a.s0.origins || a.use_lut || a.bias || a.C1) return false;
  if (a.s0.C != a.C0 || a.s0.coff != 0) return false;
  if (a.Cout % 8 || a.out_C % 8 || a.out_coff % 8) return false;
  if (a.ref && (a.ref_C % 8 || a.ref_coff % 8)) return false;
  if (variant_of(a) < 0) return false;
  return tem_get_encode() != nullptr;
}

cudaError_t tc_s2_pack_weights(const ConvArgs& a, bf16* dst, cudaStream_t st) {
  const int v = variant_of(a);
  if (v == 1) {
    PackUpwArgs q; q.w = a.w; q.ws_tap = a.ws_tap; q.ws_in = a.ws_in; q.ws_out = a.ws_out;
    q.nchunks = a.C0 / 64; q.cout = a.Cout; q.swz = wide_swizzle(); q.dst = dst; q.total = (long long)(tc_s2_packed_bytes(a) / 2);
    pack_weights_upw_kernel<<<(unsigned)((q.total + 255) / 256), 256, 0, st>>>(q); ++g_tem_launches;
    return cudaGetLastError();
  }
  if (v == 2) {
    PackDwArgs q; q.w = a.w; q.ws_tap = a.ws_tap; q.ws_in = a.ws_in; q.ws_out = a.ws_out;
    q.nch = a.C0 / 16; q.cout = a.Cout; q.np = dw_np(a.Cout); q.pad = a.pad[0]; q.dst = dst; q.total = (long long)(tc_s2_packed_bytes(a) / 2);
    pack_weights_dw_kernel<<<(unsigned)((q.total + 255) / 256), 256, 0, st>>>(q); ++g_tem_launches;
    return cudaGetLastError();
  }
  if (a.form != 1 && down2_on()) {
    PackD2Args q; q.w = a.w; q.ws_tap = a.ws_tap; q.ws_in = a.ws_in; q.ws_out = a.ws_out;
    q.pad = a.pad[0]; q.cin = a.C0; q.cols = a.Cout; q.co = cp_of(a.Cout); q.cin8 = a.C0 == 8;
    q.dst = dst; q.total = (int)(resident_packed_bytes(a) / 2);
    pack_weights_down2_kernel<<<(q.total + 255) / 256, 256, 0, st>>>(q); ++g_tem_launches;
    return cudaGetLastError();
  }
  PackS2Args p;
  p.w = a.w; p.ws_tap = a.ws_tap; p.ws_in = a.ws_in; p.ws_out = a.ws_out;
  p.up = a.form == 1; p.pad = a.pad[0]; p.cin = a.C0; p.cols = a.Cout; p.cp = cp_of(a.Cout);
  p.np = p.up ? 8 * p.cp : npad_of(a.Cout); p.cin8 = a.C0 == 8;
  p.dst = dst; p.total = (int)(resident_packed_bytes(a) / 2);
  pack_weights_s2_kernel<<<(p.total + 255) / 256, 256, 0, st>>>(p); ++g_tem_launches;
  return cudaGetLastError();
}

static bool make_map_s2(CUtensorMap* m, const void* base, int B, int Z, int Y, int X, int C, int es) {
  EncodeTiledFn enc = tem_get_encode();
  if (!enc) return false;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)Z, (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)X * C * 2, (cuuint64_t)Y * X * C * 2, (cuuint64_t)Z * Y * X * C * 2};
  // with element strides the box spans count * stride tensor elements and TMA keeps every stride-th one
  cuuint32_t box[5] = {8, (cuuint32_t)(SXV * es), (cuuint32_t)(SYR * es), 1, 1};
  cuuint32_t estr[5] = {1, (cuuint32_t)es, (cuuint32_t)es, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// [voxel][64 channels] tiles of 17 x 9 voxels in the 128B-swizzle layout (es = 2: every second voxel / row, the parity split)
static bool make_map_sw128(CUtensorMap* m, const void* base, int B, int Z, int Y, int X, int C, int es) {
  EncodeTiledFn enc = tem_get_encode();
  if (!enc) return false;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)Z, (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)X * C * 2, (cuuint64_t)Y * X * C * 2, (cuuint64_t)Z * Y * X * C * 2};
  cuuint32_t box[5] = {64, (cuuint32_t)(SXV * es), (cuuint32_t)(SYR * es), 1, 1};
  cuuint32_t estr[5] = {1, (cuuint32_t)es, (cuuint32_t)es, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// z chunking of the wide kernels (one CTA per SM): fewest waves x (slices per CTA + fixed cost)
static void pick_chunks(long long ctas_per_z, int q0, int zmax, int fixed, int& zc, int& nzc) {
  double best = 1e30; zc = zmax < q0 ? zmax : q0; 
  for (int cand = 1; cand <= zmax && cand <= q0; ++cand) {
    const int n = (q0 + cand - 1) / cand;
    const long long waves = (ctas_per_z * n + 147) / 148;
    const double cost = (double)waves * (cand + fixed);
    if (cost < best - 1e-9 || (cost < best + 1e-9 && cand > zc)) { best = cost; zc = cand; }
  }
  nzc = (q0 + zc - 1) / zc;
}

cudaError_t launch_conv_tc_s2(const ConvArgs& a, const bf16* wpacked, cudaStream_t st) {
  S2Args t; memset(&t, 0, sizeof(t));
  const int cin = a.C0;
  const bool up = a.form == 1;
  const int variant = variant_of(a);
  t.B = a.B; t.pad = a.pad[0];
  for (int i = 0; i < 3; ++i) {
    t.L[i] = a.L[i]; t.shift[i] = a.s0.shift[i];
    t.Q[i] = up ? ((a.L[i] - 1 + t.pad) >> 1) + 1 : a.L[i];
    t.out_off[i] = a.out_off[i]; t.ref_off[i] = a.ref_off[i];
  }
  t.planes = cin / 8; t.cin8 = cin == 8;
  t.wpacked = wpacked; t.wbytes = (int)tc_s2_packed_bytes(a);
  t.ntx = (t.Q[2] + TX - 1) / TX; t.nty = (t.Q[1] + TY - 1) / TY;
  const long long cols = (long long)a.B * t.ntx * t.nty;
  t.out = (bf16*)a.out; t.OZ = a.OZ; t.OY = a.OY; t.OX = a.OX; t.out_C = a.out_C; t.out_coff = a.out_coff;
  t.Cout = a.Cout; t.slope = a.slope;
  t.ref = a.ref; t.RZ = a.RZ; t.RY = a.RY; t.RX = a.RX; t.ref_C = a.ref_C; t.ref_coff = a.ref_coff; t.ref_slope = a.ref_slope;
  t.drop_key = a.drop_key; t.accumulate = a.accumulate;
  CUtensorMap m0;
  t.swz = wide_swizzle();
  t.dbg = tem_ablation_bits();
  if (variant == 1) {
    if (t.swz) { if (!make_map_sw128(&m0, a.s0.p, a.B, a.s0.Z, a.s0.Y, a.s0.X, a.s0.C, 1)) return cudaErrorInvalidValue; }
    else if (!tem_make_map_5d(&m0, a.s0.p, a.B, a.s0.Z, a.s0.Y, a.s0.X, a.s0.C, SXV, SYR)) return cudaErrorInvalidValue;
    const int groups = (a.Cout + 31) / 32;
    pick_chunks(cols * groups, t.Q[0], 1 << 20, 1, t.zc, t.nzc);
    const size_t smem = (size_t)UW_NIN * UW_IN_STAGE + (size_t)UW_WRING + 1024;
    static bool attr = false;
    if (!attr) { cudaError_t e = cudaFuncSetAttribute(conv_upw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e) return e; attr = true; }
    conv_upw_tc_kernel<<<dim3((unsigned)(cols * t.nzc), (unsigned)groups), kThreadsUpw, smem, st>>>(m0, t);
    ++g_tem_launches;
    return cudaGetLastError();
  }
  if (variant == 2) {
    if (!make_map_s2(&m0, a.s0.p, a.B, a.s0.Z, a.s0.Y, a.s0.X, a.s0.C, 2)) return cudaErrorInvalidValue;
    t.np = dw_np(a.Cout);
    const int groups = (a.Cout + t.np - 1) / t.np;
    t.ring = t.np <= 64 ? 8 : (t.np <= 128 ? 4 : 2);          // K-steps per weight stage
    pick_chunks(cols * groups, t.Q[0], 512 / t.np, 1, t.zc, t.nzc);
    const size_t smem = (size_t)DW_NIN * DW_STAGE + (size_t)DW_WMAX * DW_WSLOT + 1024;
    static bool attr = false;
    if (!attr) { cudaError_t e = cudaFuncSetAttribute(conv_downw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e) return e; attr = true; }
    conv_downw_tc_kernel<<<dim3((unsigned)(cols * t.nzc), (unsigned)groups), kThreadsDw, smem, st>>>(m0, t);
    ++g_tem_launches;
    return cudaGetLastError();
  }
  if (!up && down2_on()) {
    // input-slice-major kernel: the TMEM strip holds (zc + 2) column groups of CO columns; 256 columns keep two CTAs per SM
    // every input slice is consumed by one MMA group: a short ring suffices, which lets more CTAs share an SM
    const int co = cp_of(a.Cout);
    const size_t wb2 = (resident_packed_bytes(a) + 1023) & ~(size_t)1023;
    const size_t slot2 = (size_t)4 * (cin / 8) * SUB_STRIDE;
    static const char* ring_s = getenv("TEM_DOWN2_RING");        // debug knob
    t.ring = ring_s ? atoi(ring_s) : 4;
    if (t.ring < 2) t.ring = 2; if (t.ring > RING_MAX) t.ring = RING_MAX;
    while (t.ring > 2 && wb2 + t.ring * slot2 + 1024 > 200 * 1024) --t.ring;
    const size_t smem2 = wb2 + t.ring * slot2 + 1024;
    int sm_by_smem = (int)((227 * 1024) / (smem2 + 1024));
    if (sm_by_smem > 2048 / kThreads) sm_by_smem = 2048 / kThreads;
    if (sm_by_smem < 1) sm_by_smem = 1;
    // z chunks: fewest waves x (input slices per chunk + a fixed cost of ~6 slices for start-up / strip zeroing / drain)
    int zmax = 512 / co - 2;
    if (zmax > kDown2MaxZ) zmax = kDown2MaxZ;
    double best = 1e30; int best_nzc = (t.Q[0] + zmax - 1) / zmax;
    for (int nzc2 = (t.Q[0] + zmax - 1) / zmax; nzc2 <= t.Q[0]; ++nzc2) {
      const int zc2 = (t.Q[0] + nzc2 - 1) / nzc2;
      if (zc2 < 2 && nzc2 > 1) break;
      int tc2 = 32; while (tc2 < (zc2 + 2) * co) tc2 <<= 1;
      int per_sm = 512 / tc2; if (per_sm > sm_by_smem) per_sm = sm_by_smem;
      const long long ctas = cols * ((t.Q[0] + zc2 - 1) / zc2);
      const long long waves = (ctas + 148LL * per_sm - 1) / (148LL * per_sm);
      const double cost = (double)waves * (2 * zc2 + 2 + 6) * per_sm;      // per_sm CTAs share an SM's tensor pipe and load path
      if (cost < best - 1e-9) { best = cost; best_nzc = nzc2; }
    }
    t.zc = (t.Q[0] + best_nzc - 1) / best_nzc; t.nzc = (t.Q[0] + t.zc - 1) / t.zc;
    int tc = 32; while (tc < (t.zc + 2) * co) tc <<= 1;
    t.np = tc;
    if (!make_map_s2(&m0, a.s0.p, a.B, a.s0.Z, a.s0.Y, a.s0.X, a.s0.C, 2)) return cudaErrorInvalidValue;
    const unsigned grid2 = (unsigned)(cols * t.nzc);
    static bool attr2[12] = {};
    const int mode2 = (!a.ref && !a.accumulate) ? (a.drop_key ? 1 : 0) : ((!a.drop_key && a.slope == 1.f) ? 2 : 3);
#define LAUNCH_D2(COV, MD, IDX)                                                                                         \
    {                                                                                                                   \
      if (!attr2[IDX]) { cudaError_t e = cudaFuncSetAttribute(conv_down2_tc_kernel<COV, MD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); if (e) return e; attr2[IDX] = true; } \
      conv_down2_tc_kernel<COV, MD><<<grid2, kThreads, smem2, st>>>(m0, t);                                            \
    }
#define LAUNCH_D2_MODE(COV, IDX) { if (mode2 == 0) LAUNCH_D2(COV, 0, IDX) else if (mode2 == 1) LAUNCH_D2(COV, 1, IDX + 1) else if (mode2 == 2) LAUNCH_D2(COV, 2, IDX + 2) else LAUNCH_D2(COV, 3, IDX + 3) }
    if (co == 8) LAUNCH_D2_MODE(8, 0) else if (co == 16) LAUNCH_D2_MODE(16, 4) else LAUNCH_D2_MODE(32, 8)
#undef LAUNCH_D2_MODE
#undef LAUNCH_D2
    ++g_tem_launches;
    return cudaGetLastError();
  }
  size_t smem; t.ring = ring_of(a, smem);
  // z chunks: as many CTAs as are resident at once (one wave), chunks of at least two slices
  const int tmem_cols = up ? 16 * cp_of(a.Cout) : 2 * npad_of(a.Cout);
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm > 512 / tmem_cols) per_sm = 512 / tmem_cols;
  if (per_sm > 2048 / (up ? kThreadsUp : kThreads)) per_sm = 2048 / (up ? kThreadsUp : kThreads);
  if (up && per_sm > 2) per_sm = 2;              // register file: 320 threads x ~96 registers
  if (per_sm < 1) per_sm = 1;
  int nzc = (int)((148LL * per_sm) / cols);
  if (nzc < 1) nzc = 1;
  if (nzc > (t.Q[0] + 1) / 2) nzc = (t.Q[0] + 1) / 2;
  if (nzc < 1) nzc = 1;
  t.zc = (t.Q[0] + nzc - 1) / nzc; t.nzc = (t.Q[0] + t.zc - 1) / t.zc;
  if (up) { if (!tem_make_map_plane(&m0, &t.merged, a.s0.p, a.B, a.s0.Z, a.s0.Y, a.s0.X, a.s0.C, SXV, SYR)) return cudaErrorInvalidValue; }
  else if (!make_map_s2(&m0, a.s0.p, a.B, a.s0.Z, a.s0.Y, a.s0.X, a.s0.C, 2)) return cudaErrorInvalidValue;
  const unsigned grid = (unsigned)(cols * t.nzc);
  static bool attr[20] = {};
  // epilogue variant (see Epi::run): plain forward, forward + dropout, data gradient, anything else
  const int mode = (!a.ref && !a.accumulate) ? (a.drop_key ? 1 : 0) : ((!a.drop_key && a.slope == 1.f) ? 2 : 3);
#define LAUNCH_S2(KERNEL, IDX)                                                                                          \
  {                                                                                                                     \
    if (!attr[IDX]) { cudaError_t e = cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); if (e) return e; attr[IDX] = true; } \
    KERNEL<<<grid, up ? kThreadsUp : kThreads, smem, st>>>(m0, t);                                                                       \
  }
  if (up) {
    const int cp = cp_of(a.Cout);
#define LAUNCH_S2_MODE(K, P, IDX)                                                                                       \
  { if (mode == 0) LAUNCH_S2((K<P, 0>), IDX) else if (mode == 1) LAUNCH_S2((K<P, 1>), IDX + 1) else if (mode == 2) LAUNCH_S2((K<P, 2>), IDX + 2) else LAUNCH_S2((K<P, 3>), IDX + 3) }
    if (cp == 8) LAUNCH_S2_MODE(conv_up_tc_kernel, 8, 0) else if (cp == 16) LAUNCH_S2_MODE(conv_up_tc_kernel, 16, 4) else LAUNCH_S2_MODE(conv_up_tc_kernel, 32, 8)
  } else {
    if (npad_of(a.Cout) == 16) LAUNCH_S2_MODE(conv_down_tc_kernel, 16, 12) else LAUNCH_S2_MODE(conv_down_tc_kernel, 32, 16)
  }
#undef LAUNCH_S2_MODE
#undef LAUNCH_S2
  ++g_tem_launches;
  return cudaGetLastError();
}
