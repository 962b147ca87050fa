// tcgen05 weight gradient of the 4x4x4 stride-2 layers (strided conv, models/utils.py:80; Conv3DTranspose, :129-130):
//   dw[k][ca][cb] = sum_{b,p} S[b, 2p + k - pad][ca] * P[b, p][cb]                        
// (conv: S = layer input, P = dy, pad 0; transposed conv: S = dy, P = layer input, pad 1).
//
// k - pad = 2m + r splits every axis into a parity r and a two-tap offset m in {m_lo(r), m_lo(r)+1}:
//   dw[k(r,m')] = sum_p S_r[p + m_lo + m'] * P[p],   S_r = every second voxel of S,
// i.e. each of the 8 parity classes is a 2x2x2-tap stride-1 correlation of a de-interleaved tile of S with P, and the
// machinery of wgrad_tc.cu applies with 2 taps per axis: MN-major operands straight from [voxel][8ch] planes, the
// y tap carried by uniformly strided tile rows (two useful diagonals), the x tap by a shifted start address of the P
// operand, the z tap by the previous P slice.  The de-interleaved S_r tiles are produced by TMA itself (elementStrides
// = 2 along x and y, as in conv_tc_s2.cu).  A CTA owns one (rz, ry) parity pair (blockIdx.y) and both x parities:
// 8 accumulators (rx, m'z, m'x) of N <= 64 columns stay in TMEM for the whole persistent CTA.
#include <cuda.h>
#include <string.h>
#include <stdlib.h>
#include "tem_kernels.cuh"
#include "ptx_sm100.cuh"

extern unsigned long long g_tem_launches;

namespace {

constexpr int XR_MAX = 6, DR_MAX = 8;
constexpr int kThreads = 192;

struct WsArgs {
  int B, L[3];                 // extent of P (z,y,x)
  int pa, pb, RA, RB, M, N, NR, WA, WB;
  int shift[3], pad;
  int nrg, nzc, zc, units, XR, DR, tmem_cols;
  int xa_bytes;                // one class tile of S_r (pa planes x RA rows x WA voxels)
  int gb_bytes;                // one P slice tile
  float* dw; long long ws_tap, ws_a, ws_b;
};

__device__ __host__ __forceinline__ int mlo_of(int r, int pad) { return (((r + pad) & 1) - pad - r) / 2; }

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_s2_kernel(const __grid_constant__ CUtensorMap mapx, const __grid_constant__ CUtensorMap mapg, const WsArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t xfull[XR_MAX], xempty[XR_MAX], gfull[DR_MAX], gempty[DR_MAX], done_bar;
  __shared__ uint32_t tmem_base_s;
  const int XR = a.XR, DR = a.DR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int xslot_bytes = 2 * a.xa_bytes;                // both x parities
  uint8_t* xring = smem;
  uint8_t* gring = smem + (size_t)XR * xslot_bytes;
  const int rz = blockIdx.y >> 1, ry = blockIdx.y & 1;
  const int mlz = mlo_of(rz, a.pad), mly = mlo_of(ry, a.pad);

  auto decode = [&](int u, int& b, int& y0, int& tz0, int& ntz) {
    const int zc_i = u % a.nzc; u /= a.nzc;
    const int rg = u % a.nrg; u /= a.nrg;
    b = u; y0 = rg * a.RB; tz0 = zc_i * a.zc;
    ntz = min(a.zc, a.L[0] + 1 - tz0);
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < XR; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xempty[i], 1); }
    for (int i = 0; i < DR; ++i) { mbar_init(&gfull[i], 1); mbar_init(&gempty[i], 1); }
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)a.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      int gslot = 0; uint32_t gph = 0; int xslot = 0; uint32_t xph = 0;
      for (int u = blockIdx.x; u < a.units; u += gridDim.x) {
        int b, y0, tz0, ntz; decode(u, b, y0, tz0, ntz);
        // P slices tz0-1 .. tz0+ntz-1 (index i); step s (tz = tz0+s) reads index s+1 (m'z = 0) and s (m'z = 1)
        auto load_g = [&](int i) {
          mbar_wait(&gempty[gslot], gph ^ 1u);
          mbar_arrive_expect_tx(&gfull[gslot], (uint32_t)a.gb_bytes);
          uint8_t* dst = gring + (size_t)gslot * a.gb_bytes;
          const int plane_bytes = a.RB * a.WB * 16;
          for (int p = 0; p < a.pb; ++p)
            tma_load_5d(dst + p * plane_bytes, &mapg, &gfull[gslot], p * 8, -1, y0, tz0 - 1 + i, b);
          if (++gslot == DR) { gslot = 0; gph ^= 1u; }
        };
        load_g(0);
        for (int s = 0; s < ntz; ++s) {
          load_g(s + 1);
          mbar_wait(&xempty[xslot], xph ^ 1u);
          mbar_arrive_expect_tx(&xfull[xslot], (uint32_t)xslot_bytes);
          uint8_t* dst = xring + (size_t)xslot * xslot_bytes;
          const int plane_bytes = a.RA * a.WA * 16;
          const int cy = 2 * (y0 + mly) + ry + a.shift[1], cz = 2 * (tz0 + s + mlz) + rz + a.shift[0];
          for (int rx = 0; rx < 2; ++rx) {
            const int cx = 2 * mlo_of(rx, a.pad) + rx + a.shift[2];
            for (int p = 0; p < a.pa; ++p)
              tma_load_5d(dst + rx * a.xa_bytes + p * plane_bytes, &mapx, &xfull[xslot], p * 8, cx, cy, cz, b);
          }
          if (++xslot == XR) { xslot = 0; xph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(a.N >> 3) << 17) | ((uint32_t)(a.M >> 4) << 24);
    const uint32_t lo_fixed = (128u >> 4) << 16;                                   // LBO = 128 B (next 8 voxels)
    const uint32_t a_hi = (((uint32_t)a.WA * 16u) >> 4) | (1u << 14), b_hi = (((uint32_t)a.WB * 16u) >> 4) | (1u << 14);
    const uint32_t xbase16 = smem_u32(xring) >> 4, gbase16 = smem_u32(gring) >> 4;
    const uint32_t xs16 = (uint32_t)xslot_bytes >> 4, xa16 = (uint32_t)a.xa_bytes >> 4, gb16 = (uint32_t)a.gb_bytes >> 4;
    const uint32_t N = (uint32_t)a.N;
    int gwslot = 0; uint32_t gwph = 0;
    int xslot = 0; uint32_t xph = 0;
    int gold = 0;
    uint32_t acc = 0u;
    auto mma = [&](uint32_t d, uint32_t alo, uint32_t blo, uint32_t accf) {
      if (!elect_one()) return;
      uint64_t ad, bd;
      asm volatile("mov.b64 %0, {%1, %2};" : "=l"(ad) : "r"(alo), "r"(a_hi));
      asm volatile("mov.b64 %0, {%1, %2};" : "=l"(bd) : "r"(blo), "r"(b_hi));
      umma_bf16(d, ad, bd, idesc, accf);
    };
    for (int u = blockIdx.x; u < a.units; u += gridDim.x) {
      int b, y0, tz0, ntz; decode(u, b, y0, tz0, ntz);
      mbar_wait(&gfull[gwslot], gwph); if (++gwslot == DR) { gwslot = 0; gwph ^= 1u; }
      for (int s = 0; s < ntz; ++s) {
        mbar_wait(&gfull[gwslot], gwph); if (++gwslot == DR) { gwslot = 0; gwph ^= 1u; }
        mbar_wait(&xfull[xslot], xph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        int g1s = gold + 1; if (g1s >= DR) g1s -= DR;
        // m'z = 0 reads P slice tz (index s+1), m'z = 1 reads tz-1 (index s); +1 voxel: the P tile starts at x = -1
        const uint32_t b0 = (gbase16 + (uint32_t)g1s * gb16 + 1u) | lo_fixed;
        const uint32_t b1 = (gbase16 + (uint32_t)gold * gb16 + 1u) | lo_fixed;
        uint32_t alo = (xbase16 + (uint32_t)xslot * xs16) | lo_fixed;
        for (int r = 0; r < a.NR; ++r) {
          const uint32_t ro = (uint32_t)r * 16u;
          // accumulator index = (rx*2 + m'z)*2 + m'x
          mma(tmem_base + 0 * N, alo, b0 + ro, acc);        mma(tmem_base + 1 * N, alo, b0 + ro - 1u, acc);
          mma(tmem_base + 2 * N, alo, b1 + ro, acc);        mma(tmem_base + 3 * N, alo, b1 + ro - 1u, acc);
          mma(tmem_base + 4 * N, alo + xa16, b0 + ro, acc); mma(tmem_base + 5 * N, alo + xa16, b0 + ro - 1u, acc);
          mma(tmem_base + 6 * N, alo + xa16, b1 + ro, acc); mma(tmem_base + 7 * N, alo + xa16, b1 + ro - 1u, acc);
          alo += 16u; acc = 1u;
        }
        if (elect_one()) {
          umma_commit(&xempty[xslot]);
          umma_commit(&gempty[gold]);
          if (s == ntz - 1) umma_commit(&gempty[g1s]);
        }
        __syncwarp();
        if (++xslot == XR) { xslot = 0; xph ^= 1u; }
        if (++gold == DR) gold = 0;
      }
      if (++gold == DR) gold = 0;
    }
    if (elect_one()) umma_commit(&done_bar);
    __syncwarp();
  }
  // ---- epilogue: 16 local taps (m'z, m'y, rx, m'x) x Ca x Cb image in shared memory, one atomic per weight and CTA
  float* red = reinterpret_cast<float*>(smem);
  const int Ca = a.pa * 8, Cb = a.pb * 8;
  const int CbP = Cb + 4;                              // padded image rows: the float4 read-modify-writes of 8 consecutive ca fall into different bank groups
  const int nred = 16 * Ca * Cb;
  if (warp >= 2) {
    mbar_wait(&done_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int i = threadIdx.x - 64; i < 16 * Ca * CbP; i += 128) red[i] = 0.f;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int q = warp & 3;
    const int m = (a.M == 128) ? q * 32 + lane : q * 16 + (lane & 15);
    const bool rowok = (a.M == 128) || lane < 16;
    const int gm = m >> 3, pA = gm / a.RA, gi = gm % a.RA, ca = pA * 8 + (m & 7);
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int j = 0; j < a.RB; ++j) {
      const int my = gi - j;
      const bool use = rowok && my >= 0 && my < 2;
      for (int acc = 0; acc < 8; ++acc) {
        const int rx = acc >> 2, mz = (acc >> 1) & 1, mx = acc & 1;
        float* const rowp = red + ((size_t)(((mz * 2 + my) * 2 + rx) * 2 + mx) * Ca + ca) * CbP;
        for (int pB = 0; pB < a.pb; pB += 2) {              // two 8-column loads in flight per wait
          uint32_t r[16];
          tmem_ld8(lane_base + (uint32_t)(acc * a.N + (pB * a.RB + j) * 8), r);
          if (pB + 1 < a.pb) tmem_ld8(lane_base + (uint32_t)(acc * a.N + ((pB + 1) * a.RB + j) * 8), r + 8);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (use) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (pB + h < a.pb) {
                float4* dst = reinterpret_cast<float4*>(rowp + (pB + h) * 8);
                float4 v0 = dst[0], v1 = dst[1];
                v0.x += __uint_as_float(r[8 * h + 0]); v0.y += __uint_as_float(r[8 * h + 1]); v0.z += __uint_as_float(r[8 * h + 2]); v0.w += __uint_as_float(r[8 * h + 3]);
                v1.x += __uint_as_float(r[8 * h + 4]); v1.y += __uint_as_float(r[8 * h + 5]); v1.z += __uint_as_float(r[8 * h + 6]); v1.w += __uint_as_float(r[8 * h + 7]);
                dst[0] = v0; dst[1] = v1;
              }
            }
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    // one vector reduction per four weights and CTA (Ca, Cb are powers of two: shifts instead of divisions); every CTA starts
    // its pass at a different offset (fewer same-address collisions in L2)
    const int lg_cb4 = (Cb == 8) ? 1 : (Cb == 16 ? 2 : 3), lg_ca = (Ca == 8) ? 3 : (Ca == 16 ? 4 : 5);
    const int n4 = nred >> 2;
    const bool vec4 = a.ws_b == 1 && (reinterpret_cast<uintptr_t>(a.dw) & 15) == 0 && (a.ws_tap & 3) == 0 && (a.ws_a & 3) == 0;
    const int rot = (int)(((long long)blockIdx.x * n4 / gridDim.x) & ~127LL);
    for (int i0 = threadIdx.x - 64; i0 < n4; i0 += 128) {
      int i = i0 + rot; if (i >= n4) i -= n4;
      const float4 v = *reinterpret_cast<const float4*>(red + (size_t)(i >> lg_cb4) * CbP + (i & ((1 << lg_cb4) - 1)) * 4);
      if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) {
        const int c4 = i & ((1 << lg_cb4) - 1); int t = i >> lg_cb4; const int cA = t & (Ca - 1); t >>= lg_ca;
        const int mx = t & 1, rx = (t >> 1) & 1, my = (t >> 2) & 1, mz = t >> 3;
        const int kz = 2 * (mlz + mz) + rz + a.pad, ky = 2 * (mly + my) + ry + a.pad, kx = 2 * (mlo_of(rx, a.pad) + mx) + rx + a.pad;
        float* dst = a.dw + (long long)((kz * 4 + ky) * 4 + kx) * a.ws_tap + (long long)cA * a.ws_a + (long long)(c4 * 4) * a.ws_b;
        if (vec4) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        } else {
          atomicAdd(dst, v.x); atomicAdd(dst + a.ws_b, v.y); atomicAdd(dst + 2 * a.ws_b, v.z); atomicAdd(dst + 3 * a.ws_b, v.w);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols) : "memory");
  }
}

bool plan(const WgradArgs& w, WsArgs& t, size_t& smem) {
  memset(&t, 0, sizeof(t));
  t.B = w.B; t.pad = w.pad[0];
  for (int i = 0; i < 3; ++i) { t.L[i] = w.L[i]; t.shift[i] = w.S.shift[i]; }
  t.pa = w.Ca / 8; t.pb = w.Cb / 8;
  t.M = (t.pa == 4) ? 128 : 64;
  t.RA = (t.M / 8) / t.pa;
  int rb = t.RA - 1;
  while (rb > 1 && rb * w.Cb > 64) --rb;
  if (t.M == 128) while (rb > 1 && (rb * w.Cb) % 16) --rb;
  t.RB = rb; t.N = rb * w.Cb;
  if (t.N > 64 || (t.M == 128 && t.N % 16)) return false;
  t.NR = (w.L[2] + 1 + 15) / 16;
  t.WA = 16 * t.NR; t.WB = 16 * t.NR + 8;
  if (t.WA * 2 > 256 || t.WB > 256) return false;
  t.xa_bytes = t.pa * t.RA * t.WA * 16; t.gb_bytes = t.pb * t.RB * t.WB * 16;
  int cols = 32; while (cols < 8 * t.N) cols <<= 1;
  t.tmem_cols = cols;
  t.XR = 2; t.DR = 4;
  while (t.XR < XR_MAX && (size_t)(t.XR + 1) * 2 * t.xa_bytes + (size_t)(t.DR + 1) * t.gb_bytes <= 150 * 1024) { ++t.XR; ++t.DR; }
  smem = (size_t)t.XR * 2 * t.xa_bytes + (size_t)t.DR * t.gb_bytes + 1024;
  const size_t red = (size_t)16 * w.Ca * (w.Cb + 4) * 4;
  if (red + 1024 > smem) smem = red + 1024;
  return smem <= 200 * 1024;
}

bool make_map_strided(CUtensorMap* m, const void* base, int B, int Z, int Y, int X, int C, int wa, int ra) {
  EncodeTiledFn enc = tem_get_encode();
  if (!enc) return false;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)Z, (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)X * C * 2, (cuuint64_t)Y * X * C * 2, (cuuint64_t)Z * Y * X * C * 2};
  cuuint32_t box[5] = {8, (cuuint32_t)(wa * 2), (cuuint32_t)(ra * 2), 1, 1};
  cuuint32_t estr[5] = {1, 2, 2, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

bool wgrad_tc_s2_supported(const WgradArgs& w) {
  if (w.S.dtype != DT_BF16 || w.p_dtype != DT_BF16 || w.S.origins || w.use_lut) return false;
  for (int i = 0; i < 3; ++i) if (w.k[i] != 4 || w.stride[i] != 2 || w.p_off[i] != 0 || w.pad[i] != w.pad[0]) return false;
  if (w.pad[0] != 0 && w.pad[0] != 1) return false;
  if (!(w.Ca == 8 || w.Ca == 16 || w.Ca == 32) || !(w.Cb == 8 || w.Cb == 16 || w.Cb == 32)) return false;
  if (w.S.C != w.Ca || w.S.coff != 0 || w.p_C != w.Cb || w.p_coff != 0) return false;
  if (w.PZ != w.L[0] || w.PY != w.L[1] || w.PX != w.L[2]) return false;
  if (w.p_bstride != (long long)w.L[0] * w.L[1] * w.L[2] * w.Cb) return false;
  if (w.S.bstride != (long long)w.S.Z * w.S.Y * w.S.X * w.S.C) return false;
  static const char* mv_s = getenv("TEM_WS2_MINVOX");          // debug knob
  if ((long long)w.L[0] * w.L[1] * w.L[2] < (mv_s ? atoi(mv_s) : 128)) return false;     // one-voxel layers (d6): the per-CTA epilogue outweighs the MMAs; d4 (6^3) is 26 us here against 45 us on wgrad_mma since the epilogue leaves as red.v4
  WsArgs t; size_t smem;
  if (!plan(w, t, smem)) return false;
  return tem_get_encode() != nullptr;
}

cudaError_t launch_wgrad_tc_s2(const WgradArgs& w, cudaStream_t st) {
  WsArgs t; size_t smem;
  if (!plan(w, t, smem)) return cudaErrorInvalidConfiguration;
  if ((long long)w.B * w.L[0] * w.L[1] * w.L[2] == 0) return cudaSuccess;
  t.dw = w.dw; t.ws_tap = w.ws_tap; t.ws_a = w.ws_a; t.ws_b = w.ws_b;
  t.nrg = (w.L[1] + t.RB - 1) / t.RB;                // the work is partitioned by P rows: each carries both of its y taps
  const long long cols = (long long)w.B * t.nrg;
  const int nslices = w.L[0] + 1;
  // 4 (rz, ry) CTA groups share the 148 SMs: 37 persistent CTAs each
  const int per_group = 37;
  int best = 1; double best_eff = -1.0;
  for (int nzc = 1; nzc <= nslices && nzc <= 64; ++nzc) {
    const int zc = (nslices + nzc - 1) / nzc;
    if (zc < 3 && nzc > 1) break;
    const int real = (nslices + zc - 1) / zc;
    const long long units = cols * real;
    const long long waves = (units + per_group - 1) / per_group;
    const double eff = (double)units / (double)(waves * per_group) * (double)zc / (double)(zc + 1);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = nzc; }
  }
  t.zc = (nslices + best - 1) / best; t.nzc = (nslices + t.zc - 1) / t.zc;
  t.units = (int)(cols * t.nzc);
  CUtensorMap mx, mg;
  if (!make_map_strided(&mx, w.S.p, w.B, w.S.Z, w.S.Y, w.S.X, w.S.C, t.WA, t.RA)) return cudaErrorInvalidValue;
  if (!tem_make_map_5d(&mg, w.P, w.B, w.PZ, w.PY, w.PX, w.p_C, t.WB, t.RB)) return cudaErrorInvalidValue;
  static bool attr = false;
  if (!attr) { cudaError_t e = cudaFuncSetAttribute(wgrad_tc_s2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); if (e) return e; attr = true; }
  const long long mmas = (long long)t.units * t.zc * t.NR * 8;
  long long want = mmas / 96;
  if (want < 1) want = 1; if (want > per_group) want = per_group; if (want > t.units) want = t.units;
  wgrad_tc_s2_kernel<<<dim3((unsigned)want, 4), kThreads, smem, st>>>(mx, mg, t); ++g_tem_launches;
  return cudaGetLastError();
}
