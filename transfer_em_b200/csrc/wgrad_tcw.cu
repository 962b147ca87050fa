// tcgen05 weight gradient of the WIDE 3x3x3 stride-1 layers (Ca multiple of 64, Cb multiple of 32: the wf <= 2 models of
// BASELINE config 4):   dw[(dz,dy,dx)][ca][cb] = sum_{b,v} x[b, v + (dz,dy,dx)][ca] * g[b, v][cb].
//
// With 64+ input channels the channel planes alone fill the M dimension, so the row trick of wgrad_tc.cu is not needed:
//   D_tap[M = 64|128 input channels][N = 32 output channels] += X_tap^T[M][K = 16 voxels of a row] * G[K][N]
// with MN-major operands read straight from the [voxel][8ch] planes TMA writes (M / N groups = channel planes, uniform
// plane stride; K groups = 8 consecutive voxels).  A tap (dy,dx) is a shifted start address of the x operand; dz is the
// z offset of the x tile (blockIdx.y), so one CTA owns 9 accumulators (288 TMEM columns) for (dz, 32 output channels,
// one block of input channels) and is persistent over (sample, 4 g-rows, 32-voxel column block, z chunk) work units.
// No MMA work is wasted (every (row, run, tap) MMA is useful), per MMA ~ (4 KB + 1 KB)/128 + 16 cycles.
// The 4x4x4 stride-2 layers (strided conv, transposed conv) use the same kernel in parity-class form (see wgrad_tc_s2.cu):
// blockIdx.y carries (class, m'z), the four (m'y, m'x) taps are the accumulators, x tiles are de-interleaved by TMA.
#include <cuda.h>
#include <string.h>
#include <stdlib.h>
#include "tem_kernels.cuh"
#include "ptx_sm100.cuh"

extern unsigned long long g_tem_launches;

namespace {

constexpr int XB = 2;                        // 16-voxel runs per column block
constexpr int RB = 4;                        // g rows per work unit
constexpr int RA = RB + 2;                   // x rows of a tile (stride 2 uses RB + 1 of them)
constexpr int WB = 16 * XB, WA = 16 * XB + 8;
constexpr int NBW = 32;                      // output channels per CTA (MMA N)
constexpr int XRW = 2, GRW = 3;              // ring depths
constexpr int kThreads = 192;

struct WwArgs {
  int B, L[3];
  int M;                       // 64 or 128 input channels per CTA
  int n_ca, n_cb;              // channel blocks
  int shift[3];
  int s2, pad;                 // 4x4x4 stride-2 mode: blockIdx.y carries (parity class, m'z) instead of dz; x tiles are de-interleaved by TMA
  int nrg, ncb, nzc, zc, units;
  int xa_bytes, gb_bytes;
  float* dw; long long ws_tap, ws_a, ws_b;
  int nb;                      // MMA N: output channels per CTA (32, 64 or 128)
  int split_dy;                // stride 1: blockIdx.y carries (dz, dy), the CTA owns the three dx taps (N >= 64 needs it: 9 x N > 512 columns)
  int swz;                     // tiles as [voxel][64 | 32 ch] rows in the 128B / 64B swizzle layouts (one TMA request per voxel)
  int dbg;                     // experiment bits (TEM_S2_DBG): 1 no epilogue atomics, 4 no x loads, 8 no g loads
};

__device__ __forceinline__ uint64_t desc_mn(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm volatile("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tcw_kernel(const __grid_constant__ CUtensorMap mapx, const __grid_constant__ CUtensorMap mapg, const WwArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t xfull[XRW], xempty[XRW], gfull[GRW], gempty[GRW], done_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* xring = smem;
  uint8_t* gring = smem + (size_t)XRW * a.xa_bytes;
  int y = blockIdx.y;
  const int cab = y % a.n_ca; y /= a.n_ca;
  const int cbb = y % a.n_cb; y /= a.n_cb;
  // stride 1: y = dz (nine (dy,dx) taps) or y = dz * 3 + dy (three dx taps).  stride 2: y = class * 2 + m'z with
  // class = (rz, ry, rx); k - pad = 2m + r, m = m_lo(r) + m'
  const int NB = a.nb;
  const int dz = a.s2 ? 0 : (a.split_dy ? y / 3 : y);
  const int dyf = (!a.s2 && a.split_dy) ? y % 3 : 0;
  const int mz = a.s2 ? (y & 1) : 0, cls = a.s2 ? (y >> 1) : 0;
  const int rz = cls >> 2, ry = (cls >> 1) & 1, rx = cls & 1;
  auto mlo = [&](int r) { return (((r + a.pad) & 1) - a.pad - r) / 2; };
  const int ntap = a.s2 ? 4 : (a.split_dy ? 3 : 9);
  const int ra = a.s2 ? RB + 1 : (a.split_dy ? RB : RA);
  const int pa = a.M >> 3;                       // planes of the x tile
  const int xplane = ra * WA * 16, gplane = RB * WB * 16;

  auto decode = [&](int u, int& b, int& y0, int& x0, int& z0, int& nz) {
    const int zc_i = u % a.nzc; u /= a.nzc;
    const int cb = u % a.ncb; u /= a.ncb;
    const int rg = u % a.nrg; u /= a.nrg;
    b = u; y0 = rg * RB; x0 = cb * WB; z0 = zc_i * a.zc;
    nz = min(a.zc, a.L[0] - z0);
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < XRW; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xempty[i], 1); }
    for (int i = 0; i < GRW; ++i) { mbar_init(&gfull[i], 1); mbar_init(&gempty[i], 1); }
    mbar_init(&done_bar, 1);
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
    // convergent producer: the whole warp walks the work list, one elected lane issues the TMA loads of a step
    int xslot = 0; uint32_t xph = 0; int gslot = 0; uint32_t gph = 0;
    for (int u = blockIdx.x; u < a.units; u += gridDim.x) {
      int b, y0, x0, z0, nz; decode(u, b, y0, x0, z0, nz);
      for (int s = 0; s < nz; ++s) {
        mbar_wait(&gempty[gslot], gph ^ 1u);
        mbar_wait(&xempty[xslot], xph ^ 1u);
        if (elect_one()) {
          if (a.dbg & 8) mbar_arrive(&gfull[gslot]);
          else {
          mbar_arrive_expect_tx(&gfull[gslot], (uint32_t)a.gb_bytes);
          uint8_t* gd = gring + (size_t)gslot * a.gb_bytes;
          if (a.swz) {
            if (NB == 32) tma_load_5d(gd, &mapg, &gfull[gslot], cbb * 32, x0, y0, z0 + s, b);
            else for (int c = 0; c < (NB >> 6); ++c) tma_load_5d(gd + c * (gplane * 8), &mapg, &gfull[gslot], cbb * NB + c * 64, x0, y0, z0 + s, b);
          } else {
            for (int p = 0; p < NB / 8; ++p) tma_load_5d(gd + p * gplane, &mapg, &gfull[gslot], (cbb * (NB / 8) + p) * 8, x0, y0, z0 + s, b);
          }
          }
          if (a.dbg & 4) { mbar_arrive(&xfull[xslot]); } else {
          mbar_arrive_expect_tx(&xfull[xslot], (uint32_t)a.xa_bytes);
          uint8_t* xd = xring + (size_t)xslot * a.xa_bytes;
          const int cx = a.s2 ? 2 * (x0 + mlo(rx)) + rx + a.shift[2] : x0 + a.shift[2];
          const int cy = a.s2 ? 2 * (y0 + mlo(ry)) + ry + a.shift[1] : y0 + dyf + a.shift[1];
          const int cz = a.s2 ? 2 * (z0 + s + mlo(rz) + mz) + rz + a.shift[0] : z0 + s + dz + a.shift[0];
          if (a.swz) {
            for (int c = 0; c < (a.M >> 6); ++c)
              tma_load_5d(xd + c * (xplane * 8), &mapx, &xfull[xslot], (cab * (a.M >> 6) + c) * 64, cx, cy, cz, b);
          } else
          for (int p = 0; p < pa; ++p)
            tma_load_5d(xd + p * xplane, &mapx, &xfull[xslot], (cab * pa + p) * 8, cx, cy, cz, b);
          }
        }
        __syncwarp();
        if (++gslot == GRW) { gslot = 0; gph ^= 1u; }
        if (++xslot == XRW) { xslot = 0; xph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(a.M >> 4) << 24);
    // plane layout (MN-major INTERLEAVE): LBO = 128 B (next 8 voxels along K), SBO = plane stride (next 8 channels along M / N).
    // swizzled rows (MN-major SW128 for x, SW64 for g): a row = one voxel (64 / 32 channels), SBO = 8 rows (next 8 voxels
    // along K), LBO = next 64-channel chunk of x (M = 128); a tap is still a start address shifted by whole rows.
    const uint32_t lo_fixed = (128u >> 4) << 16;                                   // LBO = 128 B (next 8 voxels)
    const uint32_t a_lo_sw = (((uint32_t)xplane * 8u) >> 4) << 16;                 // LBO = one 64-channel chunk of the x tile
    const uint32_t a_hi = a.swz ? ((1024u >> 4) | (1u << 14) | (2u << 29)) : (((uint32_t)xplane >> 4) | (1u << 14));
    const uint32_t b_hi = a.swz ? (NB == 32 ? ((512u >> 4) | (1u << 14) | (4u << 29)) : ((1024u >> 4) | (1u << 14) | (2u << 29)))
                                : (((uint32_t)gplane >> 4) | (1u << 14));   // SBO = plane stride
    const uint32_t b_lo_sw = (((uint32_t)gplane * 8u) >> 4) << 16;                 // LBO = one 64-channel chunk of the g tile (N = 128)
    const uint32_t xbase16 = smem_u32(xring) >> 4, gbase16 = smem_u32(gring) >> 4;
    const uint32_t xa16 = (uint32_t)a.xa_bytes >> 4, gb16 = (uint32_t)a.gb_bytes >> 4;
    int xslot = 0; uint32_t xph = 0; int gslot = 0; uint32_t gph = 0;
    uint32_t acc = 0u;
    // per-tap constants of the issue loop (kept out of it: an N = 32 MMA lasts ~50 cycles, a few extra instructions per issue show)
    const uint32_t xm = a.swz ? 8u : 1u, gm = a.swz ? (NB == 32 ? 4u : 8u) : 1u;   // 16 B units per voxel row
    uint32_t tap_off[9], tap_col[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int ty = a.s2 ? (t >> 1) : (a.split_dy ? 0 : t / 3), tx = a.s2 ? (t & 1) : (a.split_dy ? t : t % 3);
      tap_off[t] = (uint32_t)(ty * WA + tx) * xm;
      tap_col[t] = tmem_base + (uint32_t)(t * NB);
    }
    for (int u = blockIdx.x; u < a.units; u += gridDim.x) {
      int b, y0, x0, z0, nz; decode(u, b, y0, x0, z0, nz);
      for (int s = 0; s < nz; ++s) {
        mbar_wait(&gfull[gslot], gph);
        mbar_wait(&xfull[xslot], xph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t xs = a.swz ? ((xbase16 + (uint32_t)xslot * xa16) | a_lo_sw) : ((xbase16 + (uint32_t)xslot * xa16) | lo_fixed);
        const uint32_t gs = a.swz ? ((gbase16 + (uint32_t)gslot * gb16) | b_lo_sw) : ((gbase16 + (uint32_t)gslot * gb16) | lo_fixed);
        if (elect_one()) {
          // measured (A/B on one box): with nine taps the table form wins (g1.wgrad 752 vs 855 us), with three / four taps
          // the branchy form does (g7.wgrad 245 vs 261 us: the table form issues nine predicated UTCHMMA slots per run)
          if (ntap == 9) {
#pragma unroll 1
            for (int j = 0; j < RB; ++j) {
#pragma unroll
              for (int r = 0; r < XB; ++r) {
                const uint64_t bd = desc_mn(gs + (uint32_t)(j * WB + 16 * r) * gm, b_hi);
                const uint32_t xrow = xs + (uint32_t)(j * WA + 16 * r) * xm;
#pragma unroll
                for (int t = 0; t < 9; ++t) umma_bf16(tap_col[t], desc_mn(xrow + tap_off[t], a_hi), bd, idesc, acc);
                acc = 1u;
              }
            }
          } else {
#pragma unroll 1
            for (int j = 0; j < RB; ++j) {
#pragma unroll
              for (int r = 0; r < XB; ++r) {
                const uint64_t bd = desc_mn(gs + (uint32_t)(j * WB + 16 * r) * gm, b_hi);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  if (t < ntap) {
                    const int ty = a.s2 ? (t >> 1) : 0, tx = a.s2 ? (t & 1) : t;
                    umma_bf16(tmem_base + (uint32_t)(t * NB), desc_mn(xs + (uint32_t)((j + ty) * WA + 16 * r + tx) * xm, a_hi), bd, idesc, acc);
                  }
                }
                acc = 1u;
              }
            }
          }
          umma_commit(&xempty[xslot]);
          umma_commit(&gempty[gslot]);
        }
        __syncwarp();
        acc = 1u;
        if (++gslot == GRW) { gslot = 0; gph ^= 1u; }
        if (++xslot == XRW) { xslot = 0; xph ^= 1u; }
      }
    }
    if (elect_one()) umma_commit(&done_bar);
    __syncwarp();
  } else {
    // epilogue: row = input channel of this block, 9 accumulators x 32 output channels -> global atomics
    mbar_wait(&done_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;
    const int m = (a.M == 128) ? q * 32 + lane : q * 16 + (lane & 15);
    const bool rowok = (a.M == 128) || lane < 16;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const int ca = cab * a.M + m;
    const bool any = blockIdx.x < (unsigned)a.units;          // CTAs without work hold an uninitialised accumulator
    for (int t = 0; t < ntap; ++t) {
      int tap = a.split_dy ? dz * 9 + dyf * 3 + t : dz * 9 + t;
      if (a.s2) {
        const int kz = 2 * (mlo(rz) + mz) + rz + a.pad, ky = 2 * (mlo(ry) + (t >> 1)) + ry + a.pad, kx = 2 * (mlo(rx) + (t & 1)) + rx + a.pad;
        tap = (kz * 4 + ky) * 4 + kx;
      }
      float* dst = a.dw + (long long)tap * a.ws_tap + (long long)ca * a.ws_a + (long long)(cbb * NB) * a.ws_b;
      for (int c = 0; c < NB; c += 8) {
        uint32_t r[8];
        tmem_ld8(lane_base + (uint32_t)(t * NB + c), r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (rowok && any && !(a.dbg & 1)) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float v = __uint_as_float(r[e]);
            if (v != 0.f) atomicAdd(dst + (long long)(c + e) * a.ws_b, v);
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

}  // namespace

bool wgrad_tcw_supported(const WgradArgs& w) {
  if (w.S.dtype != DT_BF16 || w.p_dtype != DT_BF16 || w.S.origins || w.use_lut) return false;
  const bool s1 = w.k[0] == 3 && w.stride[0] == 1 && w.pad[0] == 0, s2 = w.k[0] == 4 && w.stride[0] == 2 && (w.pad[0] == 0 || w.pad[0] == 1);
  if (!s1 && !s2) return false;
  for (int i = 0; i < 3; ++i) if (w.k[i] != w.k[0] || w.stride[i] != w.stride[0] || w.pad[i] != w.pad[0] || w.p_off[i] != 0) return false;
  if (w.Ca < 64 || w.Ca % 64 || (w.Ca > 64 && w.Ca % 128) || w.Cb < 32 || w.Cb % 32) return false;
  if (w.S.C != w.Ca || w.S.coff != 0 || w.p_C != w.Cb || w.p_coff != 0) return false;
  if (w.PZ != w.L[0] || w.PY != w.L[1] || w.PX != w.L[2]) return false;
  if (w.p_bstride != (long long)w.L[0] * w.L[1] * w.L[2] * w.Cb) return false;
  if (w.S.bstride != (long long)w.S.Z * w.S.Y * w.S.X * w.S.C) return false;
  return tem_get_encode() != nullptr;
}

cudaError_t launch_wgrad_tcw(const WgradArgs& w, cudaStream_t st) {
  if ((long long)w.B * w.L[0] * w.L[1] * w.L[2] == 0) return cudaSuccess;
  WwArgs t; memset(&t, 0, sizeof(t));
  t.B = w.B; for (int i = 0; i < 3; ++i) { t.L[i] = w.L[i]; t.shift[i] = w.S.shift[i]; }
  t.s2 = w.stride[0] == 2; t.pad = w.pad[0];
  t.M = w.Ca >= 128 ? 128 : 64;
  static const bool v1 = getenv("TEM_WGRAD_TCW_V1") != nullptr;      // debug knob: N = 32 everywhere, nine taps per CTA in the stride-1 form
  t.nb = v1 ? 32 : (w.Cb % 128 == 0 ? 128 : (w.Cb % 64 == 0 ? 64 : 32));
  if (!t.s2 && t.M == 64 && t.nb == 64) t.nb = 32;     // 64 -> 64 (g1): measured 748 us with nine N = 32 taps per CTA, 879 us split
  t.split_dy = !t.s2 && t.nb > 32;
  t.n_ca = w.Ca / t.M; t.n_cb = w.Cb / t.nb;
  const int ra = t.s2 ? RB + 1 : (t.split_dy ? RB : RA);
  t.xa_bytes = (t.M / 8) * ra * WA * 16; t.gb_bytes = (t.nb / 8) * RB * WB * 16;
  t.dw = w.dw; t.ws_tap = w.ws_tap; t.ws_a = w.ws_a; t.ws_b = w.ws_b;
  t.dbg = tem_ablation_bits();
  t.nrg = (w.L[1] + RB - 1) / RB; t.ncb = (w.L[2] + WB - 1) / WB;
  const int gy = (t.s2 ? 16 : (t.split_dy ? 9 : 3)) * t.n_ca * t.n_cb;
  int gx = 148 / gy; if (gx < 1) gx = 1;
  // z chunks: enough units for the persistent CTAs of one (dz, channel block) group
  const long long cols = (long long)w.B * t.nrg * t.ncb;
  int nzc = 1;
  while (cols * nzc < 4LL * gx && (w.L[0] + nzc) / (nzc + 1) >= 4) ++nzc;
  t.zc = (w.L[0] + nzc - 1) / nzc; t.nzc = (w.L[0] + t.zc - 1) / t.zc;
  t.units = (int)(cols * t.nzc);
  if (gx > t.units) gx = t.units;
  static const int swz = getenv("TEM_WGRAD_TCW_NO_SWIZZLE") ? 0 : 1;     // debug knob: 8-channel plane tiles
  t.swz = swz;
  CUtensorMap mx, mg;
  {     // x: (de-interleaved in the stride-2 form: every second voxel of every second row through TMA element strides, as in conv_tc_s2.cu)
    EncodeTiledFn enc = tem_get_encode();
    if (!enc) return cudaErrorInvalidValue;
    const int es = t.s2 ? 2 : 1;
    cuuint64_t dims[5] = {(cuuint64_t)w.S.C, (cuuint64_t)w.S.X, (cuuint64_t)w.S.Y, (cuuint64_t)w.S.Z, (cuuint64_t)w.B};
    cuuint64_t strides[4] = {(cuuint64_t)w.S.C * 2, (cuuint64_t)w.S.X * w.S.C * 2, (cuuint64_t)w.S.Y * w.S.X * w.S.C * 2, (cuuint64_t)w.S.Z * w.S.Y * w.S.X * w.S.C * 2};
    cuuint32_t box[5] = {(cuuint32_t)(swz ? 64 : 8), (cuuint32_t)(WA * es), (cuuint32_t)(ra * es), 1, 1};
    cuuint32_t estr[5] = {1, (cuuint32_t)es, (cuuint32_t)es, 1, 1};
    if (enc(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(w.S.p), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return cudaErrorInvalidValue;
    cuuint64_t gdims[5] = {(cuuint64_t)w.p_C, (cuuint64_t)w.PX, (cuuint64_t)w.PY, (cuuint64_t)w.PZ, (cuuint64_t)w.B};
    cuuint64_t gstr[4] = {(cuuint64_t)w.p_C * 2, (cuuint64_t)w.PX * w.p_C * 2, (cuuint64_t)w.PY * w.PX * w.p_C * 2, (cuuint64_t)w.PZ * w.PY * w.PX * w.p_C * 2};
    cuuint32_t gbox[5] = {(cuuint32_t)(swz ? (t.nb == 32 ? 32 : 64) : 8), (cuuint32_t)WB, (cuuint32_t)RB, 1, 1};
    cuuint32_t ones[5] = {1, 1, 1, 1, 1};
    if (enc(&mg, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(w.P), gdims, gstr, gbox, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
            swz ? (t.nb == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B) : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return cudaErrorInvalidValue;
  }
  const size_t smem = (size_t)XRW * t.xa_bytes + (size_t)GRW * t.gb_bytes + 1024;
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  static bool attr = false;
  if (!attr) { cudaError_t e = cudaFuncSetAttribute(wgrad_tcw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); if (e) return e; attr = true; }
  wgrad_tcw_kernel<<<dim3((unsigned)gx, (unsigned)gy), kThreads, smem, st>>>(mx, mg, t); ++g_tem_launches;
  return cudaGetLastError();
}
