// tcgen05 weight gradient of the single-channel ends of the networks (3x3x3, stride 1):
//   dw[dz][dy][dx][cb] = sum_{b,p} s[b, p + (dz,dy,dx)] * P[b, p][cb]
//   * first layers g0 / d0 (generator.py:54, discriminator.py:39): s = the one-channel input (uint8 through the standardise
//     LUT, fp32 / bf16 fakes with virtual zero padding), P = dy with Cb = 8 channels;
//   * last generator layer g11 (generator.py:110) with the operands swapped by launch_wgrad_c1 (s = d(output) padded by
//     2, P = a10 with Cb = 16 channels, taps flipped).
// wgrad_c1.cu does this with packed FMAs at ~17 TFLOP/s (g0: 77 us at batch 8).  Here the reduction over voxels is the K
// dimension of an MMA and the dz taps are folded into the N dimension (the transpose of conv_c1tc.cu's Toeplitz trick):
//   D_{dy,dx}[(jz, cb)][iz] += sum_{x < 16} P[z0 + jz][y][x0 + x][cb] * s[z0 + iz][y + dy][x0 + x + dx]      (one MMA)
//   dw[dz][dy][dx][cb] = sum_{jz} D_{dy,dx}[(jz, cb)][jz + dz]                                     (once per CTA)
// M = 128 = (16 planes x 8 channels) or (2 channel groups x 8 planes x 8), N = 32 / 16 input planes, K = 16 consecutive x.
//   * A = P (MN-major, no swizzle): copied with cp.async into [plane slot][y][x][16 B]: the 8 K-rows of a core matrix are 8
//     consecutive x (16 B apart: the natural channels-last layout), M groups are plane slots (SBO = one y-x plane);
//   * B = s (MN-major): the same bf16 tile layout as conv_c1tc.cu, [z chunk of 8 planes][row = hy*18 + hx][8 planes]: N groups
//     are z chunks (SBO = rows*16 B), the K-rows 16 consecutive hx, the (dy, dx) tap a row offset of the start address.
//   * nine accumulators D_{dy,dx} (N columns each) stay in TMEM for the whole CTA; the diagonal fold, a shared-memory
//     reduction over jz and one atomicAdd per weight and CTA happen once at the end.
// s keeps fp32-like precision as a bf16 pair (hi = bf16(s), lo = bf16(s - hi)): the hi and lo planes of a window are
// stacked along N (N = 2 x 24 or 2 x 16 columns, [hi | lo] chunks of the window's own tile image), so one MMA covers both
// and the fold adds the two halves; the result equals the CUDA-core kernels' fp32 s x bf16 P to 2^-17.
// CTAs are persistent over work items (sample, y block of 8, x block of 16, z range of W windows), one per SM.
#include <cuda.h>
#include <string.h>
#include <stdlib.h>
#include "tem_kernels.cuh"
#include "ptx_sm100.cuh"

extern unsigned long long g_tem_launches;

namespace {

constexpr int WT_TX = 16, WT_HX = 18, WT_TY = 8, WT_HY = 10;
constexpr int WT_ROWS = WT_HY * WT_HX;       // 180 staged rows (hy, hx)
constexpr int WT_R = 184;                    // row pitch of a z chunk
constexpr int WT_PB = 5;                     // chunks gathered per producer thread and batch (a whole item: one L2 round trip)
constexpr int WT_NP = 256;                   // producer threads
constexpr int WT_THREADS = 32 + WT_NP;       // MMA warp + 8 producer warps (the first four also fold the accumulators at the end)
constexpr int WT_WIN_BYTES = 16 * WT_TY * WT_TX * 16;     // one window of P: 16 plane slots x 8 rows x 16 voxels x 16 B

struct Wc1Args {
  SrcView S; int use_lut; float lut_mean, lut_std;
  const bf16* P; int PZ, PY, PX, p_C, p_coff, p_off[3]; long long p_bstride;
  int Cb, nzw, N, W;           // channels of P (8 / 16), planes per window (128 / Cb), MMA N = 2 * nch * 8, windows per item
  int nch;                     // z chunks a window reads: nzw / 8 + 1
  int B, L[3];
  int nyb, nxb, nzr, items;
  int gbytes, tbytes;          // bytes of one P / s buffer (s: per window [hi chunks | lo chunks])
  float* dw; long long ws_tap, ws_b;
  int dbg;                     // stage-ablation bits (-DTEM_ABLATION builds only): 2 no atomics, 4 no s gather, 8 no P copies, 16 no MMAs
};

__device__ __forceinline__ void wc_tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}
// x -> (hi | lo << 16), hi = bf16(x), lo = bf16(x - hi)
__device__ __forceinline__ uint32_t wc_split_bf16(float x) {
  const bf16 hi = __float2bfloat16_rn(x);
  const bf16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
  return (uint32_t)*reinterpret_cast<const uint16_t*>(&hi) | ((uint32_t)*reinterpret_cast<const uint16_t*>(&lo) << 16);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

template <int SDT>
__global__ void __launch_bounds__(WT_THREADS, 1) wgrad_c1tc_kernel(const Wc1Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[2], empty[2], done_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ uint32_t lut[256];
  __shared__ float red[4][27][8];                  // per-warp partial sums over the warp's four plane slots
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* const gbuf = smem;                               // [2][W windows][16 slots][8 y][16 x][16 B]
  uint8_t* const tbuf = smem + 2 * (size_t)a.gbytes;        // [2][W windows][hi, lo][nch][WT_R][16 B]
  const uint32_t tmem_cols = 9 * a.N > 256 ? 512u : 256u;

  auto decode = [&](int it, int& b, int& y0, int& x0, int& z0) {
    int w = it;
    const int zr = w % a.nzr; w /= a.nzr;
    const int xb = w % a.nxb; w /= a.nxb;
    const int yb = w % a.nyb; w /= a.nyb;
    b = w; y0 = yb * WT_TY; x0 = xb * WT_TX; z0 = zr * a.W * a.nzw;
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&full[i], WT_NP); mbar_init(&empty[i], 1); }
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (SDT == DT_U8) {
    for (int i = threadIdx.x; i < 256; i += WT_THREADS) {
      lut[i] = wc_split_bf16(tem_standardize((float)i, a.lut_mean, a.lut_std));
    }
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ---- MMA issuer
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(a.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t g16 = smem_u32(gbuf) >> 4, t16 = smem_u32(tbuf) >> 4;
    const uint32_t a_hi = (uint32_t)((WT_TY * WT_TX * 16) >> 4) | (1u << 14);      // SBO = one plane slot
    const uint32_t b_hi = (uint32_t)((WT_R * 16) >> 4) | (1u << 14);               // SBO = one z chunk
    const uint32_t lbo = (128u >> 4) << 16;                                        // LBO = 8 voxels along x
    uint32_t ph = 0; int buf = 0; uint32_t acc = 0u;
    for (int it = blockIdx.x; it < a.items; it += gridDim.x) {
      int b, y0, x0, z0; decode(it, b, y0, x0, z0);
      const int ny = min(WT_TY, a.L[1] - y0);
      const int nwin = min(a.W, (a.L[0] - z0 + a.nzw - 1) / a.nzw);
      mbar_wait(&full[buf], (ph >> buf) & 1u); ph ^= 1u << buf;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        if (!(a.dbg & 16)) {
          const uint32_t gb = g16 + (uint32_t)buf * ((uint32_t)a.gbytes >> 4), tb = t16 + (uint32_t)buf * ((uint32_t)a.tbytes >> 4);
          for (int w = 0; w < nwin; ++w) {
            for (int y = 0; y < ny; ++y) {
              const uint32_t alo = (gb + (uint32_t)(w * (WT_WIN_BYTES >> 4) + y * (WT_TX * 16 >> 4))) | lbo;
              const uint32_t brow = tb + (uint32_t)(w * 2 * a.nch * WT_R + y * WT_HX);
#pragma unroll
              for (int t = 0; t < 9; ++t) {
                const uint32_t blo = (brow + (uint32_t)((t / 3) * WT_HX + (t % 3))) | lbo;
                umma_bf16(tmem_base + (uint32_t)(t * a.N), ((uint64_t)a_hi << 32) | alo, ((uint64_t)b_hi << 32) | blo, idesc, acc);
              }
              acc = 1u;
            }
          }
        }
        umma_commit(&empty[buf]);
      }
      __syncwarp();
      acc = 1u;
      buf ^= 1;
    }
    if (elect_one()) umma_commit(&done_bar);
    __syncwarp();
  } else {
    // ---- producers: P window(s) by cp.async, s tile gathered and converted to bf16
    const int ptid = threadIdx.x - 32;
    const SrcView& S = a.S;
    const long long szs = (long long)S.Y * S.X * S.C;
    const long long pzs = (long long)a.PY * a.PX * a.p_C;
    const int gx = ptid & 15, gy = (ptid >> 4) & 7, ghalf = ptid >> 7;     // this thread's (x, y) of the P tile, odd / even plane slots
    const int ncg = a.Cb >> 3;
    uint32_t ph = 0; int buf = 0;
    for (int it = blockIdx.x; it < a.items; it += gridDim.x) {
      int b, y0, x0, z0; decode(it, b, y0, x0, z0);
      mbar_wait(&empty[buf], ((ph >> buf) & 1u) ^ 1u); ph ^= 1u << buf;
      // P: slot (w, cg, jz) of window w <- plane z0 + w*nzw + jz, channels cg*8 .. +7, zero outside the extent
      if (!(a.dbg & 8)) {
        const bool xy_ok = (y0 + gy) < a.L[1] && (x0 + gx) < a.L[2];
        const bf16* const pcol = a.P + (long long)b * a.p_bstride + (((long long)a.p_off[0] * a.PY + y0 + gy + a.p_off[1]) * a.PX + x0 + gx + a.p_off[2]) * a.p_C + a.p_coff;
        const uint32_t gdst = smem_u32(gbuf + (size_t)buf * a.gbytes) + (uint32_t)((gy * WT_TX + gx) * 16);
        const int nslot = a.W * 16;
        for (int sl = ghalf; sl < nslot; sl += WT_NP / 128) {
          const int w = sl >> 4, s16 = sl & 15;
          const int cg = (ncg == 2) ? (s16 >> 3) : 0, jz = (ncg == 2) ? (s16 & 7) : s16;
          const int z = z0 + w * a.nzw + jz;
          const bool ok = xy_ok && z < a.L[0];
          cp_async16(gdst + (uint32_t)(sl * (WT_TY * WT_TX * 16)), ok ? (const void*)(pcol + (long long)z * pzs + cg * 8) : (const void*)a.P, ok ? 16u : 0u);
        }
      }
      // s: bf16 tile [z chunk][row = (hy, hx)][8 planes]
      const int zb = z0 + S.shift[0], yb = y0 + S.shift[1], xb = x0 + S.shift[2];
      const long long sbase = (long long)b * S.bstride;
      uint8_t* const dst = tbuf + (size_t)buf * a.tbytes;
      const int total = (a.dbg & 4) ? 0 : a.W * a.nch * WT_ROWS;
      const int lo_off = a.nch * WT_R * 16;
      for (int base = ptid; base < total; base += WT_NP * WT_PB) {
        uint32_t e[WT_PB][8];
#pragma unroll
        for (int k = 0; k < WT_PB; ++k) {
          const int idx = base + k * WT_NP;
          const int c = idx / WT_ROWS, r = idx - c * WT_ROWS;           // c = window * nch + chunk of the window
          const int hy = r / WT_HX, hx = r - hy * WT_HX;
          const int cw = c / a.nch;
          const int y = yb + hy, x = xb + hx, zc = zb + cw * a.nzw + (c - cw * a.nch) * 8;
          const bool ok = idx < total && y >= 0 && y < S.Y && x >= 0 && x < S.X;
          const long long off = sbase + (((long long)zc * S.Y + y) * S.X + x) * S.C + S.coff;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            e[k][u] = 0xFFFFFFFFu;                                     // marks "outside": virtual zero padding
            if (ok && zc + u >= 0 && zc + u < S.Z) {
              if (SDT == DT_U8) e[k][u] = reinterpret_cast<const uint8_t*>(S.p)[off + u * szs];
              else if (SDT == DT_BF16) e[k][u] = reinterpret_cast<const uint16_t*>(S.p)[off + u * szs];
              else e[k][u] = __float_as_uint(reinterpret_cast<const float*>(S.p)[off + u * szs]);
            }
          }
        }
#pragma unroll
        for (int k = 0; k < WT_PB; ++k) {
          const int idx = base + k * WT_NP;
          if (idx >= total) break;
          const int c = idx / WT_ROWS, r = idx - c * WT_ROWS;
          uint32_t h[8];                                              // hi | lo << 16
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            // 0xFFFFFFFF is a NaN pattern a finite gradient / image never holds
            if (SDT == DT_F32) h[u] = e[k][u] == 0xFFFFFFFFu ? 0u : wc_split_bf16(__uint_as_float(e[k][u]));
            else if (SDT == DT_U8) h[u] = e[k][u] == 0xFFFFFFFFu ? 0u : lut[e[k][u]];
            else h[u] = e[k][u] == 0xFFFFFFFFu ? 0u : e[k][u];
          }
          uint4 ph, pl;
          ph.x = __byte_perm(h[0], h[1], 0x5410); ph.y = __byte_perm(h[2], h[3], 0x5410); ph.z = __byte_perm(h[4], h[5], 0x5410); ph.w = __byte_perm(h[6], h[7], 0x5410);
          pl.x = __byte_perm(h[0], h[1], 0x7632); pl.y = __byte_perm(h[2], h[3], 0x7632); pl.z = __byte_perm(h[4], h[5], 0x7632); pl.w = __byte_perm(h[6], h[7], 0x7632);
          const int cw = c / a.nch;
          uint8_t* const d = dst + ((size_t)(c + cw * a.nch) * WT_R + r) * 16;       // slot (window, hi, chunk)
          *reinterpret_cast<uint4*>(d) = ph;
          *reinterpret_cast<uint4*>(d + lo_off) = pl;
        }
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes (cp.async, st.shared) -> the MMA's async-proxy reads
      mbar_arrive(&full[buf]);
      buf ^= 1;
    }
    // ---- final fold (first four producer warps): dw[dz][t][cb] = sum_jz D_t[(jz, cb)][jz + dz]
    if (ptid < 128) {
    mbar_wait(&done_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int slot = m >> 3;                                 // lane = slot * 8 + channel of the group
    const int jz = (ncg == 2) ? (slot & 7) : slot;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    // a warp holds four plane slots of one channel group (lane = slot_in_warp * 8 + c8): shuffles add them up (a shared-memory
    // float atomicAdd is a compare-and-swap loop: with 16 lanes per address it cost 35 us per launch)
    // the three columns a lane needs (jz, jz + 1, jz + 2) differ from lane to lane while tcgen05.ld reads the same columns for
    // the whole warp: every lane parks its row in shared memory (the tile buffers are free now; pitch 49: conflict-free) and
    // picks its diagonal from there.  (Selecting in registers compiled to a divergent jump table: 17 us per launch.)
    float* const scr = reinterpret_cast<float*>(smem) + m * 49;
    const int half = a.N >> 1;
    for (int t = 0; t < 9; ++t) {
      for (int c0 = 0; c0 < a.N; c0 += 16) {
        uint32_t r[16];
        wc_tmem_ld16(lane_base + (uint32_t)(t * a.N + c0), r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 16; ++c) scr[c0 + c] = __uint_as_float(r[c]);
      }
      float v[3];
#pragma unroll
      for (int dz = 0; dz < 3; ++dz) v[dz] = scr[jz + dz] + scr[half + jz + dz];     // hi + lo columns
#pragma unroll
      for (int dz = 0; dz < 3; ++dz) {
        float sum = v[dz];
        sum += __shfl_xor_sync(0xffffffffu, sum, 8);
        sum += __shfl_xor_sync(0xffffffffu, sum, 16);
        if (lane < 8) red[q][dz * 9 + t][lane] = sum;
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    for (int i = ptid; i < 27 * a.Cb; i += 128) {
      const int tap = i / a.Cb, c = i - tap * a.Cb;
      // Cb = 8: the four warps hold planes 0-3, 4-7, 8-11, 12-15; Cb = 16: warps 2 cg and 2 cg + 1 hold channel group cg
      const float s = (ncg == 2) ? red[2 * (c >> 3)][tap][c & 7] + red[2 * (c >> 3) + 1][tap][c & 7]
                                 : red[0][tap][c] + red[1][tap][c] + red[2][tap][c] + red[3][tap][c];
      if (s != 0.f && !(a.dbg & 2)) atomicAdd(a.dw + (long long)tap * a.ws_tap + (long long)c * a.ws_b, s);
    }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

bool plan(const WgradArgs& w, Wc1Args& t, size_t& smem) {
  memset(&t, 0, sizeof(t));
  t.Cb = w.Cb; t.nzw = 128 / w.Cb; t.nch = t.nzw / 8 + 1; t.N = 2 * t.nch * 8;
  t.B = w.B; for (int i = 0; i < 3; ++i) { t.L[i] = w.L[i]; t.p_off[i] = w.p_off[i]; }
  const int nwin_all = (w.L[0] + t.nzw - 1) / t.nzw;
  t.W = 2;                                                    // two windows (2 x 32 KB of P) per item
  if (t.W > nwin_all) t.W = nwin_all;
  t.gbytes = t.W * WT_WIN_BYTES; t.tbytes = t.W * 2 * t.nch * WT_R * 16;
  smem = 2 * (size_t)t.gbytes + 2 * (size_t)t.tbytes + 1024;
  return smem <= 208 * 1024;
}

}  // namespace

bool wgrad_c1tc_supported(const WgradArgs& w) {
  static const bool off = getenv("TEM_NO_WGRAD_C1TC") != nullptr;     // debug knob: CUDA-core kernels of wgrad_c1.cu
  if (off) return false;
  if (w.Ca != 1 || w.Ca1 || !(w.Cb == 8 || w.Cb == 16) || w.p_dtype != DT_BF16) return false;
  for (int i = 0; i < 3; ++i) if (w.k[i] != 3 || w.stride[i] != 1 || w.pad[i] != 0) return false;
  if (w.S.origins || (w.use_lut ? w.S.dtype != DT_U8 : w.S.dtype == DT_U8)) return false;
  if (w.p_C % 8 || w.p_coff % 8 || (reinterpret_cast<uintptr_t>(w.P) & 15)) return false;
  Wc1Args t; size_t smem;
  return plan(w, t, smem);
}

cudaError_t launch_wgrad_c1tc(const WgradArgs& w, cudaStream_t st) {
  Wc1Args t; size_t smem;
  if (!plan(w, t, smem)) return cudaErrorInvalidConfiguration;
  if ((long long)w.B * w.L[0] * w.L[1] * w.L[2] == 0) return cudaSuccess;
  t.S = w.S; t.use_lut = w.use_lut; t.lut_mean = w.lut_mean; t.lut_std = w.lut_std;
  t.P = (const bf16*)w.P; t.PZ = w.PZ; t.PY = w.PY; t.PX = w.PX; t.p_C = w.p_C; t.p_coff = w.p_coff; t.p_bstride = w.p_bstride;
  t.nyb = (w.L[1] + WT_TY - 1) / WT_TY; t.nxb = (w.L[2] + WT_TX - 1) / WT_TX;
  t.nzr = (w.L[0] + t.W * t.nzw - 1) / (t.W * t.nzw);
  t.items = w.B * t.nyb * t.nxb * t.nzr;
  t.dw = w.dw; t.ws_tap = w.ws_tap; t.ws_b = w.ws_b;
  t.dbg = tem_ablation_bits();
  const unsigned grid = (unsigned)(t.items < 148 ? t.items : 148);
  static bool attr[3] = {false, false, false};
#define LAUNCH_WC1(SDT, IDX)                                                                                                        \
  {                                                                                                                                 \
    if (!attr[IDX]) { cudaError_t e = cudaFuncSetAttribute(wgrad_c1tc_kernel<SDT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024); if (e) return e; attr[IDX] = true; } \
    wgrad_c1tc_kernel<SDT><<<grid, WT_THREADS, smem, st>>>(t);                                                                      \
  }
  if (w.S.dtype == DT_U8) LAUNCH_WC1(DT_U8, 0) else if (w.S.dtype == DT_BF16) LAUNCH_WC1(DT_BF16, 1) else LAUNCH_WC1(DT_F32, 2)
#undef LAUNCH_WC1
  ++g_tem_launches;
  return cudaGetLastError();
}
