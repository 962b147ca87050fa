// tcgen05 weight gradient of the 3x3x3 stride-1 VALID convolutions (models/utils.py:73,122; generator.py:54-110):
//   dw[(dz,dy,dx)][ca][cb] = sum_{b,v} x[b, v + (dz,dy,dx)][ca] * g[b, v][cb]            
//
// GEMM view: D[M][N] += A[M][K] * B[N][K]^T with K = 16 consecutive x-voxels of one (z, y) row.  Both operands are
// read in their natural channels-last layout as MN-MAJOR UMMA operands: a core matrix is 8 voxels (K) x 8 channels
// (16 B, MN), i.e. exactly the [voxel][8 channels] plane layout TMA writes.  The taps are carried by the 8-channel
// GROUPS of the M and N dimensions, which only need ONE uniform stride each:
//   * M groups = (channel plane, row) of the x tile: group stride = one row of the tile.  With R_A rows per plane the
//     MMA sees x rows y0 .. y0+R_A-1 at once;
//   * N groups = (channel plane, row) of the g tile: rows y0 .. y0+RB-1.
//   Block (x row i, g row j) of D is the partial sum of tap dy = i - j over the voxels of g row j: the three
//   diagonals dy = 0, 1, 2 are useful, everything else is discarded.  Different j are different voxels, so nothing is
//   computed twice; the diagonal blocks are added up in the epilogue.
//   * dx is a shifted start address of the g operand (16 B per voxel), dz a different g slice: 9 (dz,dx) accumulators
//     of N columns each stay resident in TMEM for the whole CTA.
// One CTA owns (sample, RB g-rows, a chunk of x z-slices, the full x extent); per x z-slice the x tile and the newest
// g slice arrive through TMA rings (OOB zero fill gives the virtual zero padding of g), one thread issues
// (runs x 9) M64/128 x N x K16 MMAs, and after the last slice four warps fold the diagonals of the 9 accumulators
// into a shared-memory dw image that leaves with one atomicAdd per weight and CTA.
// wgrad_mma.cu explains why taps cannot be the M dimension directly; this kernel side-steps that by letting rows of
// the tile (uniformly strided) play the role of the dy taps.
#include <cuda.h>
#include <string.h>
#include <stdlib.h>
#include "tem_kernels.cuh"
#include "ptx_sm100.cuh"

extern unsigned long long g_tem_launches;

namespace {

constexpr int XR_MAX = 8;                    // pipeline slots (chosen per launch)
constexpr int kThreads = 192;

struct WtArgs {
  int B, L[3];                 // extent of g (z,y,x)
  int pa, pb;                  // 8-channel planes of x / g
  int pa0;                     // x planes [0, pa0) come from mapx, [pa0, pa) from mapx1 (two-source input of a concat layer)
  int shift1[3];               // window shift of the second x tensor
  int RA, RB;                  // rows per plane in the x tile / in the g tile (RB = useful g rows per CTA)
  int M, N;                    // MMA shape: M = 8*pa*RA (64 or 128), N = 8*pb*RB
  int NR;                      // 16-voxel runs along x
  int WA, WB;                  // tile widths in voxels: WA = 16*NR, WB = 16*NR + 8 (g tile starts at x = -2)
  int shift[3];                // x tensor coordinate = x-window coordinate + shift
  int nrg, nzc, zc;            // row groups, z chunks, x slices per chunk
  int units;                   // work units (sample, row group, z chunk); CTAs are persistent over them
  int XR, sb;                  // pipeline slots, x slices per pipeline step
  int tmem_cols;
  int xa_bytes, gb_bytes;      // bytes of one ring slot
  float* dw; long long ws_tap, ws_a, ws_b;
  int vec4;                    // dw rows are contiguous in cb and 16 B aligned: red.global.add.v4.f32
  int dbg;                     // stage-ablation bits (-DTEM_ABLATION builds only): 1 no epilogue, 2 no atomics, 4 no x loads, 8 no g loads, 16 no MMAs
};

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapx, const __grid_constant__ CUtensorMap mapx1, const __grid_constant__ CUtensorMap mapg, const WtArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[XR_MAX], empty[XR_MAX], done_bar;
  const int NSLOT = a.XR, SB = a.sb;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // One pipeline step = SB consecutive x z-slices of a work unit plus the SB + 2 g slices they correlate with (the two
  // g slices shared with the next step are simply loaded again: they come from L2).  Measured (tools/ubench/sync_latency.cu):
  // a full / empty hand-over costs the issuing warp ~300-400 cycles whatever the ring depth, the 18 MMAs of one g7 slice
  // ~500: with one hand-over per slice the issue loop bounded the small layers (stage ablation: 23 of 31 us remained on g7
  // with loads and MMAs switched off).  Layout of a slot: [SB x-slices][SB + 2 g-slices].
  const uint32_t slot_bytes = (uint32_t)(SB * a.xa_bytes + (SB + 2) * a.gb_bytes);
  const uint32_t g_off = (uint32_t)(SB * a.xa_bytes);

  // work unit u -> (sample b, first g row y0, first x slice zx0, slices nzx)
  auto decode = [&](int u, int& b, int& y0, int& zx0, int& nzx) {
    const int zc_i = u % a.nzc; u /= a.nzc;
    const int rg = u % a.nrg; u /= a.nrg;
    b = u; y0 = rg * a.RB; zx0 = zc_i * a.zc;
    nzx = min(a.zc, a.L[0] + 2 - zx0);
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSLOT; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
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
      int slot = 0; uint32_t ph = 0;
      const int xplane = a.RA * a.WA * 16, gplane = a.RB * a.WB * 16;
      for (int u = blockIdx.x; u < a.units; u += gridDim.x) {
        int b, y0, zx0, nzx; decode(u, b, y0, zx0, nzx);
        for (int s0 = 0; s0 < nzx; s0 += SB) {
          const int n = min(SB, nzx - s0);
          mbar_wait(&empty[slot], ph ^ 1u);
          uint8_t* base = smem + (size_t)slot * slot_bytes;
          const int ng = (a.dbg & 8) ? 0 : n + 2, nx = (a.dbg & 4) ? 0 : n;
          if (ng + nx == 0) mbar_arrive(&full[slot]);
          else mbar_arrive_expect_tx(&full[slot], (uint32_t)(nx * a.xa_bytes + ng * a.gb_bytes));
          // g slice i of the step is z = zx0 + s0 + i - 2 (OOB rows / slices are zero filled: the padding of g)
          for (int i = 0; i < ng; ++i)
            for (int p = 0; p < a.pb; ++p)
              tma_load_5d(base + g_off + i * a.gb_bytes + p * gplane, &mapg, &full[slot], p * 8, -2, y0, zx0 + s0 + i - 2, b);
          for (int i = 0; i < nx; ++i)
            for (int p = 0; p < a.pa; ++p) {
              if (p < a.pa0) tma_load_5d(base + i * a.xa_bytes + p * xplane, &mapx, &full[slot], p * 8, a.shift[2], y0 + a.shift[1], zx0 + s0 + i + a.shift[0], b);
              else tma_load_5d(base + i * a.xa_bytes + p * xplane, &mapx1, &full[slot], (p - a.pa0) * 8, a.shift1[2], y0 + a.shift1[1], zx0 + s0 + i + a.shift1[0], b);
            }
          if (++slot == NSLOT) { slot = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // the whole warp runs the issue loop convergently (uniform values stay in uniform registers); one elected lane
    // executes the MMAs / commits
    {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)((3 * a.N) >> 3) << 17) | ((uint32_t)(a.M >> 4) << 24);      // N = the g rows of three consecutive g slices (dz = 2, 1, 0)
      // descriptors are affine in the start address: only the low word (start >> 4, LBO) changes inside the loops, so
      // the single issuing thread spends a couple of integer adds per MMA instead of rebuilding 64-bit descriptors
      const uint32_t lo_fixed = (128u >> 4) << 16;                                   // LBO = 128 B (next 8 voxels)
      const uint32_t a_hi = (((uint32_t)a.WA * 16u) >> 4) | (1u << 14), b_hi = (((uint32_t)a.WB * 16u) >> 4) | (1u << 14);   // SBO, version
      const uint32_t sbase16 = smem_u32(smem) >> 4;
      const uint32_t xa16 = (uint32_t)a.xa_bytes >> 4, gb16 = (uint32_t)a.gb_bytes >> 4, slot16 = slot_bytes >> 4, goff16 = g_off >> 4;
      const uint32_t N = (uint32_t)a.N;
      int slot = 0; uint32_t ph = 0;
      uint32_t acc = 0u;
      for (int u = blockIdx.x; u < a.units; u += gridDim.x) {
        int b, y0, zx0, nzx; decode(u, b, y0, zx0, nzx);
        for (int s0 = 0; s0 < nzx; s0 += SB) {
          const int n = min(SB, nzx - s0);
          mbar_wait(&full[slot], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one()) {
            const uint32_t xb = sbase16 + (uint32_t)slot * slot16, gb = xb + goff16 + 2u;     // +2 voxels: the g tile starts at x = -2
            if (!(a.dbg & 16)) {
              for (int i = 0; i < n; ++i) {
                // x slice i correlates with the g slices i, i+1, i+2 of the step (dz = 2, 1, 0: zd = zx - dz).  The three slices
                // are adjacent in the slot and their rows continue each other's row stride, so ONE MMA takes all three as
                // N = 3 * RB * Cb columns [dz = 2 | dz = 1 | dz = 0]: a third of the MMAs, three times as wide (the M = 64 x N = 48
                // MMAs of g1 kept the tensor pipe busy 53 % of the time for 24 % of its throughput)
                const uint32_t bg = (gb + (uint32_t)i * gb16) | lo_fixed;
                uint32_t alo = (xb + (uint32_t)i * xa16) | lo_fixed;
                for (int r = 0; r < a.NR; ++r) {
                  const uint32_t ro = (uint32_t)r * 16u;           // 16 voxels = 256 B = 16 descriptor units
                  const uint64_t ad = ((uint64_t)a_hi << 32) | alo;
#define WT_MMA(DX, BLO) umma_bf16(tmem_base + (DX) * 3u * N, ad, ((uint64_t)b_hi << 32) | (BLO), idesc, acc)
                  WT_MMA(0, bg + ro); WT_MMA(1, bg + ro - 1u); WT_MMA(2, bg + ro - 2u);
#undef WT_MMA
                  alo += 16u; acc = 1u;
                }
              }
            }
            umma_commit(&empty[slot]);
          }
          __syncwarp();
          acc = 1u;
          if (++slot == NSLOT) { slot = 0; ph ^= 1u; }
        }
      }
      if (elect_one()) umma_commit(&done_bar);
      __syncwarp();
    }
  }
  // ---- epilogue: fold the useful diagonals of the 9 accumulators into a shared dw image, then one vector reduction per
  // four weights and CTA.  The image rows ([tap][ca] x Cb floats) are padded by 4 floats: the eight lanes of a quarter
  // warp hold eight consecutive ca, i.e. rows 144 B apart (Cb = 32), and their float4 read-modify-writes fall into
  // eight different 16 B bank groups (unpadded: all into the same one, 1.8 M conflicts on g7 in the round-1 capture).
  float* red = reinterpret_cast<float*>(smem);
  const int Ca = a.pa * 8, Cb = a.pb * 8;
  const int CbP = Cb + 4;
  const int nrow = 27 * Ca;
  if (warp >= 2 && !(a.dbg & 1)) {
    mbar_wait(&done_bar, 0);                                 // all MMAs retired: rings are free, accumulators final
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int tid = threadIdx.x - 64;
    for (int i = tid; i < nrow * CbP / 4; i += 128) reinterpret_cast<float4*>(red)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int q = warp & 3;
    // M = 128: TMEM lane = row.  M = 64: rows 16q .. 16q+15 live in lanes 0..15 of subpartition q.
    const int m = (a.M == 128) ? q * 32 + lane : q * 16 + (lane & 15);
    const bool rowok = (a.M == 128) || lane < 16;
    const int gm = m >> 3, pA = gm / a.RA, gi = gm % a.RA, ca = pA * 8 + (m & 7);
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    // Round j folds the column groups of g row j: in one round an address (tap, ca, cb) is touched by exactly one
    // thread (x row i <-> dy = i - j), so plain read-modify-writes suffice; rounds are separated by a named barrier.
    for (int j = 0; j < a.RB; ++j) {
      const int ty = gi - j;
      const bool use = rowok && ty >= 0 && ty < 3;
      for (int acc = 0; acc < 9; ++acc) {
        const int dz = acc / 3, dx = acc % 3;
        float* rowp = red + ((size_t)((dz * 3 + ty) * 3 + dx) * Ca + ca) * CbP;
        for (int pB = 0; pB < a.pb; pB += 2) {              // two 8-column loads in flight per wait
          uint32_t r[16];
          const uint32_t cbase = (uint32_t)((dx * 3 + (2 - dz)) * a.N);          // accumulator dx, column block of g slice 2 - dz
          tmem_ld8(lane_base + cbase + (uint32_t)((pB * a.RB + j) * 8), r);
          if (pB + 1 < a.pb) tmem_ld8(lane_base + cbase + (uint32_t)(((pB + 1) * a.RB + j) * 8), r + 8);
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
    // every CTA starts its pass over the dw image at a different offset (fewer same-address collisions in L2); Ca and Cb
    // are powers of two (8 / 16 / 32): shifts instead of divisions
    const int lg_cb4 = (Cb == 8) ? 1 : (Cb == 16 ? 2 : 3), lg_ca = (Ca == 8) ? 3 : (Ca == 16 ? 4 : 5);
    const int n4 = nrow << lg_cb4;
    const int rot = (int)(((long long)blockIdx.x * n4 / gridDim.x) & ~127LL);
    if (!(a.dbg & 2)) {
      for (int i0 = tid; i0 < n4; i0 += 128) {
        int i = i0 + rot; if (i >= n4) i -= n4;
        const int row = i >> lg_cb4, c4 = i & ((1 << lg_cb4) - 1);
        const float4 v = *reinterpret_cast<const float4*>(red + (size_t)row * CbP + c4 * 4);
        if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) {
          const int tap = row >> lg_ca, cA = row & (Ca - 1);
          float* dst = a.dw + (long long)tap * a.ws_tap + (long long)cA * a.ws_a + (long long)(c4 * 4) * a.ws_b;
          if (a.vec4) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
          } else {
            atomicAdd(dst, v.x); atomicAdd(dst + a.ws_b, v.y); atomicAdd(dst + 2 * a.ws_b, v.z); atomicAdd(dst + 3 * a.ws_b, v.w);
          }
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

bool plan(const WgradArgs& w, WtArgs& t, size_t& smem) {
  memset(&t, 0, sizeof(t));
  t.B = w.B; for (int i = 0; i < 3; ++i) { t.L[i] = w.L[i]; t.shift[i] = w.S.shift[i]; }
  t.pa = w.Ca / 8; t.pb = w.Cb / 8; t.pa0 = (w.Ca - w.Ca1) / 8;
  for (int i = 0; i < 3; ++i) t.shift1[i] = w.S1.shift[i];
  t.M = (t.pa == 4 || (w.Ca1 && t.pa == 2 && w.Cb <= 16)) ? 128 : 64;     // two 8-channel sources: 8 x rows per plane, three g rows
  t.RA = (t.M / 8) / t.pa;
  int rb = t.RA - 2;
  static const char* nmax_s = getenv("TEM_WTC_NMAX");     // debug knob: 28 keeps the 9 accumulators within 256 TMEM columns
  const int nmax = nmax_s ? atoi(nmax_s) : 56;
  while (rb > 1 && rb * w.Cb > nmax) --rb;
  if (t.M == 128) while (rb > 1 && (rb * w.Cb) % 16) --rb;
  t.RB = rb; t.N = rb * w.Cb;
  if (t.N > 56 || (t.M == 128 && t.N % 16)) return false;   // Cb = 32 keeps N = 32 even under a smaller cap
  t.NR = (w.L[2] + 2 + 15) / 16;
  t.WA = 16 * t.NR; t.WB = 16 * t.NR + 8;
  if (t.WB > 256) return false;
  t.xa_bytes = t.pa * t.RA * t.WA * 16; t.gb_bytes = t.pb * t.RB * t.WB * 16;
  int cols = 32; while (cols < 9 * t.N) cols <<= 1;
  t.tmem_cols = cols;
  // pipeline step: sb x-slices with >= ~64 MMAs between two hand-overs (at least 2: a step re-loads two g slices);
  // as many slots as ~150 KB allow (one CTA per SM anyway: the accumulators take most of TMEM), at least two
  static const char* sb_s = getenv("TEM_WTC_SB");        // debug knob
  // (measured, profiles/negative_results_r2.md: the loads run at ~2.7 TB/s from L2 whatever the request granularity, so what counts is
  // the (sb + 2) / sb re-read factor of g: four slices per step took g1 from 64.7 to 58.2 us, g10 from 49.9 to 41.5)
  int sb = sb_s ? atoi(sb_s) : (64 + 9 * t.NR - 1) / (9 * t.NR);
  if (!sb_s && sb < 4) sb = 4;
  if (sb < 2) sb = 2; if (sb > 6) sb = 6;
  auto slot_of = [&](int n) { return (size_t)n * t.xa_bytes + (size_t)(n + 2) * t.gb_bytes; };
  while (sb > 1 && 2 * slot_of(sb) > 180 * 1024) --sb;
  t.sb = sb;
  t.XR = 2;
  while (t.XR < XR_MAX && (size_t)(t.XR + 1) * slot_of(sb) <= 150 * 1024) ++t.XR;
  smem = (size_t)t.XR * slot_of(sb) + 1024;
  const size_t red = (size_t)27 * w.Ca * (w.Cb + 4) * 4;
  if (red + 1024 > smem) smem = red + 1024;
  return smem <= 200 * 1024;
}

}  // namespace

bool wgrad_tc_supported(const WgradArgs& w) {
  if (w.S.dtype != DT_BF16 || w.p_dtype != DT_BF16 || w.S.origins || w.use_lut) return false;
  for (int i = 0; i < 3; ++i) if (w.k[i] != 3 || w.stride[i] != 1 || w.pad[i] != 0 || w.p_off[i] != 0) return false;
  if (!(w.Ca == 8 || w.Ca == 16 || w.Ca == 32) || !(w.Cb == 8 || w.Cb == 16 || w.Cb == 32)) return false;
  if (w.S.C != w.Ca - w.Ca1 || w.S.coff != 0 || w.p_C != w.Cb || w.p_coff != 0) return false;
  if (w.Ca1) {
    if (w.Ca1 % 8 || (w.Ca - w.Ca1) % 8 || w.Ca1 >= w.Ca || w.S1.dtype != DT_BF16 || w.S1.origins || w.S1.C != w.Ca1 || w.S1.coff != 0) return false;
    if (w.S1.bstride != (long long)w.S1.Z * w.S1.Y * w.S1.X * w.S1.C) return false;
  }
  if (w.PZ != w.L[0] || w.PY != w.L[1] || w.PX != w.L[2]) return false;      // OOB zero fill is the padding of g
  if (w.p_bstride != (long long)w.L[0] * w.L[1] * w.L[2] * w.Cb) return false;
  if (w.S.bstride != (long long)w.S.Z * w.S.Y * w.S.X * w.S.C) return false;
  WtArgs t; size_t smem;
  if (!plan(w, t, smem)) return false;
  return tem_get_encode() != nullptr;
}

cudaError_t launch_wgrad_tc(const WgradArgs& w, cudaStream_t st) {
  WtArgs t; size_t smem;
  if (!plan(w, t, smem)) return cudaErrorInvalidConfiguration;
  if ((long long)w.B * w.L[0] * w.L[1] * w.L[2] == 0) return cudaSuccess;
  t.dbg = tem_ablation_bits();
  t.dw = w.dw; t.ws_tap = w.ws_tap; t.ws_a = w.ws_a; t.ws_b = w.ws_b;
  t.vec4 = w.ws_b == 1 && (reinterpret_cast<uintptr_t>(w.dw) & 15) == 0 && (w.ws_tap & 3) == 0 && (w.ws_a & 3) == 0;
  t.nrg = (w.L[1] + t.RB - 1) / t.RB;
  // z chunks: one CTA per SM (the accumulators take most of TMEM); pick the chunk count with the best wave efficiency
  const long long cols = (long long)w.B * t.nrg;
  const int nslices = w.L[0] + 2;
  int best = 1; double best_eff = -1.0;
  for (int nzc = 1; nzc <= nslices && nzc <= 64; ++nzc) {
    const int zc = (nslices + nzc - 1) / nzc;
    if (zc < 3 && nzc > 1) break;
    const int real = (nslices + zc - 1) / zc;
    const long long units = cols * real;
    const long long waves = (units + 147) / 148;
    const double eff = (double)units / (double)(waves * 148) * (double)zc / (double)(zc + 2);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = nzc; }
  }
  t.zc = (nslices + best - 1) / best; t.nzc = (nslices + t.zc - 1) / t.zc;
  CUtensorMap mx, mx1, mg;
  if (!tem_make_map_5d(&mx, w.S.p, w.B, w.S.Z, w.S.Y, w.S.X, w.S.C, t.WA, t.RA)) return cudaErrorInvalidValue;
  if (w.Ca1) { if (!tem_make_map_5d(&mx1, w.S1.p, w.B, w.S1.Z, w.S1.Y, w.S1.X, w.S1.C, t.WA, t.RA)) return cudaErrorInvalidValue; }
  else mx1 = mx;
  if (!tem_make_map_5d(&mg, w.P, w.B, w.PZ, w.PY, w.PX, w.p_C, t.WB, t.RB)) return cudaErrorInvalidValue;
  static bool attr = false;
  if (!attr) { cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); if (e) return e; attr = true; }
  t.units = (int)(cols * t.nzc);
  // every CTA ends with a pass of 27*Ca*Cb atomics: small layers get fewer, longer-lived CTAs (>= ~512 MMAs each)
  static const char* mpc_s = getenv("TEM_WGRAD_TC_MPC");
  const long long mmas = (long long)t.units * t.zc * t.NR * 9;
  long long want = mmas / (mpc_s ? atoi(mpc_s) : 96);
  if (want < 1) want = 1; if (want > 148) want = 148; if (want > t.units) want = t.units;
  const unsigned grid = (unsigned)want;
  wgrad_tc_kernel<<<grid, kThreads, smem, st>>>(mx, mx1, mg, t); ++g_tem_launches;
  return cudaGetLastError();
}
