// tcgen05 forward / data-gradient kernel for the single-input-channel 3x3x3 stride-1 layers:
//   g0 / d0 forward (1 -> 8: generator.py:54, discriminator.py:39; uint8 patches through the standardise LUT, fp32 / bf16
//   fakes, per-tile origins of tiled inference) and the data gradient of the 16 -> 1 layer g11 (flipped taps, zero padding 2,
//   LeakyReLU' mask of a10).
//
// With one input channel the reduction of an implicit GEMM is only the 27 taps, and an im2col of them costs more
// instructions than the FMAs it replaces (conv_c1.cu runs at ~48 % of the packed-FMA pipe, 59 us for g0 at batch 8).  Here the dz taps are
// folded into a banded (Toeplitz) weight matrix instead, so that no im2col is needed:
//   D[row = (y, x)][(j, co)] = sum_{dy,dx} sum_{i < 16} in[8 w + i][y + dy][x + dx] * T_{dy,dx}[i][(j, co)],
//   T_{dy,dx}[i][(j, co)] = w[i - j][dy][dx][co] for 0 <= i - j <= 2, else 0          (j = 0..7: output z = 8 w + j)
// i.e. per window of 8 output planes along z one M128 x N(8*CO) x K16 MMA per (dy, dx): 9 MMAs.
//   * the input keeps fp32-like precision as a bf16 pair: in = hi + lo with hi = bf16(in), lo = bf16(in - hi), each tap is
//     two MMAs (hi and lo tile, same weights) -- the result equals the CUDA-core kernels' fp32 input x bf16 weights to 2^-17.
//   * A operand (K-major, no swizzle): the halo tile is staged as bf16 in the layout [z chunk of 8 planes][row][16 B] with
//     row = hy * 18 + hx: the 128 rows of an MMA are 128 consecutive rows (SBO = 128 B), the two K halves are two
//     z chunks (LBO = rows * 16 B), and the (dy, dx) tap is a row offset dy * 18 + dx of the start address.  M = 128 rows
//     = 7 y-rows x 18 columns of which the 16 first per row are outputs (the others are junk rows, never stored).
//   * B operand: the 9 Toeplitz matrices (bf16, 2-4 KB each), packed once per parameter version, resident in shared memory.
//   * D: two TMEM buffers of 128 columns (2 windows at CO = 8, 1 at CO = 16); four epilogue warps (thread = voxel (y, x))
//     drain a buffer -- LeakyReLU / LeakyReLU' mask / dropout, bf16 pack, one 16 B store per voxel and plane, 16 lanes
//     = 256 B contiguous -- while the MMAs of the next unit run.
//   * four producer warps gather the next item's halo (uint8 -> LUT -> bf16, or fp32 / bf16 -> bf16) into the second
//     tile buffer, 48 loads in flight per thread (the ~1-2 us L2 round trip is paid twice per item, not once per
//     chunk); generic-proxy writes are fenced (fence.proxy.async) before the mbarrier hand-over to the MMA warp.
// CTAs are persistent over work items (sample, y block of 7, x block of 16, z range of <= 8 windows), two per SM.
#include <cuda.h>
#include <string.h>
#include <stdlib.h>
#include "tem_kernels.cuh"
#include "ptx_sm100.cuh"

extern unsigned long long g_tem_launches;

namespace {

constexpr int CT_TX = 16, CT_HX = 18, CT_TY = 7, CT_HY = 9;
constexpr int CT_ROWS = CT_HY * CT_HX;       // 162 staged rows (hy, hx)
constexpr int CT_PB = 6;                     // chunks gathered per producer thread and batch (48 loads in flight)
constexpr int CT_R = 169;                    // row pitch of an x chunk: rows up to 127 + 38 are read; 169 * 16 B = 16 (mod 128): conflict-free chunk-major stores
constexpr int CT_THREADS = 288;              // MMA warp, 4 epilogue warps, 4 producer warps

struct C1tArgs {
  SrcView S; int pad; int use_lut; float lut_mean, lut_std;
  int B, L[3];
  int CO, N;                   // padded output channels (8 / 16), MMA N = 8 * CO
  int Cout;
  const bf16* wimg; int wbytes;
  int nyb, nxb, nzs, wps, nwin, items;
  int tbuf, lo_off;            // bytes of one tile buffer (hi + lo image), offset of the lo image
  bf16* out; int OZ, OY, OX, out_C, out_coff, out_off[3];
  float slope;
  const bf16* ref; int RZ, RY, RX, ref_C, ref_coff, ref_off[3]; float ref_slope;
  uint32_t drop_key;
  int dbg;                     // stage-ablation bits (-DTEM_ABLATION builds only): 1 no epilogue work, 4 no input gather, 16 no MMAs
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}

// x -> (hi | lo << 16), hi = bf16(x), lo = bf16(x - hi)
__device__ __forceinline__ uint32_t split_bf16(float x) {
  const bf16 hi = __float2bfloat16_rn(x);
  const bf16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
  return (uint32_t)*reinterpret_cast<const uint16_t*>(&hi) | ((uint32_t)*reinterpret_cast<const uint16_t*>(&lo) << 16);
}

// MASK: the epilogue applies the LeakyReLU' mask of `ref` and / or dropout (data gradients); false: plain forward epilogue
template <int SDT, int CO, bool MASK>
__global__ void __launch_bounds__(CT_THREADS, 2) conv_c1tc_kernel(const C1tArgs a) {
  constexpr int N = 8 * CO;                 // columns of one window
  constexpr int WPU = 128 / N;              // windows per TMEM buffer
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t w_bar, tin_full[2], tin_empty[2], tfull[2], tempty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ uint32_t lut[256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t wpad = (uint32_t)((a.wbytes + 1023) & ~1023);
  uint8_t* wsm = smem;
  uint8_t* tbuf = smem + wpad;

  auto decode = [&](int it, int& b, int& y0, int& x0, int& w0, int& w1) {
    int w = it;
    const int zs = w % a.nzs; w /= a.nzs;
    const int xb = w % a.nxb; w /= a.nxb;
    const int yb = w % a.nyb; w /= a.nyb;
    b = w; y0 = yb * CT_TY; x0 = xb * CT_TX; w0 = zs * a.wps; w1 = min(a.nwin, w0 + a.wps);
  };

  if (threadIdx.x == 0) {
    mbar_init(&w_bar, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tin_full[i], 4); mbar_init(&tin_empty[i], 1); mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (SDT == DT_U8) {
    for (int i = threadIdx.x; i < 256; i += CT_THREADS) {
      lut[i] = split_bf16(tem_standardize((float)i, a.lut_mean, a.lut_std));
    }
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ---- MMA issuer
    if (lane == 0) { mbar_arrive_expect_tx(&w_bar, (uint32_t)a.wbytes); bulk_load(wsm, a.wimg, (uint32_t)a.wbytes, &w_bar); }
    __syncwarp();
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    mbar_wait(&w_bar, 0);
    const uint32_t a_hi = (128u >> 4) | (1u << 14);                               // SBO = 128 B: consecutive rows
    const uint32_t a_lbo = (((uint32_t)CT_R * 16u) >> 4) << 16;                   // LBO = next x chunk
    const uint32_t b_hi = (128u >> 4) | (1u << 14);
    const uint32_t b_lbo = (((uint32_t)N * 16u) >> 4) << 16;
    const uint32_t wb16 = smem_u32(wsm) >> 4, tb16 = smem_u32(tbuf) >> 4;
    uint32_t tin_ph = 0, te_ph = 0;          // phase bits (bit i = buffer i)
    int tb = 0, buf = 0;
    for (int it = blockIdx.x; it < a.items; it += gridDim.x) {
      int b, y0, x0, w0, w1; decode(it, b, y0, x0, w0, w1);
      mbar_wait(&tin_full[tb], (tin_ph >> tb) & 1u); tin_ph ^= 1u << tb;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t t16 = tb16 + (uint32_t)tb * ((uint32_t)a.tbuf >> 4), lo16 = (uint32_t)a.lo_off >> 4;
      for (int wu = w0; wu < w1; wu += WPU) {
        mbar_wait(&tempty[buf], ((te_ph >> buf) & 1u) ^ 1u); te_ph ^= 1u << buf;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          for (int wi = 0; wi < WPU && wu + wi < w1 && !(a.dbg & 16); ++wi) {
            const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 128 + wi * N);
            const uint32_t arow = t16 + (uint32_t)((wu + wi - w0) * CT_R);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              const uint32_t alo = (arow + (uint32_t)((t / 3) * CT_HX + (t % 3))) | a_lbo;
              const uint32_t blo = (wb16 + (uint32_t)(t * (N * 32 / 16))) | b_lbo;
              umma_bf16(d_tmem, ((uint64_t)a_hi << 32) | alo, ((uint64_t)b_hi << 32) | blo, idesc, t > 0 ? 1u : 0u);
              umma_bf16(d_tmem, ((uint64_t)a_hi << 32) | (alo + lo16), ((uint64_t)b_hi << 32) | blo, idesc, 1u);
            }
          }
          umma_commit(&tfull[buf]);
        }
        __syncwarp();
        buf ^= 1;
      }
      if (elect_one()) umma_commit(&tin_empty[tb]);          // all MMAs reading this tile buffer have retired when it fires
      __syncwarp();
      tb ^= 1;
    }
  } else if (warp <= 4) {
    // ---- epilogue: thread = accumulator row = voxel (hy, hx) of the tile
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int hy = row / CT_HX, hx = row - hy * CT_HX;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const long long out_zs = (long long)a.OY * a.OX * a.out_C, ref_zs = (long long)a.RY * a.RX * a.ref_C;
    const uint32_t di_zs = (uint32_t)(a.L[1] * a.L[2] * a.Cout);
    uint32_t tf_ph = 0;
    int buf = 0;
    for (int it = blockIdx.x; it < a.items; it += gridDim.x) {
      int b, y0, x0, w0, w1; decode(it, b, y0, x0, w0, w1);
      const int oy = y0 + hy, ox = x0 + hx;
      const bool rowok = hy < CT_TY && hx < CT_TX && oy < a.L[1] && ox < a.L[2];
      bf16* const ocol = a.out + ((((long long)b * a.OZ + a.out_off[0]) * a.OY + oy + a.out_off[1]) * a.OX + ox + a.out_off[2]) * a.out_C + a.out_coff;
      const bf16* const rcol = a.ref ? a.ref + ((((long long)b * a.RZ + a.ref_off[0]) * a.RY + oy + a.ref_off[1]) * a.RX + ox + a.ref_off[2]) * a.ref_C + a.ref_coff : nullptr;
      const uint32_t di_col = (uint32_t)((((long long)b * a.L[0]) * a.L[1] + oy) * a.L[2] + ox) * (uint32_t)a.Cout;
      for (int wu = w0; wu < w1; wu += WPU) {
        mbar_wait(&tfull[buf], (tf_ph >> buf) & 1u); tf_ph ^= 1u << buf;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // 16 columns per step: 2 planes (CO = 8) or 1 (CO = 16).  The loop is deliberately NOT unrolled: unrolled (with the mask /
        // dropout variants inlined) the epilogue was 48 KB of code run by four warps per CTA and stalled on instruction
        // fetch (28 of 54 us on g0 with loads, MMAs, TMEM reads and stores all switched off)
        const int nstep = min(WPU, w1 - wu) * (N / 16);
#pragma unroll 1
        for (int s = 0; s < nstep; ++s) {
          uint32_t r[16];
          if (a.dbg & 32) {
#pragma unroll
            for (int c = 0; c < 16; ++c) r[c] = 0x3f800000u + lane + c;
          } else {
            tmem_ld16(lane_base + (uint32_t)(buf * 128 + s * 16), r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          }
          if (!rowok || (a.dbg & 1)) continue;
#pragma unroll
          for (int v = 0; v < 16 / CO; ++v) {
            const int oz = wu * 8 + s * (16 / CO) + v;
            if (oz >= a.L[0]) continue;
            float o[CO];
#pragma unroll
            for (int c = 0; c < CO; ++c) o[c] = __uint_as_float(r[v * CO + c]);
            if (MASK) {
              if (rcol) {
#pragma unroll
                for (int c = 0; c < CO; c += 8) {
                  if (c < a.Cout) {
                    float f[8]; unpack8(__ldg(reinterpret_cast<const uint4*>(rcol + (long long)oz * ref_zs + c)), f);
#pragma unroll
                    for (int u = 0; u < 8; ++u) o[c + u] *= (f[u] > 0.f) ? 1.f : a.ref_slope;
                  }
                }
              }
              if (a.drop_key) {
                const uint32_t di = di_col + (uint32_t)oz * di_zs;
#pragma unroll
                for (int c = 0; c < CO; ++c) o[c] *= 2.f * tem_keep(a.drop_key, di + c);
              }
            }
#pragma unroll
            for (int c = 0; c < CO; ++c) o[c] = o[c] > 0.f ? o[c] : o[c] * a.slope;      // slope = 1: identity
#pragma unroll
            for (int c = 0; c < CO; c += 8) {
              if (c < a.Cout) {
                uint4 pk; pk.x = pack2(o[c], o[c + 1]); pk.y = pack2(o[c + 2], o[c + 3]); pk.z = pack2(o[c + 4], o[c + 5]); pk.w = pack2(o[c + 6], o[c + 7]);
                if (!(a.dbg & 2) || pk.x == 0x12345678u) *reinterpret_cast<uint4*>(ocol + (long long)oz * out_zs + c) = pk;
              }
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[buf]);
        buf ^= 1;
      }
    }
  } else {
    // ---- producers: halo of the item -> bf16 tile [z chunk][row = (hy, hx)][8 planes]
    const int ptid = threadIdx.x - 5 * 32;
    const SrcView& S = a.S;
    const uint32_t fill = (SDT == DT_U8 && S.origins) ? lut[0] : 0u;   // as conv_c1.cu: outside a tiled volume reads raw 0
    const long long zstride = (long long)S.Y * S.X * S.C;
    uint32_t pe_ph = 0;
    int tb = 0;
    for (int it = blockIdx.x; it < a.items; it += gridDim.x) {
      int b, y0, x0, w0, w1; decode(it, b, y0, x0, w0, w1);
      const int nch = w1 - w0 + 1;
      int oz = 0, oy = 0, ox = 0; long long sbase = (long long)b * S.bstride;
      if (S.origins) { oz = S.origins[b * 3]; oy = S.origins[b * 3 + 1]; ox = S.origins[b * 3 + 2]; sbase = 0; }
      const int zb = w0 * 8 - a.pad + S.shift[0] + oz, yb = y0 - a.pad + S.shift[1] + oy, xb = x0 - a.pad + S.shift[2] + ox;
      mbar_wait(&tin_empty[tb], ((pe_ph >> tb) & 1u) ^ 1u); pe_ph ^= 1u << tb;
      uint8_t* const dst = tbuf + (size_t)tb * a.tbuf;
      const int total = (a.dbg & 4) ? 0 : nch * CT_ROWS;
      for (int base = ptid; base < total; base += 128 * CT_PB) {
        // all loads of a batch are issued before the first one is consumed
        uint32_t e[CT_PB][8];
#pragma unroll
        for (int k = 0; k < CT_PB; ++k) {
          const int idx = base + k * 128;
          const int c = idx / CT_ROWS, r = idx - c * CT_ROWS;        // consecutive threads: consecutive rows (x fastest) of one chunk
          const int hy = r / CT_HX, hx = r - hy * CT_HX;
          const int y = yb + hy, x = xb + hx, z0 = zb + c * 8;
          const bool ok = idx < total && y >= 0 && y < S.Y && x >= 0 && x < S.X;
          const long long off = sbase + (((long long)z0 * S.Y + y) * S.X + x) * S.C + S.coff;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            e[k][u] = 0xFFFFFFFFu;                                     // marks "outside"
            if (ok && z0 + u >= 0 && z0 + u < S.Z) {
              if (SDT == DT_U8) e[k][u] = reinterpret_cast<const uint8_t*>(S.p)[off + u * zstride];
              else if (SDT == DT_BF16) e[k][u] = reinterpret_cast<const uint16_t*>(S.p)[off + u * zstride];
              else e[k][u] = __float_as_uint(reinterpret_cast<const float*>(S.p)[off + u * zstride]);
            }
          }
        }
#pragma unroll
        for (int k = 0; k < CT_PB; ++k) {
          const int idx = base + k * 128;
          if (idx >= total) break;
          const int c = idx / CT_ROWS, r = idx - c * CT_ROWS;
          uint32_t h[8];                                              // hi | lo << 16
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            // 0xFFFFFFFF is a NaN pattern a finite gradient / image never holds
            if (SDT == DT_F32) h[u] = e[k][u] == 0xFFFFFFFFu ? fill : split_bf16(__uint_as_float(e[k][u]));
            else if (SDT == DT_U8) h[u] = e[k][u] == 0xFFFFFFFFu ? fill : lut[e[k][u]];
            else h[u] = e[k][u] == 0xFFFFFFFFu ? fill : e[k][u];
          }
          uint4 ph, pl;
          ph.x = __byte_perm(h[0], h[1], 0x5410); ph.y = __byte_perm(h[2], h[3], 0x5410); ph.z = __byte_perm(h[4], h[5], 0x5410); ph.w = __byte_perm(h[6], h[7], 0x5410);
          pl.x = __byte_perm(h[0], h[1], 0x7632); pl.y = __byte_perm(h[2], h[3], 0x7632); pl.z = __byte_perm(h[4], h[5], 0x7632); pl.w = __byte_perm(h[6], h[7], 0x7632);
          *reinterpret_cast<uint4*>(dst + ((size_t)c * CT_R + r) * 16) = ph;
          *reinterpret_cast<uint4*>(dst + a.lo_off + ((size_t)c * CT_R + r) * 16) = pl;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the MMA's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&tin_full[tb]);
      tb ^= 1;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

// Toeplitz B image: [tap (dy,dx)][k half][n group][8 rows][8 elems], n = j * CO + co, k = input plane i of the window
struct PackTArgs { const float* w; long long ws_tap, ws_out; int flip, CO, cout, total; bf16* dst; };
__global__ void pack_toeplitz_kernel(const PackTArgs a) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.total) return;
  int t = idx;
  const int e = t & 7; t >>= 3;
  const int r = t & 7; t >>= 3;
  const int ng = a.CO;                       // N / 8 = CO groups
  const int g = t % ng; t /= ng;
  const int kh = t & 1; t >>= 1;
  const int tap9 = t;                        // dy * 3 + dx
  const int n = g * 8 + r, i = kh * 8 + e;
  const int j = n / a.CO, co = n % a.CO;
  const int dz = i - j;
  float v = 0.f;
  if (dz >= 0 && dz < 3 && co < a.cout) {
    int tap = dz * 9 + tap9;
    if (a.flip) tap = 26 - tap;
    v = a.w[tap * a.ws_tap + (long long)co * a.ws_out];
  }
  a.dst[idx] = __float2bfloat16_rn(v);
}

int co_pad(int cout) { return cout <= 8 ? 8 : 16; }

}  // namespace

bool c1tc_supported(const ConvArgs& a) {
  static const bool off = getenv("TEM_NO_CONV_C1TC") != nullptr;      // debug knob: CUDA-core kernels of conv_c1.cu
  if (off) return false;
  if (a.C0 != 1 || a.C1 != 0 || a.bias || a.accumulate || a.split) return false;
  for (int i = 0; i < 3; ++i) if (a.k[i] != 3 || a.stride[i] != 1 || a.conv_off[i]) return false;
  if (a.form != 0 && a.form != 1) return false;
  if (!(a.Cout == 8 || a.Cout == 16)) return false;
  if (a.out_dtype != DT_BF16 || a.out_C % 8 || a.out_coff % 8) return false;
  if (a.ref && (a.ref_C % 8 || a.ref_coff % 8)) return false;
  if (a.use_lut ? a.s0.dtype != DT_U8 : a.s0.dtype == DT_U8) return false;
  if (a.st_out) return false;
  return true;
}

size_t c1tc_packed_bytes(const ConvArgs& a) { return (size_t)9 * (8 * co_pad(a.Cout)) * 32; }

cudaError_t c1tc_pack_weights(const ConvArgs& a, bf16* dst, cudaStream_t st) {
  PackTArgs p;
  p.w = a.w; p.ws_tap = a.ws_tap; p.ws_out = a.ws_out; p.flip = a.form == 1; p.CO = co_pad(a.Cout); p.cout = a.Cout;
  p.total = (int)(c1tc_packed_bytes(a) / 2); p.dst = dst;
  pack_toeplitz_kernel<<<(p.total + 255) / 256, 256, 0, st>>>(p); ++g_tem_launches;
  return cudaGetLastError();
}

cudaError_t launch_conv_c1tc(const ConvArgs& a, const bf16* wimg, cudaStream_t st) {
  C1tArgs t; memset(&t, 0, sizeof(t));
  t.S = a.s0; t.pad = a.form == 1 ? 2 : 0; t.use_lut = a.use_lut; t.lut_mean = a.lut_mean; t.lut_std = a.lut_std;
  t.B = a.B; for (int i = 0; i < 3; ++i) { t.L[i] = a.L[i]; t.out_off[i] = a.out_off[i]; t.ref_off[i] = a.ref_off[i]; }
  if ((long long)a.B * a.L[0] * a.L[1] * a.L[2] == 0) return cudaSuccess;
  t.CO = co_pad(a.Cout); t.N = 8 * t.CO; t.Cout = a.Cout;
  t.wimg = wimg; t.wbytes = (int)c1tc_packed_bytes(a);
  t.nyb = (a.L[1] + CT_TY - 1) / CT_TY; t.nxb = (a.L[2] + CT_TX - 1) / CT_TX;
  t.nwin = (a.L[0] + 7) / 8;
  // z ranges: at most wmax windows each (two tile buffers of wmax + 1 chunks, hi + lo, next to the weights in <= 110 KB:
  // two CTAs per SM); more, shorter ranges when the items would not fill two CTAs per SM evenly
  const size_t wpad = ((size_t)t.wbytes + 1023) & ~(size_t)1023;
  int wmax = (int)((110 * 1024 - 1024 - wpad) / (4 * (size_t)CT_R * 16)) - 1;
  if (wmax > t.nwin) wmax = t.nwin;
  if (wmax < 1) return cudaErrorInvalidConfiguration;
  int nzs = (t.nwin + wmax - 1) / wmax;
  const long long base_items = (long long)a.B * t.nyb * t.nxb;
  auto eff = [&](int n) { const long long items = base_items * n; const long long waves = (items + 295) / 296; return (double)items / (double)(waves * 296); };
  int best = nzs; double best_eff = eff(nzs);
  for (int n = nzs + 1; n <= t.nwin && n <= nzs + 3; ++n) {
    if ((t.nwin + n - 1) / n < 2) break;                      // every range re-stages one extra z chunk
    if (eff(n) > best_eff + 0.05) { best_eff = eff(n); best = n; }
  }
  t.nzs = best; t.wps = (t.nwin + t.nzs - 1) / t.nzs; t.nzs = (t.nwin + t.wps - 1) / t.wps;
  t.items = (int)(base_items * t.nzs);
  t.out = (bf16*)a.out; t.OZ = a.OZ; t.OY = a.OY; t.OX = a.OX; t.out_C = a.out_C; t.out_coff = a.out_coff;
  t.slope = a.slope;
  t.ref = a.ref; t.RZ = a.RZ; t.RY = a.RY; t.RX = a.RX; t.ref_C = a.ref_C; t.ref_coff = a.ref_coff; t.ref_slope = a.ref_slope;
  t.drop_key = a.drop_key;
  t.dbg = tem_ablation_bits();
  t.lo_off = (t.wps + 1) * CT_R * 16; t.tbuf = 2 * t.lo_off;
  const size_t smem = wpad + 2 * (size_t)t.tbuf + 1024;
  const unsigned grid = (unsigned)(t.items < 296 ? t.items : 296);
  static bool attr[12] = {false, false, false, false, false, false, false, false, false, false, false, false};
#define LAUNCH_C1T(SDT, COV, MK, IDX)                                                                                              \
  {                                                                                                                                \
    if (!attr[IDX]) { cudaError_t e = cudaFuncSetAttribute(conv_c1tc_kernel<SDT, COV, MK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024); if (e) return e; attr[IDX] = true; } \
    conv_c1tc_kernel<SDT, COV, MK><<<grid, CT_THREADS, smem, st>>>(t);                                                             \
  }
#define LAUNCH_C1T_DT(COV, MK, IDX)                                                                                                \
  { if (dt == DT_U8) LAUNCH_C1T(DT_U8, COV, MK, IDX) else if (dt == DT_BF16) LAUNCH_C1T(DT_BF16, COV, MK, IDX + 1) else LAUNCH_C1T(DT_F32, COV, MK, IDX + 2) }
  const int dt = a.s0.dtype;
  const bool mask = a.ref != nullptr || a.drop_key != 0;
  if (t.CO == 8) { if (mask) LAUNCH_C1T_DT(8, true, 0) else LAUNCH_C1T_DT(8, false, 3) }
  else { if (mask) LAUNCH_C1T_DT(16, true, 6) else LAUNCH_C1T_DT(16, false, 9) }
#undef LAUNCH_C1T_DT
#undef LAUNCH_C1T
  ++g_tem_launches;
  return cudaGetLastError();
}
