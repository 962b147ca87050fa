// tcgen05 implicit-GEMM 3x3x3 stride-1 convolution for sm_100a (forward and, with flipped/transposed packed
// weights and OOB zero-fill as padding, the data gradient).  transfer_em/models/utils.py:73,122; generator.py:54-110.
//
// GEMM view:  D[M = 128 output voxels, N = Cout] += A[M, K = 27*Cin] * B[K, N]
//   * one CTA owns an (x-tile 8) x (y-tile 16) column of one sample and marches along z; every output z-slice is
//     one 128-row accumulator tile (TMEM lane r = y_local*8 + x_local), double-buffered in TMEM;
//   * the halo of each input z-slice ((16+2) x (8+2) voxels x Cin) is staged ONCE by TMA into a ring of shared
//     memory slots as channel planes [plane = 8 channels][y][x][8ch] (no-swizzle K-major "core matrices" are
//     8 voxels x 16 B contiguous, so a tap (dz,dy,dx) is just a shifted descriptor start address: slot(dz) +
//     (dy*HX+dx)*16 B, SBO = HX*16 B between y rows, LBO = plane stride between the two 8-channel K halves).
//     Nothing is re-read per tap: 27*Cin/16 MMAs are issued per output slice straight from the halo ring;
//   * fused crop-and-concat: the planes of a slot may come from two tensor maps (up-sampled tensor + cropped skip);
//   * weights are pre-packed (bf16) into the UMMA B-operand image [k-step][k-half][n-group][8][8] and loaded once
//     per CTA with one bulk copy;
//   * warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer, warps 2-5 = epilogue
//     (tcgen05.ld -> LeakyReLU / LeakyReLU' * dropout mask / accumulate -> bf16 NDHWC stores).
#include <cuda.h>
#include <string.h>
#include "tem_kernels.cuh"
#include "ptx_sm100.cuh"

extern unsigned long long g_tem_launches;

namespace {

constexpr int TX = 8, TY = 16, HX = TX + 2, HY = TY + 2;
constexpr int PLANE_BYTES = HY * HX * 16;          // 2880: one 8-channel plane of a halo slice
constexpr int PLANE_STRIDE = 2944;                 // padded to a multiple of 128 B (TMA destination alignment)
constexpr int RING_MAX = 12;
constexpr int kTcThreads = 192;

struct TcArgs {
  int B, L[3];                 // conv output extent (z,y,x)
  int planes0, planes1;        // 8-channel planes taken from map0 / map1
  int merged0, merged1;        // map folds (channel, x): see tem_make_map_c8
  int shift0[3], shift1[3];    // tensor coordinate = conv-input coordinate + shift (z,y,x)
  int spd;                     // k-steps (K=16 MMAs) per dz
  int cin8;                    // 1 when Cin == 8 (tap-pair k-steps)
  const bf16* wpacked; int wbytes;
  int ntx, nty, nzc, zc, ring;
  bf16* out; int OZ, OY, OX, out_C, out_coff, out_off[3];
  int Cout;
  float slope;
  const bf16* ref; int RZ, RY, RX, ref_C, ref_coff, ref_off[3]; float ref_slope;
  uint32_t drop_key;
  int accumulate;
};

template <int NPAD>
__global__ void __launch_bounds__(kTcThreads, (NPAD == 16) ? 4 : 3)
conv3_tc_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1, const TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[RING_MAX], empty_bar[RING_MAX], w_bar, tfull_bar[2], tempty_bar[2];
  const int RING = a.ring;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int planes = a.planes0 + a.planes1;
  const uint32_t wbytes_pad = (uint32_t)((a.wbytes + 1023) & ~1023);
  uint8_t* wsm = smem;
  uint8_t* ring = smem + wbytes_pad;
  const uint32_t slot_bytes = (uint32_t)planes * PLANE_STRIDE;

  if (threadIdx.x == 0) {
    for (int i = 0; i < RING; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&w_bar, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr uint32_t kTmemCols = (2 * NPAD < 32) ? 32 : 2 * NPAD;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  // work item
  int w = blockIdx.x;
  const int zc_i = w % a.nzc; w /= a.nzc;
  const int tx_i = w % a.ntx; w /= a.ntx;
  const int ty_i = w % a.nty; w /= a.nty;
  const int b = w;
  const int x0 = tx_i * TX, y0 = ty_i * TY, z0 = zc_i * a.zc;
  const int nz = min(a.zc, a.L[0] - z0);
  const int nslices = nz + 2;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&w_bar, (uint32_t)a.wbytes);
      bulk_load(wsm, a.wpacked, (uint32_t)a.wbytes, &w_bar);
      int slot = 0; uint32_t ph = 0;                       // ring position kept incrementally (no runtime division)
      for (int s = 0; s < nslices; ++s) {
        mbar_wait(&empty_bar[slot], ph ^ 1u);
        mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)planes * PLANE_BYTES);
        uint8_t* dst = ring + (size_t)slot * slot_bytes;
        for (int p = 0; p < a.planes0; ++p)
          tma_load_plane(dst + p * PLANE_STRIDE, &map0, &full_bar[slot], a.merged0, p, x0 + a.shift0[2], y0 + a.shift0[1], z0 + s + a.shift0[0], b);
        for (int p = 0; p < a.planes1; ++p)
          tma_load_plane(dst + (a.planes0 + p) * PLANE_STRIDE, &map1, &full_bar[slot], a.merged1, p, x0 + a.shift1[2], y0 + a.shift1[1], z0 + s + a.shift1[0], b);
        if (++slot == RING) { slot = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NPAD >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      mbar_wait(&w_bar, 0);
      const uint32_t wbase = smem_u32(wsm);
      const uint32_t rbase = smem_u32(ring);
      int waited = 0, wslot = 0, zslot = 0; uint32_t wph = 0;
      for (int zo = 0; zo < nz; ++zo) {
        while (waited < zo + 3) { mbar_wait(&full_bar[wslot], wph); ++waited; if (++wslot == RING) { wslot = 0; wph ^= 1u; } }
        mbar_wait(&tempty_bar[zo & 1], (((uint32_t)(zo >> 1)) & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(zo & 1) * NPAD;
        uint32_t acc = 0;
        int step = 0;
        for (int dz = 0; dz < 3; ++dz) {
          int sl = zslot + dz; if (sl >= RING) sl -= RING;
          const uint32_t sbase = rbase + (uint32_t)sl * slot_bytes;
          if (a.cin8) {
            // Cin == 8: one plane; a K=16 step covers two taps (LBO = distance between the taps), the 10th tap is a
            // zero-weight dummy
#pragma unroll
            for (int p = 0; p < 5; ++p) {
              const int t0 = 2 * p, t1 = (2 * p + 1 < 9) ? 2 * p + 1 : 2 * p;
              const uint32_t o0 = (uint32_t)((t0 / 3) * HX + (t0 % 3)) * 16u;
              const uint32_t o1 = (uint32_t)((t1 / 3) * HX + (t1 % 3)) * 16u;
              const uint32_t lbo = (t1 == t0) ? 0u : (o1 - o0);   // dummy half re-reads tap 8 (finite data x zero weights)
              umma_bf16(d_tmem, umma_desc(sbase + o0, lbo, HX * 16), umma_desc(wbase + (uint32_t)step * (NPAD * 32), NPAD * 16, 128), idesc, acc);
              acc = 1; ++step;
            }
          } else {
            const int kcs = planes >> 1;
            for (int t = 0; t < 9; ++t) {
              const uint32_t o = (uint32_t)((t / 3) * HX + (t % 3)) * 16u;
              for (int kc = 0; kc < kcs; ++kc) {
                umma_bf16(d_tmem, umma_desc(sbase + (uint32_t)(2 * kc) * PLANE_STRIDE + o, PLANE_STRIDE, HX * 16),
                          umma_desc(wbase + (uint32_t)step * (NPAD * 32), NPAD * 16, 128), idesc, acc);
                acc = 1; ++step;
              }
            }
          }
        }
        umma_commit(&tfull_bar[zo & 1]);     // accumulator ready for the epilogue
        umma_commit(&empty_bar[zslot]);      // input slice zo is no longer needed
        if (++zslot == RING) zslot = 0;
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int yl = row >> 3, xl = row & 7;
    const int oy = y0 + yl, ox = x0 + xl;
    const bool inside = oy < a.L[1] && ox < a.L[2];
    for (int zo = 0; zo < nz; ++zo) {
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
      const int oz = z0 + zo;
      float v[NPAD];
#pragma unroll
      for (int c = 0; c < NPAD; ++c) v[c] = __uint_as_float(r[c]);
      if (a.ref) {
        const long long ro = ((((long long)b * a.RZ + oz + a.ref_off[0]) * a.RY + oy + a.ref_off[1]) * a.RX + ox + a.ref_off[2]) * a.ref_C + a.ref_coff;
#pragma unroll
        for (int c = 0; c < NPAD; c += 8) {
          if (c < a.Cout) {
            float f[8];
            unpack8(*reinterpret_cast<const uint4*>(a.ref + ro + c), f);
#pragma unroll
            for (int u = 0; u < 8; ++u) v[c + u] *= (f[u] > 0.f) ? 1.f : a.ref_slope;
          }
        }
      }
      if (a.drop_key) {
        const uint32_t di = (uint32_t)(((((long long)b * a.L[0] + oz) * a.L[1] + oy) * a.L[2] + ox) * a.Cout);
#pragma unroll
        for (int c = 0; c < NPAD; ++c) v[c] *= 2.f * tem_keep(a.drop_key, di + c);
      }
      bf16* op = a.out + ((((long long)b * a.OZ + oz + a.out_off[0]) * a.OY + oy + a.out_off[1]) * a.OX + ox + a.out_off[2]) * a.out_C + a.out_coff;
#pragma unroll
      for (int c = 0; c < NPAD; c += 8) {
        if (c < a.Cout) {
          float o[8];
          if (a.accumulate) unpack8(*reinterpret_cast<const uint4*>(op + c), o);
          else {
#pragma unroll
            for (int u = 0; u < 8; ++u) o[u] = 0.f;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            o[u] += v[c + u];
            if (a.slope != 1.f) o[u] = o[u] > 0.f ? o[u] : o[u] * a.slope;
          }
          uint4 pk;
          pk.x = pack2(o[0], o[1]); pk.y = pack2(o[2], o[3]); pk.z = pack2(o[4], o[5]); pk.w = pack2(o[6], o[7]);
          *reinterpret_cast<uint4*>(op + c) = pk;
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

// packs fp32 master weights into the bf16 UMMA B image [k-step][k-half][n-group][8 rows][8 elems]
struct PackArgs {
  const float* w; long long ws_tap, ws_in, ws_out;
  int flip, cin, cols, npad, spd, cin8;
  bf16* dst; int total;
};
__global__ void pack_weights_kernel(const PackArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.total) return;
  int t = i;
  const int e = t & 7; t >>= 3;
  const int r = t & 7; t >>= 3;
  const int ng = a.npad >> 3;
  const int g = t % ng; t /= ng;
  const int j = t & 1; t >>= 1;
  const int s = t;
  const int dz = s / a.spd, within = s % a.spd;
  int tap9, ci;
  if (a.cin8) { tap9 = 2 * within + j; ci = e; }
  else { const int kcs = a.cin >> 4; tap9 = within / kcs; ci = (2 * (within % kcs) + j) * 8 + e; }
  const int co = g * 8 + r;
  float v = 0.f;
  if (tap9 < 9 && co < a.cols) {
    int tap = dz * 9 + tap9;
    if (a.flip) tap = 26 - tap;
    v = a.w[tap * a.ws_tap + (long long)ci * a.ws_in + (long long)co * a.ws_out];
  }
  a.dst[i] = __float2bfloat16_rn(v);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// host interface
// ------------------------------------------------------------------------------------------------
bool tc_conv_supported(const ConvArgs& a) {
  if (a.form != 0 && !(a.form == 1 && a.stride[0] == 1 && a.stride[1] == 1 && a.stride[2] == 1)) return false;
  if (a.k[0] != 3 || a.k[1] != 3 || a.k[2] != 3) return false;
  if (a.stride[0] != 1 || a.stride[1] != 1 || a.stride[2] != 1) return false;
  if (a.s0.dtype != DT_BF16 || a.out_dtype != DT_BF16) return false;
  if (a.s0.origins || a.use_lut || a.bias) return false;
  const int cin = a.C0 + a.C1;
  if (!(cin == 8 || cin % 16 == 0)) return false;
  if (a.C0 % 8 || a.C1 % 8 || a.s0.C % 8 || a.s0.coff != 0 || a.s0.C != a.C0) return false;
  if (a.C1 && (a.s1.dtype != DT_BF16 || a.s1.C % 8 || a.s1.coff != 0 || a.s1.C != a.C1)) return false;
  if (a.Cout % 8 || a.Cout > 32 || a.out_C % 8 || a.out_coff % 8) return false;
  if (a.ref && (a.ref_C % 8 || a.ref_coff % 8)) return false;
  const int steps = 3 * (cin == 8 ? 5 : 9 * (cin / 16));
  const int npad = a.Cout <= 16 ? 16 : 32;
  const size_t smem = (((size_t)steps * npad * 32 + 1023) & ~(size_t)1023) + (size_t)5 * (cin / 8) * PLANE_STRIDE + 1024;
  if (smem > 200 * 1024) return false;
  if (a.conv_off[0] || a.conv_off[1] || a.conv_off[2]) return false;
  return tem_get_encode() != nullptr;
}

size_t tc_packed_bytes(int cin, int cout) {
  const int steps = 3 * (cin == 8 ? 5 : 9 * (cin / 16));
  const int npad = cout <= 16 ? 16 : 32;
  return (size_t)steps * npad * 32;
}

// pack weights for a conv described by `a` (same weight addressing as the direct kernel) into dst
cudaError_t tc_pack_weights(const ConvArgs& a, bf16* dst, cudaStream_t st) {
  PackArgs p;
  const int cin = a.C0 + a.C1;
  p.w = a.w; p.ws_tap = a.ws_tap; p.ws_in = a.ws_in; p.ws_out = a.ws_out;
  p.flip = (a.form == 1) ? 1 : 0;      // data gradient = correlation with the flipped kernel
  p.cin = cin; p.cols = a.Cout; p.npad = a.Cout <= 16 ? 16 : 32;
  p.cin8 = cin == 8; p.spd = p.cin8 ? 5 : 9 * (cin / 16);
  p.dst = dst; p.total = (int)(tc_packed_bytes(cin, a.Cout) / 2);
  pack_weights_kernel<<<(p.total + 255) / 256, 256, 0, st>>>(p); ++g_tem_launches;
  return cudaGetLastError();
}

cudaError_t launch_conv_tc(const ConvArgs& a, const bf16* wpacked, cudaStream_t st) {
  TcArgs t; memset(&t, 0, sizeof(t));
  const int cin = a.C0 + a.C1;
  t.B = a.B; for (int i = 0; i < 3; ++i) t.L[i] = a.L[i];
  t.planes0 = a.C0 / 8; t.planes1 = a.C1 / 8;
  // form 1 (stride-1 data gradient): out[j] = sum_k in[j - k] w[k] = sum_k' in[j + k' - 2] w[2-k'] -> pad 2 via OOB zero fill
  const int pad = (a.form == 1) ? 2 : 0;
  for (int i = 0; i < 3; ++i) { t.shift0[i] = a.s0.shift[i] - pad; t.shift1[i] = a.s1.shift[i] - pad; }
  t.cin8 = cin == 8; t.spd = t.cin8 ? 5 : 9 * (cin / 16);
  t.wpacked = wpacked; t.wbytes = (int)tc_packed_bytes(cin, a.Cout);
  t.ntx = (a.L[2] + TX - 1) / TX; t.nty = (a.L[1] + TY - 1) / TY;
  // z chunking: enough CTAs to fill the machine, but at least ~6 slices per CTA to amortise the 2-slice halo + setup
  const long long cols = (long long)a.B * t.ntx * t.nty;
  int nzc = 1;
  while (cols * nzc < 2 * 148 && (a.L[0] + nzc) / (nzc + 1) >= 6) ++nzc;
  t.zc = (a.L[0] + nzc - 1) / nzc; t.nzc = (a.L[0] + t.zc - 1) / t.zc;
  t.out = (bf16*)a.out; t.OZ = a.OZ; t.OY = a.OY; t.OX = a.OX; t.out_C = a.out_C; t.out_coff = a.out_coff;
  for (int i = 0; i < 3; ++i) { t.out_off[i] = a.out_off[i]; t.ref_off[i] = a.ref_off[i]; }
  t.Cout = a.Cout; t.slope = a.slope;
  t.ref = a.ref; t.RZ = a.RZ; t.RY = a.RY; t.RX = a.RX; t.ref_C = a.ref_C; t.ref_coff = a.ref_coff; t.ref_slope = a.ref_slope;
  t.drop_key = a.drop_key; t.accumulate = a.accumulate;
  CUtensorMap m0, m1;
  if (!tem_make_map_plane(&m0, &t.merged0, a.s0.p, a.B, a.s0.Z, a.s0.Y, a.s0.X, a.s0.C, HX, HY)) return cudaErrorInvalidValue;
  if (a.C1) { if (!tem_make_map_plane(&m1, &t.merged1, a.s1.p, a.B, a.s1.Z, a.s1.Y, a.s1.X, a.s1.C, HX, HY)) return cudaErrorInvalidValue; }
  else { m1 = m0; t.merged1 = t.merged0; }
  const int npad = a.Cout <= 16 ? 16 : 32;
  // ring depth 5..12 within ~16 KB: measured, occupancy (CTAs per SM) hides TMA latency better than a deeper ring
  int ring = (int)((16 * 1024) / ((size_t)(cin / 8) * PLANE_STRIDE));
  if (ring > RING_MAX) ring = RING_MAX;
  if (ring < 5) ring = 5;
  t.ring = ring;
  const size_t smem = (((size_t)t.wbytes + 1023) & ~(size_t)1023) + (size_t)ring * (cin / 8) * PLANE_STRIDE + 1024;
  const unsigned grid = (unsigned)(cols * t.nzc);
  static bool attr16 = false, attr32 = false;
  if (npad == 16) {
    if (!attr16) { cudaError_t e = cudaFuncSetAttribute(conv3_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); if (e) return e; attr16 = true; }
    conv3_tc_kernel<16><<<grid, kTcThreads, smem, st>>>(m0, m1, t);
  } else {
    if (!attr32) { cudaError_t e = cudaFuncSetAttribute(conv3_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); if (e) return e; attr32 = true; }
    conv3_tc_kernel<32><<<grid, kTcThreads, smem, st>>>(m0, m1, t);
  }
  ++g_tem_launches;
  return cudaGetLastError();
}
