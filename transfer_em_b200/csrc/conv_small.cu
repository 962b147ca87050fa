// Strided 4x4x4 convolution forward for SMALL output volumes with a long reduction (discriminator.py:66-74: d4 maps
// 14^3 x 32 -> 6^3 x 32, d6 maps 4^3 x 32 -> 1 x 32; K = 64 taps x Cin = 2048 per output).  The tile shapes of the
// tensor-core kernels (128 output voxels per MMA, 128 KB of bf16 weights beside a TMA ring) do not fit these layers;
// they are tiny GEMMs with M = a few hundred voxels, so a CUDA-core kernel that reuses each weight for 8 outputs wins:
//   one CTA = one sample x a 2x2x2 block of output voxels x all output channels (lane = channel),
//   the 6^3 x Cin input block sits in shared memory (broadcast reads), the 64 taps are split over the 8 warps,
//   weights are read coalesced from the fp32 masters (L2 resident) and rounded to bf16 like every other conv here,
//   the 8 partial sums per (voxel, channel) are combined in shared memory, then LeakyReLU -> bf16.
#include <string.h>
#include "tem_kernels.cuh"

extern unsigned long long g_tem_launches;

namespace {

constexpr int kWarpsS = 8;
constexpr int IB = 6;                      // input block edge: 2*2 + 4 - 2

__global__ void __launch_bounds__(kWarpsS * 32) conv_small_s2_kernel(const ConvArgs a, const int nbz, const int nby, const int nbx) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int Cin = a.C0;
  bf16* xin = reinterpret_cast<bf16*>(smem_raw);                                   // [6][6][6][Cin]
  float* red = reinterpret_cast<float*>(smem_raw + (size_t)IB * IB * IB * Cin * 2);   // [8 warps][8 voxels][32]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int t = blockIdx.x;
  const int bx = t % nbx; t /= nbx;
  const int by = t % nby; t /= nby;
  const int bz = t % nbz; t /= nbz;
  const int b = t;
  const int oz0 = bz * 2, oy0 = by * 2, ox0 = bx * 2;
  // stage the input block (16 B chunks), zero outside the tensor
  const bf16* src = reinterpret_cast<const bf16*>(a.s0.p) + (long long)b * a.s0.bstride;
  const int chunks = Cin >> 3;
  for (int i = tid; i < IB * IB * IB * chunks; i += kWarpsS * 32) {
    const int c = i % chunks; int v = i / chunks;
    const int x = v % IB; v /= IB; const int y = v % IB; const int z = v / IB;
    const int gz = oz0 * 2 + z + a.s0.shift[0], gy = oy0 * 2 + y + a.s0.shift[1], gx = ox0 * 2 + x + a.s0.shift[2];
    uint4 q = make_uint4(0, 0, 0, 0);
    if (gz >= 0 && gz < a.s0.Z && gy >= 0 && gy < a.s0.Y && gx >= 0 && gx < a.s0.X)
      q = __ldg(reinterpret_cast<const uint4*>(src + (((long long)gz * a.s0.Y + gy) * a.s0.X + gx) * a.s0.C + c * 8));
    reinterpret_cast<uint4*>(xin)[i] = q;
  }
  __syncthreads();
  float acc[8];
#pragma unroll
  for (int v = 0; v < 8; ++v) acc[v] = 0.f;
  const bool co_ok = lane < a.Cout;
  // warp w owns taps (kz, ky) = (w >> 1, 2*(w & 1) + {0,1}) and all kx
  const int kz = warp >> 1;
  for (int kyi = 0; kyi < 2; ++kyi) {
    const int ky = (warp & 1) * 2 + kyi;
    for (int kx = 0; kx < 4; ++kx) {
      const float* wp = a.w + (long long)((kz * 4 + ky) * 4 + kx) * a.ws_tap + (long long)lane * a.ws_out;
      for (int c0 = 0; c0 < Cin; c0 += 16) {         // 16 independent weight loads in flight before the FMA block
        float wv[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          wv[u] = (co_ok && c0 + u < Cin) ? __ldg(wp + (long long)(c0 + u) * a.ws_in) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) wv[u] = __bfloat162float(__float2bfloat16_rn(wv[u]));
        if (c0 + 16 <= Cin) {
#pragma unroll
          for (int u = 0; u < 16; u += 2) {
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              const int z = ((v >> 2) & 1) * 2 + kz, y = ((v >> 1) & 1) * 2 + ky, x = (v & 1) * 2 + kx;
              const __nv_bfloat162 xv = *reinterpret_cast<const __nv_bfloat162*>(xin + (size_t)((z * IB + y) * IB + x) * Cin + c0 + u);
              acc[v] = fmaf(__low2float(xv), wv[u], acc[v]);
              acc[v] = fmaf(__high2float(xv), wv[u + 1], acc[v]);
            }
          }
        } else {                                     // Cin = 8, 24, 40, 56: the last half chunk
#pragma unroll
          for (int u = 0; u < 8; u += 2) {
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              const int z = ((v >> 2) & 1) * 2 + kz, y = ((v >> 1) & 1) * 2 + ky, x = (v & 1) * 2 + kx;
              const __nv_bfloat162 xv = *reinterpret_cast<const __nv_bfloat162*>(xin + (size_t)((z * IB + y) * IB + x) * Cin + c0 + u);
              acc[v] = fmaf(__low2float(xv), wv[u], acc[v]);
              acc[v] = fmaf(__high2float(xv), wv[u + 1], acc[v]);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int v = 0; v < 8; ++v) red[(warp * 8 + v) * 32 + lane] = acc[v];
  __syncthreads();
  {
    const int v = tid >> 5, co = tid & 31;            // 256 threads = 8 voxels x 32 channels
    float sum = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < kWarpsS; ++w8) sum += red[(w8 * 8 + v) * 32 + co];
    const int oz = oz0 + ((v >> 2) & 1), oy = oy0 + ((v >> 1) & 1), ox = ox0 + (v & 1);
    if (co < a.Cout && oz < a.L[0] && oy < a.L[1] && ox < a.L[2]) {
      if (a.slope != 1.f) sum = sum > 0.f ? sum : sum * a.slope;
      bf16* op = reinterpret_cast<bf16*>(a.out) + ((((long long)b * a.OZ + oz + a.out_off[0]) * a.OY + oy + a.out_off[1]) * a.OX + ox + a.out_off[2]) * a.out_C + a.out_coff + co;
      *op = __float2bfloat16_rn(sum);
    }
  }
}

}  // namespace

bool conv_small_supported(const ConvArgs& a) {
  if (a.form != 0 || a.C1 || a.bias || a.ref || a.drop_key || a.accumulate || a.use_lut || a.s0.origins) return false;
  for (int i = 0; i < 3; ++i) if (a.k[i] != 4 || a.stride[i] != 2 || a.pad[i] != 0 || a.conv_off[i]) return false;
  if (a.s0.dtype != DT_BF16 || a.out_dtype != DT_BF16) return false;
  if (a.C0 % 8 || a.C0 > 64 || a.s0.C != a.C0 || a.s0.coff != 0) return false;
  if (a.Cout > 32) return false;
  if ((long long)a.L[0] * a.L[1] * a.L[2] > 512) return false;
  return true;
}

cudaError_t launch_conv_small(const ConvArgs& a, cudaStream_t st) {
  const int nbz = (a.L[0] + 1) / 2, nby = (a.L[1] + 1) / 2, nbx = (a.L[2] + 1) / 2;
  const size_t smem = (size_t)IB * IB * IB * a.C0 * 2 + (size_t)kWarpsS * 8 * 32 * 4;
  conv_small_s2_kernel<<<(unsigned)((long long)a.B * nbz * nby * nbx), kWarpsS * 32, smem, st>>>(a, nbz, nby, nbx); ++g_tem_launches;
  return cudaGetLastError();
}
