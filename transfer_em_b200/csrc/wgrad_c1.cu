// Weight gradients of the single-channel ends of the networks (memory-bound, CUDA cores + warp shuffles):
//   * first layers g0 / d0 (generator.py:54, discriminator.py:39): S has ONE channel (uint8 standardised through
//     the LUT, fp32 fakes with virtual zero padding, or tile origins), P = dy with Cb channels;
//   * last generator layer g11 (generator.py:110): P = d(output) has ONE channel (fp32), S = a10 with Ca channels.
// dw[tap][ca][cb] += sum_{b,p} S[b, p + tap][ca] * P[b,p][cb]     (stride 1, VALID)
#include "tem_kernels.cuh"

extern unsigned long long g_tem_launches;

namespace {

__device__ __forceinline__ float load1(const SrcView& S, long long off, const float* lut) {
  if (S.dtype == DT_U8) return lut[reinterpret_cast<const uint8_t*>(S.p)[off]];
  if (S.dtype == DT_BF16) return bf2f(reinterpret_cast<const bf16*>(S.p)[off]);
  return reinterpret_cast<const float*>(S.p)[off];
}

// Ca == 1.  blockIdx.y = kz tap plane, blockIdx.z = 8-channel block of P.  Each thread owns KY*KX x 8 accumulators
// over its positions (lanes = consecutive positions -> coalesced P and S reads), then a shuffle reduction.
template <int KYX>
__global__ void __launch_bounds__(128) wgrad_cin1_kernel(const WgradArgs a) {
  __shared__ float lut[256];
  if (a.use_lut) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = tem_standardize((float)i, a.lut_mean, a.lut_std);
    __syncthreads();
  }
  const int dz = blockIdx.y, cb0 = blockIdx.z * 8;
  const int kx = a.k[2];
  float acc[KYX][8];
#pragma unroll
  for (int t = 0; t < KYX; ++t)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[t][c] = 0.f;
  const SrcView& S = a.S;
  const long long v0 = (long long)blockIdx.x * a.vox_per_cta, v1 = min(v0 + a.vox_per_cta, a.nvox);
  for (long long v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
    long long t = v;
    const int lx = (int)(t % a.L[2]); t /= a.L[2];
    const int ly = (int)(t % a.L[1]); t /= a.L[1];
    const int lz = (int)(t % a.L[0]); t /= a.L[0];
    const int b = (int)t;
    const long long po = (long long)b * a.p_bstride + ((((long long)lz + a.p_off[0]) * a.PY + ly + a.p_off[1]) * a.PX + lx + a.p_off[2]) * a.p_C + a.p_coff + cb0;
    float pb[8];
    if (a.p_dtype == DT_BF16) unpack8(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(a.P) + po)), pb);
    else {
      const float4* fp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.P) + po);
      const float4 x0 = fp[0], x1 = fp[1];
      pb[0] = x0.x; pb[1] = x0.y; pb[2] = x0.z; pb[3] = x0.w; pb[4] = x1.x; pb[5] = x1.y; pb[6] = x1.z; pb[7] = x1.w;
    }
    int tz = lz + dz - a.pad[0] + S.shift[0], ty0 = ly - a.pad[1] + S.shift[1], tx0 = lx - a.pad[2] + S.shift[2];
    long long sbase;
    if (S.origins) { tz += S.origins[b * 3]; ty0 += S.origins[b * 3 + 1]; tx0 += S.origins[b * 3 + 2]; sbase = 0; }
    else sbase = (long long)b * S.bstride;
    const float fill = (a.use_lut && S.origins) ? lut[0] : 0.f;
    const bool zin = tz >= 0 && tz < S.Z;
#pragma unroll
    for (int tp = 0; tp < KYX; ++tp) {
      const int ty = ty0 + tp / kx, tx = tx0 + tp % kx;
      float sv = fill;
      if (zin && ty >= 0 && ty < S.Y && tx >= 0 && tx < S.X)
        sv = load1(S, sbase + (((long long)tz * S.Y + ty) * S.X + tx) * S.C + S.coff, lut);
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[tp][c] = fmaf(sv, pb[c], acc[tp][c]);
    }
  }
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int tp = 0; tp < KYX; ++tp)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float s = acc[tp][c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == ((tp * 8 + c) & 31) && cb0 + c < a.Cb && s != 0.f)
        atomicAdd(a.dw + (long long)(dz * KYX + tp) * a.ws_tap + (long long)(cb0 + c) * a.ws_b, s);
    }
}

// Cb == 1 (P fp32 or bf16, one channel).  blockIdx.y = (kz, ky) tap row, blockIdx.z = 8-channel block of S.
template <int KX>
__global__ void __launch_bounds__(128) wgrad_cout1_kernel(const WgradArgs a) {
  const int dz = blockIdx.y / a.k[1], dy = blockIdx.y % a.k[1], ca0 = blockIdx.z * 8;
  float acc[KX][8];
#pragma unroll
  for (int t = 0; t < KX; ++t)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[t][c] = 0.f;
  const SrcView& S = a.S;
  const long long v0 = (long long)blockIdx.x * a.vox_per_cta, v1 = min(v0 + a.vox_per_cta, a.nvox);
  for (long long v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
    long long t = v;
    const int lx = (int)(t % a.L[2]); t /= a.L[2];
    const int ly = (int)(t % a.L[1]); t /= a.L[1];
    const int lz = (int)(t % a.L[0]); t /= a.L[0];
    const int b = (int)t;
    const long long po = (long long)b * a.p_bstride + ((((long long)lz + a.p_off[0]) * a.PY + ly + a.p_off[1]) * a.PX + lx + a.p_off[2]) * a.p_C + a.p_coff;
    const float pv = (a.p_dtype == DT_BF16) ? bf2f(reinterpret_cast<const bf16*>(a.P)[po]) : reinterpret_cast<const float*>(a.P)[po];
    const int tz = lz + dz - a.pad[0] + S.shift[0], ty = ly + dy - a.pad[1] + S.shift[1], tx0 = lx - a.pad[2] + S.shift[2];
    if (tz < 0 || tz >= S.Z || ty < 0 || ty >= S.Y) continue;
    const long long so = (long long)b * S.bstride + (((long long)tz * S.Y + ty) * S.X + tx0) * S.C + S.coff + ca0;
#pragma unroll
    for (int tp = 0; tp < KX; ++tp) {
      const int tx = tx0 + tp;
      if (tx < 0 || tx >= S.X) continue;
      float sv[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(S.p) + so + (long long)tp * S.C)), sv);
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[tp][c] = fmaf(sv[c], pv, acc[tp][c]);
    }
  }
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int tp = 0; tp < KX; ++tp)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float s = acc[tp][c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == ((tp * 8 + c) & 31) && ca0 + c < a.Ca && s != 0.f)
        atomicAdd(a.dw + (long long)((dz * a.k[1] + dy) * KX + tp) * a.ws_tap + (long long)(ca0 + c) * a.ws_a, s);
    }
}

}  // namespace

bool wgrad_c1_supported(const WgradArgs& w) {
  for (int i = 0; i < 3; ++i) if (w.stride[i] != 1) return false;
  if (w.k[1] != 3 || w.k[2] != 3) return false;
  if (w.Ca == 1 && w.Cb >= 8 && w.Cb % 8 == 0 && w.p_C % 8 == 0 && w.p_coff % 8 == 0) return true;
  if (w.Cb == 1 && w.Ca >= 8 && w.Ca % 8 == 0 && w.S.dtype == DT_BF16 && w.S.C % 8 == 0 && w.S.coff % 8 == 0 && !w.S.origins) return true;
  return false;
}

cudaError_t launch_wgrad_c1(const WgradArgs& w_in, cudaStream_t st) {
  WgradArgs a = w_in;
  a.nvox = (long long)a.B * a.L[0] * a.L[1] * a.L[2];
  if (a.nvox == 0) return cudaSuccess;
  long long per = 128 * 48;                       // ~48 positions per thread amortise the shuffle reduction
  long long gx = (a.nvox + per - 1) / per;
  if (gx > 148 * 8) { gx = 148 * 8; per = (a.nvox + gx - 1) / gx; }
  a.vox_per_cta = per;
  gx = (a.nvox + per - 1) / per;
  if (a.Ca == 1) {
    dim3 grid((unsigned)gx, a.k[0], a.Cb / 8);
    wgrad_cin1_kernel<9><<<grid, 128, 0, st>>>(a);
  } else {
    dim3 grid((unsigned)gx, a.k[0] * a.k[1], a.Ca / 8);
    wgrad_cout1_kernel<3><<<grid, 128, 0, st>>>(a);
  }
  ++g_tem_launches;
  return cudaGetLastError();
}
