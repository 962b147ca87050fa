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

constexpr int WZ = 4, WY = 8, WX = 32;                 // positions per tile; 256 threads = 32 x 8, 4 z-positions each
constexpr int HZw = WZ + 2, HYw = WY + 2, HXw = WX + 2;

__device__ __forceinline__ void decode_tile(long long t, int ntx, int nty, int ntz, int& b, int& z0, int& y0, int& x0) {
  const int tx = (int)(t % ntx); t /= ntx;
  const int ty = (int)(t % nty); t /= nty;
  const int tz = (int)(t % ntz); t /= ntz;
  b = (int)t; z0 = tz * WZ; y0 = ty * WY; x0 = tx * WX;
}

// Ca == 1 (3x3x3, stride 1).  blockIdx.y = kz tap plane, blockIdx.z = 8-channel block of P.  The S halo of a tile is
// staged in shared memory as float; every thread keeps 9 x 8 accumulators over all the tiles of its CTA, then one
// shuffle reduction + atomics.
__global__ void __launch_bounds__(256) wgrad_cin1_kernel(const WgradArgs a, const int ntx, const int nty, const int ntz,
                                                         const long long ntiles, const long long tiles_per_cta) {
  __shared__ float lut[256];
  __shared__ float tile[HZw * HYw * HXw];
  const int tid = threadIdx.x;
  if (a.use_lut) for (int i = tid; i < 256; i += 256) lut[i] = tem_standardize((float)i, a.lut_mean, a.lut_std);
  const int dz = blockIdx.y, cb0 = blockIdx.z * 8;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[t][c] = 0.f;
  const SrcView& S = a.S;
  const int lx = tid & 31, ly = tid >> 5;
  const long long t0 = (long long)blockIdx.x * tiles_per_cta, t1 = min(t0 + tiles_per_cta, ntiles);
  for (long long tl = t0; tl < t1; ++tl) {
    int b, z0, y0, x0; decode_tile(tl, ntx, nty, ntz, b, z0, y0, x0);
    int oz = 0, oy = 0, ox = 0; long long sbase = (long long)b * S.bstride;
    if (S.origins) { oz = S.origins[b * 3]; oy = S.origins[b * 3 + 1]; ox = S.origins[b * 3 + 2]; sbase = 0; }
    __syncthreads();
    const float fill = (a.use_lut && S.origins) ? lut[0] : 0.f;
    // only the z-planes dz .. dz+WZ-1 of the halo are needed by this CTA
    {
      // all loads of a thread are issued before any is consumed (6 independent requests in flight per thread)
      constexpr int NS = (WZ * HYw * HXw + 255) / 256;
      float raw[NS]; bool okv[NS];
      const int zb = z0 + dz - a.pad[0] + S.shift[0] + oz, yb = y0 - a.pad[1] + S.shift[1] + oy, xb = x0 - a.pad[2] + S.shift[2] + ox;
#pragma unroll
      for (int sI = 0; sI < NS; ++sI) {
        const int i = tid + sI * 256;
        const int hx = i % HXw, r = i / HXw, hy = r % HYw, hz = r / HYw;
        const int z = zb + hz, y = yb + hy, x = xb + hx;
        okv[sI] = i < WZ * HYw * HXw && z >= 0 && z < S.Z && y >= 0 && y < S.Y && x >= 0 && x < S.X;
        const long long off = sbase + (((long long)z * S.Y + y) * S.X + x) * S.C + S.coff;
        raw[sI] = 0.f;
        if (okv[sI]) {
          if (S.dtype == DT_U8) raw[sI] = (float)reinterpret_cast<const uint8_t*>(S.p)[off];
          else if (S.dtype == DT_BF16) raw[sI] = bf2f(reinterpret_cast<const bf16*>(S.p)[off]);
          else raw[sI] = reinterpret_cast<const float*>(S.p)[off];
        }
      }
#pragma unroll
      for (int sI = 0; sI < NS; ++sI) {
        const int i = tid + sI * 256;
        if (i < WZ * HYw * HXw) tile[i] = okv[sI] ? ((S.dtype == DT_U8) ? lut[(int)raw[sI]] : raw[sI]) : fill;
      }
    }
    __syncthreads();
    const int py = y0 + ly, px = x0 + lx;
    if (py < a.L[1] && px < a.L[2]) {
#pragma unroll
      for (int j = 0; j < WZ; ++j) {
        const int pz = z0 + j;
        if (pz >= a.L[0]) break;
        const long long po = (long long)b * a.p_bstride + ((((long long)pz + a.p_off[0]) * a.PY + py + a.p_off[1]) * a.PX + px + a.p_off[2]) * a.p_C + a.p_coff + cb0;
        float pb[8];
        if (a.p_dtype == DT_BF16) unpack8(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(a.P) + po)), pb);
        else {
          const float4* fp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.P) + po);
          const float4 x0v = fp[0], x1v = fp[1];
          pb[0] = x0v.x; pb[1] = x0v.y; pb[2] = x0v.z; pb[3] = x0v.w; pb[4] = x1v.x; pb[5] = x1v.y; pb[6] = x1v.z; pb[7] = x1v.w;
        }
#pragma unroll
        for (int tp = 0; tp < 9; ++tp) {
          const float sv = tile[(j * HYw + ly + tp / 3) * HXw + lx + tp % 3];
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[tp][c] = fmaf(sv, pb[c], acc[tp][c]);
        }
      }
    }
  }
  // warp shuffle reduction, then one partial per warp in shared memory, then ONE atomic per output per CTA
  __shared__ float red[8][72];
  const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int tp = 0; tp < 9; ++tp)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float sum = acc[tp][c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == ((tp * 8 + c) & 31)) red[wid][tp * 8 + c] = sum;
    }
  __syncthreads();
  if (tid < 72) {
    float sum = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) sum += red[w8][tid];
    const int tp = tid >> 3, c = tid & 7;
    if (cb0 + c < a.Cb && sum != 0.f) atomicAdd(a.dw + (long long)(dz * 9 + tp) * a.ws_tap + (long long)(cb0 + c) * a.ws_b, sum);
  }
}

// Cb == 1 (P fp32 or bf16, one channel).  blockIdx.y = (kz, ky) tap row, blockIdx.z = 8-channel block of S.
__global__ void __launch_bounds__(256) wgrad_cout1_kernel(const WgradArgs a, const int ntx, const int nty, const int ntz,
                                                          const long long ntiles, const long long tiles_per_cta) {
  __shared__ uint4 tile[WZ * WY * HXw];          // the (dz,dy)-shifted rows of this CTA's 8-channel block
  const int tid = threadIdx.x;
  const int dz = blockIdx.y / 3, dy = blockIdx.y % 3, ca0 = blockIdx.z * 8;
  float acc[3][8];
#pragma unroll
  for (int t = 0; t < 3; ++t)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[t][c] = 0.f;
  const SrcView& S = a.S;
  const int lx = tid & 31, ly = tid >> 5;
  const long long t0 = (long long)blockIdx.x * tiles_per_cta, t1 = min(t0 + tiles_per_cta, ntiles);
  for (long long tl = t0; tl < t1; ++tl) {
    int b, z0, y0, x0; decode_tile(tl, ntx, nty, ntz, b, z0, y0, x0);
    const bf16* Sb = reinterpret_cast<const bf16*>(S.p) + (long long)b * S.bstride;
    __syncthreads();
    {
      constexpr int NS = (WZ * WY * HXw + 255) / 256;
      uint4 raw[NS];
      const int zb = z0 + dz - a.pad[0] + S.shift[0], yb = y0 + dy - a.pad[1] + S.shift[1], xb = x0 - a.pad[2] + S.shift[2];
#pragma unroll
      for (int sI = 0; sI < NS; ++sI) {
        const int i = tid + sI * 256;
        const int hx = i % HXw, r = i / HXw, hy = r % WY, hz = r / WY;
        const int z = zb + hz, y = yb + hy, x = xb + hx;
        raw[sI] = make_uint4(0, 0, 0, 0);
        if (i < WZ * WY * HXw && z >= 0 && z < S.Z && y >= 0 && y < S.Y && x >= 0 && x < S.X)
          raw[sI] = __ldg(reinterpret_cast<const uint4*>(Sb + (((long long)z * S.Y + y) * S.X + x) * S.C + S.coff + ca0));
      }
#pragma unroll
      for (int sI = 0; sI < NS; ++sI) { const int i = tid + sI * 256; if (i < WZ * WY * HXw) tile[i] = raw[sI]; }
    }
    __syncthreads();
    const int py = y0 + ly, px = x0 + lx;
    if (py < a.L[1] && px < a.L[2]) {
#pragma unroll
      for (int j = 0; j < WZ; ++j) {
        const int pz = z0 + j;
        if (pz >= a.L[0]) break;
        const long long po = (long long)b * a.p_bstride + ((((long long)pz + a.p_off[0]) * a.PY + py + a.p_off[1]) * a.PX + px + a.p_off[2]) * a.p_C + a.p_coff;
        const float pv = (a.p_dtype == DT_BF16) ? bf2f(reinterpret_cast<const bf16*>(a.P)[po]) : reinterpret_cast<const float*>(a.P)[po];
#pragma unroll
        for (int tp = 0; tp < 3; ++tp) {
          float sv[8]; unpack8(tile[(j * WY + ly) * HXw + lx + tp], sv);
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[tp][c] = fmaf(sv[c], pv, acc[tp][c]);
        }
      }
    }
  }
  __shared__ float red[8][24];
  const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int tp = 0; tp < 3; ++tp)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float sum = acc[tp][c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == tp * 8 + c) red[wid][tp * 8 + c] = sum;
    }
  __syncthreads();
  if (tid < 24) {
    float sum = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) sum += red[w8][tid];
    const int tp = tid >> 3, c = tid & 7;
    if (ca0 + c < a.Ca && sum != 0.f) atomicAdd(a.dw + (long long)((dz * 3 + dy) * 3 + tp) * a.ws_tap + (long long)(ca0 + c) * a.ws_a, sum);
  }
}

}  // namespace

bool wgrad_c1_supported(const WgradArgs& w) {
  for (int i = 0; i < 3; ++i) if (w.stride[i] != 1) return false;
  if (w.k[1] != 3 || w.k[2] != 3 || (w.k[0] != 3 && w.k[0] != 1)) return false;
  if (w.Ca == 1 && w.Cb >= 8 && w.Cb % 8 == 0 && w.p_C % 8 == 0 && w.p_coff % 8 == 0) return true;
  if (w.Cb == 1 && w.Ca >= 8 && w.Ca % 8 == 0 && w.S.dtype == DT_BF16 && w.S.C % 8 == 0 && w.S.coff % 8 == 0 && !w.S.origins) return true;
  return false;
}

cudaError_t launch_wgrad_c1(const WgradArgs& w_in, cudaStream_t st) {
  WgradArgs a = w_in;
  a.nvox = (long long)a.B * a.L[0] * a.L[1] * a.L[2];
  if (a.nvox == 0) return cudaSuccess;
  const int ntx = (a.L[2] + WX - 1) / WX, nty = (a.L[1] + WY - 1) / WY, ntz = (a.L[0] + WZ - 1) / WZ;
  const long long ntiles = (long long)a.B * ntx * nty * ntz;
  const int gy = (a.Ca == 1) ? a.k[0] : a.k[0] * a.k[1];
  const int gz = (a.Ca == 1) ? a.Cb / 8 : a.Ca / 8;
  long long gx = (148 * 3 + gy * gz - 1) / (gy * gz);          // ~3 CTAs per SM in total: few, long-lived CTAs keep the atomic tail short
  if (gx > ntiles) gx = ntiles;
  const long long per = (ntiles + gx - 1) / gx;
  gx = (ntiles + per - 1) / per;
  dim3 grid((unsigned)gx, gy, gz);
  if (a.Ca == 1) wgrad_cin1_kernel<<<grid, 256, 0, st>>>(a, ntx, nty, ntz, ntiles, per);
  else wgrad_cout1_kernel<<<grid, 256, 0, st>>>(a, ntx, nty, ntz, ntiles, per);
  ++g_tem_launches;
  return cudaGetLastError();
}
