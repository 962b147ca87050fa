// Weight gradients of the single-channel ends of the networks (memory-bound, CUDA cores + warp shuffles):
//   * first layers g0 / d0 (generator.py:54, discriminator.py:39): S has ONE channel (uint8 standardised through
//     the LUT, fp32 fakes with virtual zero padding, or tile origins), P = dy with Cb channels;
//   * last generator layer g11 (generator.py:110): P = d(output) has ONE channel (fp32), S = a10 with Ca channels.
// dw[tap][ca][cb] += sum_{b,p} S[b, p + tap][ca] * P[b,p][cb]     (stride 1, VALID)
#include <stdlib.h>
#include <string.h>
#include "tem_kernels.cuh"

extern unsigned long long g_tem_launches;
static const char* g_wgrad_c1_name = "wgrad_c1_kernel";    // kernel family the last launch_wgrad_c1 call used

namespace {

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

// ------------------------------------------------------------------------------------------------
// Second-generation single-channel weight gradient: dw[tap][cb] += sum_p s[p + tap] * P[p][cb0 + c]
//   * a thread owns VPT = 3 adjacent x positions: the five s values of a halo row serve 3 taps x 3 positions, so the
//     FFMA : (LDS + LDG + unpack) ratio is ~3:1 instead of ~1:1;
//   * the halo of the NEXT tile is fetched into registers before the FFMA block of the current one (no exposed
//     global latency between tiles), P vectors of a z-position are fetched one position ahead;
//   * tiles are 24 x 16 x 4 positions (72 = 3 x 24: no ragged x tile on the 72^3 first-layer gradients).
// The same kernel takes the Cout = 1 case (g11) with the operands swapped: see launch_wgrad_c1.
// ------------------------------------------------------------------------------------------------
constexpr int VPT = 3, V2_LX = 8, V2_TX = VPT * V2_LX, V2_TY = 16, V2_TZ = 4, V2_NT = 128;
constexpr int V2_HX = V2_TX + 2, V2_HY = V2_TY + 2, V2_HALO = V2_TZ * V2_HY * V2_HX;
constexpr int V2_NS = (V2_HALO + V2_NT - 1) / V2_NT;

__device__ __forceinline__ void v2_decode(long long t, int ntx, int nty, int ntz, int& b, int& z0, int& y0, int& x0) {
  const int tx = (int)(t % ntx); t /= ntx;
  const int ty = (int)(t % nty); t /= nty;
  const int tz = (int)(t % ntz); t /= ntz;
  b = (int)t; z0 = tz * V2_TZ; y0 = ty * V2_TY; x0 = tx * V2_TX;
}

template <int SDT>   // dtype of the single-channel operand
__device__ __forceinline__ void v2_fetch(const WgradArgs& a, long long tl, int ntx, int nty, int ntz, int dz, int tid, float* raw) {
  int b, z0, y0, x0; v2_decode(tl, ntx, nty, ntz, b, z0, y0, x0);
  const SrcView& S = a.S;
  const long long sbase = (long long)b * S.bstride;
  const int zb = z0 + dz - a.pad[0] + S.shift[0], yb = y0 - a.pad[1] + S.shift[1], xb = x0 - a.pad[2] + S.shift[2];
#pragma unroll
  for (int sI = 0; sI < V2_NS; ++sI) {
    const int i = tid + sI * V2_NT;
    const int hx = i % V2_HX, r = i / V2_HX, hy = r % V2_HY, hz = r / V2_HY;
    const int z = zb + hz, y = yb + hy, x = xb + hx;
    const bool ok = i < V2_HALO && z >= 0 && z < S.Z && y >= 0 && y < S.Y && x >= 0 && x < S.X;
    const long long off = sbase + (((long long)z * S.Y + y) * S.X + x) * S.C + S.coff;
    float v = -1.f;                                  // < 0 marks "outside": virtual zero padding
    if (ok) {
      if (SDT == DT_U8) v = (float)reinterpret_cast<const uint8_t*>(S.p)[off];
      else if (SDT == DT_BF16) v = bf2f(reinterpret_cast<const bf16*>(S.p)[off]);
      else v = reinterpret_cast<const float*>(S.p)[off];
    } else if (SDT != DT_U8) v = 0.f;
    raw[sI] = v;
  }
}

template <int SDT>
__global__ void __launch_bounds__(V2_NT, 3) wgrad_cin1_v2_kernel(const WgradArgs a, const int ntx, const int nty, const int ntz,
                                                                 const long long ntiles, const long long tiles_per_cta) {
  __shared__ float lut[256];
  __shared__ float tile[V2_HALO];
  __shared__ float red[4][72];
  const int tid = threadIdx.x;
  if (SDT == DT_U8) for (int i = tid; i < 256; i += V2_NT) lut[i] = tem_standardize((float)i, a.lut_mean, a.lut_std);
  const int dz = blockIdx.y, cb0 = blockIdx.z * 8;
  unsigned long long acc2[9][4];                   // channel pairs (FFMA2)
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc2[t][c] = 0ull;
  const int lx = tid & (V2_LX - 1), ly = tid >> 3;
  const long long t0 = (long long)blockIdx.x * tiles_per_cta, t1 = min(t0 + tiles_per_cta, ntiles);
  const bf16* Pp = reinterpret_cast<const bf16*>(a.P);
  float raw[V2_NS];
  if (t0 < t1) v2_fetch<SDT>(a, t0, ntx, nty, ntz, dz, tid, raw);
  for (long long tl = t0; tl < t1; ++tl) {
    __syncthreads();                                 // previous tile fully consumed (and the LUT is ready)
#pragma unroll
    for (int sI = 0; sI < V2_NS; ++sI) {
      const int i = tid + sI * V2_NT;
      if (i < V2_HALO) tile[i] = (SDT == DT_U8) ? (raw[sI] < 0.f ? 0.f : lut[(int)raw[sI]]) : raw[sI];
    }
    __syncthreads();
    if (tl + 1 < t1) v2_fetch<SDT>(a, tl + 1, ntx, nty, ntz, dz, tid, raw);   // in flight during the FFMA block
    int b, z0, y0, x0; v2_decode(tl, ntx, nty, ntz, b, z0, y0, x0);
    const int py = y0 + ly, px = x0 + lx * VPT;
    const bool rowok = py < a.L[1];
    const long long pbase = (long long)b * a.p_bstride + (((long long)a.p_off[0] * a.PY + py + a.p_off[1]) * a.PX + px + a.p_off[2]) * a.p_C + a.p_coff + cb0;
    const long long zstride = (long long)a.PY * a.PX * a.p_C;
    uint4 pq[2][VPT];
    auto loadp = [&](int j, uint4* q) {
      const bool zok = rowok && (z0 + j) < a.L[0];
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        q[v] = make_uint4(0, 0, 0, 0);
        if (zok && px + v < a.L[2]) q[v] = __ldg(reinterpret_cast<const uint4*>(Pp + pbase + (long long)(z0 + j) * zstride + (long long)v * a.p_C));
      }
    };
    loadp(0, pq[0]);
#pragma unroll
    for (int j = 0; j < V2_TZ; ++j) {
      if (j + 1 < V2_TZ) loadp(j + 1, pq[(j + 1) & 1]);
      unsigned long long pf[VPT][4];
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        float f[8]; unpack8(pq[j & 1][v], f);
#pragma unroll
        for (int c = 0; c < 4; ++c) pf[v][c] = tem_pk2(f[2 * c], f[2 * c + 1]);
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float* srow = &tile[(j * V2_HY + ly + r) * V2_HX + lx * VPT];
        unsigned long long sv[VPT + 2];              // the input value in both halves
#pragma unroll
        for (int q = 0; q < VPT + 2; ++q) { const float t = srow[q]; sv[q] = tem_pk2(t, t); }
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
          for (int v = 0; v < VPT; ++v)
#pragma unroll
            for (int c = 0; c < 4; ++c) tem_ffma2(acc2[r * 3 + dx][c], sv[v + dx], pf[v][c]);
      }
    }
  }
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < 4; ++c) tem_upk2(acc2[t][c], acc[t][2 * c], acc[t][2 * c + 1]);
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
    const float sum = red[0][tid] + red[1][tid] + red[2][tid] + red[3][tid];
    const int tp = tid >> 3, c = tid & 7;
    if (cb0 + c < a.Cb && sum != 0.f) atomicAdd(a.dw + (long long)(dz * 9 + tp) * a.ws_tap + (long long)(cb0 + c) * a.ws_b, sum);
  }
}

static cudaError_t launch_cin1_v2(const WgradArgs& a, cudaStream_t st) {
  const int ntx = (a.L[2] + V2_TX - 1) / V2_TX, nty = (a.L[1] + V2_TY - 1) / V2_TY, ntz = (a.L[0] + V2_TZ - 1) / V2_TZ;
  const long long ntiles = (long long)a.B * ntx * nty * ntz;
  const int gy = a.k[0], gz = a.Cb / 8;
  long long gx = (148 * 3 + gy * gz - 1) / (gy * gz);
  if (gx > ntiles) gx = ntiles;
  const long long per = (ntiles + gx - 1) / gx;
  gx = (ntiles + per - 1) / per;
  dim3 grid((unsigned)gx, gy, gz);
  if (a.S.dtype == DT_U8) wgrad_cin1_v2_kernel<DT_U8><<<grid, V2_NT, 0, st>>>(a, ntx, nty, ntz, ntiles, per);
  else if (a.S.dtype == DT_BF16) wgrad_cin1_v2_kernel<DT_BF16><<<grid, V2_NT, 0, st>>>(a, ntx, nty, ntz, ntiles, per);
  else wgrad_cin1_v2_kernel<DT_F32><<<grid, V2_NT, 0, st>>>(a, ntx, nty, ntz, ntiles, per);
  ++g_tem_launches;
  return cudaGetLastError();
}

}  // namespace

bool wgrad_c1_supported(const WgradArgs& w) {
  for (int i = 0; i < 3; ++i) if (w.stride[i] != 1) return false;
  if (w.k[1] != 3 || w.k[2] != 3 || (w.k[0] != 3 && w.k[0] != 1)) return false;
  if (w.Ca == 1 && w.Cb >= 8 && w.Cb % 8 == 0 && w.p_C % 8 == 0 && w.p_coff % 8 == 0) return true;
  if (w.Cb == 1 && w.Ca >= 8 && w.Ca % 8 == 0 && w.S.dtype == DT_BF16 && w.S.C % 8 == 0 && w.S.coff % 8 == 0 && !w.S.origins) return true;
  return false;
}

const char* wgrad_c1_last_name() { return g_wgrad_c1_name; }

cudaError_t launch_wgrad_c1(const WgradArgs& w_in, cudaStream_t st) {
  g_wgrad_c1_name = "wgrad_c1_kernel";
  WgradArgs a = w_in;
  a.nvox = (long long)a.B * a.L[0] * a.L[1] * a.L[2];
  if (a.nvox == 0) return cudaSuccess;
  static const bool old_c1 = getenv("TEM_WGRAD_C1_V1") != nullptr;   // debug knob: first-generation kernels
  if (!old_c1 && a.k[0] == 3) {
    if (a.Ca == 1 && a.p_dtype == DT_BF16 && !a.S.origins && (a.use_lut ? a.S.dtype == DT_U8 : a.S.dtype != DT_U8)) {
      if (wgrad_c1tc_supported(a)) { g_wgrad_c1_name = "wgrad_c1tc_kernel"; return launch_wgrad_c1tc(a, st); }
      return launch_cin1_v2(a, st);
    }
    if (a.Cb == 1 && a.S.dtype == DT_BF16 && a.p_dtype != DT_U8 && a.p_C == 1 && a.p_coff == 0 && a.p_off[0] == 0 && a.p_off[1] == 0 &&
        a.p_off[2] == 0 && a.PZ == a.L[0] && a.PY == a.L[1] && a.PX == a.L[2]) {
      // dw[tap][ca] = sum_p S[p + tap][ca] P[p] = sum_u S[u][ca] P'[u + tap' - 2], tap' = 2 - tap: the single-channel
      // operand becomes dy with a virtual zero padding of 2, the multi-channel one the layer input, taps are flipped
      WgradArgs s; memset(&s, 0, sizeof(s));
      s.S.p = a.P; s.S.dtype = a.p_dtype; s.S.Z = a.PZ; s.S.Y = a.PY; s.S.X = a.PX; s.S.C = 1; s.S.coff = 0;
      s.S.bstride = a.p_bstride; s.S.origins = nullptr;
      for (int i = 0; i < 3; ++i) { s.S.shift[i] = -2; s.L[i] = a.L[i] + 2; s.p_off[i] = a.S.shift[i]; s.k[i] = 3; s.stride[i] = 1; s.pad[i] = 0; }
      s.Ca = 1; s.Cb = a.Ca;
      s.P = a.S.p; s.p_dtype = DT_BF16; s.PZ = a.S.Z; s.PY = a.S.Y; s.PX = a.S.X; s.p_C = a.S.C; s.p_coff = a.S.coff; s.p_bstride = a.S.bstride;
      s.B = a.B;
      s.dw = a.dw + 26 * a.ws_tap; s.ws_tap = -a.ws_tap; s.ws_a = 0; s.ws_b = a.ws_a;
      s.use_lut = 0;
      if (wgrad_c1tc_supported(s)) { g_wgrad_c1_name = "wgrad_c1tc_kernel"; return launch_wgrad_c1tc(s, st); }
      return launch_cin1_v2(s, st);
    }
  }
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
