// Dedicated memory-bound kernels for the single-channel ends of the networks (no tensor cores: K = 27 or N = 1):
//   conv_c1in_kernel<CO>:  1 -> CO channels, 3x3x3 stride 1 (g0 / d0 forward: generator.py:54, discriminator.py:39;
//                          and, with flipped weights + zero padding, the data gradient of the Cout = 1 layer g11)
//   conv_c1out_kernel<CI>: CI -> 1 channel, 3x3x3 stride 1 (g11 forward: generator.py:110), fp32 output
// A CTA stages the input halo of a 4 x 8 x 32 output tile in shared memory once (uint8 goes through the standardise
// LUT, fp32/bf16 fakes get their virtual zero padding, tiles may come from per-sample origins in one volume), every
// thread then produces 4 z-consecutive voxels with a register sliding window: FMA-bound inner loop, 16 B coalesced
// stores, fused LeakyReLU / LeakyReLU' * dropout epilogues.
#include <stdlib.h>
#include <string.h>
#include "tem_kernels.cuh"

extern unsigned long long g_tem_launches;

namespace {

constexpr int TZ = 4, TY = 8, TXT = 32;
constexpr int HZc = TZ + 2, HYc = TY + 2, HXc = TXT + 2;

template <int CO>
__global__ void __launch_bounds__(256) conv_c1in_kernel(const ConvArgs a, const int ntx, const int nty, const int ntz, const int flip) {
  __shared__ float lut[256];
  __shared__ float tile[HZc * HYc * HXc];
  __shared__ __align__(16) float wsm[27 * CO];
  const int tid = threadIdx.x;
  if (a.use_lut) for (int i = tid; i < 256; i += 256) lut[i] = tem_standardize((float)i, a.lut_mean, a.lut_std);
  for (int i = tid; i < 27 * CO; i += 256) {
    const int co = i % CO; int tap = i / CO;
    if (flip) tap = 26 - tap;
    wsm[i] = (co < a.Cout) ? bf2f(__float2bfloat16_rn(a.w[tap * a.ws_tap + (long long)co * a.ws_out])) : 0.f;
  }
  int t = blockIdx.x;
  const int tx = t % ntx; t /= ntx;
  const int ty = t % nty; t /= nty;
  const int tz = t % ntz; t /= ntz;
  const int b = t;
  const int z0 = tz * TZ, y0 = ty * TY, x0 = tx * TXT;
  __syncthreads();
  // stage: logical conv-input coordinate of halo voxel (hz,hy,hx) = (z0,y0,x0) + h - pad ; tensor = logical + shift
  const SrcView& S = a.s0;
  const int pad = flip ? 2 : 0;
  int oz = 0, oy = 0, ox = 0; long long sbase = (long long)b * S.bstride;
  if (S.origins) { oz = S.origins[b * 3]; oy = S.origins[b * 3 + 1]; ox = S.origins[b * 3 + 2]; sbase = 0; }
  const float fill = (a.use_lut && S.origins) ? lut[0] : 0.f;
  {
    // all loads of a thread are issued before any is consumed
    constexpr int NS = (HZc * HYc * HXc + 255) / 256;
    float raw[NS]; bool okv[NS];
    const int zb = z0 - pad + S.shift[0] + oz, yb = y0 - pad + S.shift[1] + oy, xb = x0 - pad + S.shift[2] + ox;
#pragma unroll
    for (int sI = 0; sI < NS; ++sI) {
      const int i = tid + sI * 256;
      const int hx = i % HXc, r = i / HXc, hy = r % HYc, hz = r / HYc;
      const int z = zb + hz, y = yb + hy, x = xb + hx;
      okv[sI] = i < HZc * HYc * HXc && z >= 0 && z < S.Z && y >= 0 && y < S.Y && x >= 0 && x < S.X;
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
      if (i < HZc * HYc * HXc) tile[i] = okv[sI] ? ((S.dtype == DT_U8) ? lut[(int)raw[sI]] : raw[sI]) : fill;
    }
  }
  __syncthreads();
  const int lx = tid & 31, ly = tid >> 5;        // 32 x 8 threads, 4 voxels in z each
  unsigned long long acc[TZ][CO / 2];              // channel pairs (FFMA2)
#pragma unroll
  for (int j = 0; j < TZ; ++j)
#pragma unroll
    for (int c = 0; c < CO / 2; ++c) acc[j][c] = 0ull;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy)
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      unsigned long long col[HZc];                   // the input value in both halves
#pragma unroll
      for (int hz = 0; hz < HZc; ++hz) { const float t = tile[(hz * HYc + ly + dy) * HXc + lx + dx]; col[hz] = tem_pk2(t, t); }
#pragma unroll
      for (int dz = 0; dz < 3; ++dz) {
        const float* wr = wsm + ((dz * 3 + dy) * 3 + dx) * CO;
        unsigned long long wv[CO / 2];
#pragma unroll
        for (int c = 0; c < CO / 2; c += 2) { const ulonglong2 w4 = *reinterpret_cast<const ulonglong2*>(wr + 2 * c); wv[c] = w4.x; wv[c + 1] = w4.y; }
#pragma unroll
        for (int j = 0; j < TZ; ++j)
#pragma unroll
          for (int c = 0; c < CO / 2; ++c) tem_ffma2(acc[j][c], col[j + dz], wv[c]);
      }
    }
  const int oy_ = y0 + ly, ox_ = x0 + lx;
  if (oy_ >= a.L[1] || ox_ >= a.L[2]) return;
#pragma unroll
  for (int j = 0; j < TZ; ++j) {
    const int oz_ = z0 + j;
    if (oz_ >= a.L[0]) break;
    float v[CO];
#pragma unroll
    for (int c = 0; c < CO / 2; ++c) tem_upk2(acc[j][c], v[2 * c], v[2 * c + 1]);
    if (a.ref) {
      const long long ro = ((((long long)b * a.RZ + oz_ + a.ref_off[0]) * a.RY + oy_ + a.ref_off[1]) * a.RX + ox_ + a.ref_off[2]) * a.ref_C + a.ref_coff;
#pragma unroll
      for (int c = 0; c < CO; c += 8) {
        float f[8]; unpack8(*reinterpret_cast<const uint4*>(a.ref + ro + c), f);
#pragma unroll
        for (int u = 0; u < 8; ++u) v[c + u] *= (f[u] > 0.f) ? 1.f : a.ref_slope;
      }
    }
    if (a.drop_key) {
      const uint32_t di = (uint32_t)(((((long long)b * a.L[0] + oz_) * a.L[1] + oy_) * a.L[2] + ox_) * a.Cout);
#pragma unroll
      for (int c = 0; c < CO; ++c) v[c] *= 2.f * tem_keep(a.drop_key, di + c);
    }
    bf16* op = reinterpret_cast<bf16*>(a.out) + ((((long long)b * a.OZ + oz_ + a.out_off[0]) * a.OY + oy_ + a.out_off[1]) * a.OX + ox_ + a.out_off[2]) * a.out_C + a.out_coff;
#pragma unroll
    for (int c = 0; c < CO; c += 8) {
      float o[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { o[u] = v[c + u]; if (a.slope != 1.f) o[u] = o[u] > 0.f ? o[u] : o[u] * a.slope; }
      uint4 pk; pk.x = pack2(o[0], o[1]); pk.y = pack2(o[2], o[3]); pk.z = pack2(o[4], o[5]); pk.w = pack2(o[6], o[7]);
      *reinterpret_cast<uint4*>(op + c) = pk;
    }
  }
}

// CI -> 1 channel, fp32 output (linear), bf16 input staged as 8-channel planes of uint4.  flip = 1: the data gradient of a
// 1 -> CI first layer (g0 / d0 into the fakes, cgan.py:161,170): taps reversed, zero padding 2, a window of the padded input
// (conv_off) and accumulation into the fp32 gradient.
template <int CI>
__global__ void __launch_bounds__(256) conv_c1out_kernel(const ConvArgs a, const int ntx, const int nty, const int ntz, const int flip) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint4* tile = reinterpret_cast<uint4*>(smem_raw);                 // [CI/8][HZ][HY][HX]
  __shared__ __align__(16) float wsm[27 * CI];
  const int tid = threadIdx.x;
  for (int i = tid; i < 27 * CI; i += 256) {
    const int tap = flip ? 26 - i / CI : i / CI;
    wsm[i] = bf2f(__float2bfloat16_rn(a.w[tap * a.ws_tap + (long long)(i % CI) * a.ws_in]));
  }
  int t = blockIdx.x;
  const int tx = t % ntx; t /= ntx;
  const int ty = t % nty; t /= nty;
  const int tz = t % ntz; t /= ntz;
  const int b = t;
  const int z0 = tz * TZ, y0 = ty * TY, x0 = tx * TXT;
  const SrcView& S = a.s0;
  const int pad = flip ? 2 : 0;
  constexpr int HV = HZc * HYc * HXc, PL = CI / 8;
  const bf16* Sb = reinterpret_cast<const bf16*>(S.p) + (long long)b * S.bstride;
  for (int i = tid; i < HV * PL; i += 256) {
    const int pl = i % PL; const int v = i / PL;
    const int hx = v % HXc; const int r = v / HXc; const int hy = r % HYc; const int hz = r / HYc;
    const int z = z0 + hz + a.conv_off[0] - pad + S.shift[0], y = y0 + hy + a.conv_off[1] - pad + S.shift[1], x = x0 + hx + a.conv_off[2] - pad + S.shift[2];
    uint4 q = make_uint4(0, 0, 0, 0);
    if (z >= 0 && z < S.Z && y >= 0 && y < S.Y && x >= 0 && x < S.X)
      q = __ldg(reinterpret_cast<const uint4*>(Sb + (((long long)z * S.Y + y) * S.X + x) * S.C + S.coff + pl * 8));
    tile[pl * HV + v] = q;
  }
  __syncthreads();
  const int lx = tid & 31, ly = tid >> 5;
  float acc[TZ] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int dy = 0; dy < 3; ++dy)
#pragma unroll
    for (int dx = 0; dx < 3; ++dx)
#pragma unroll
      for (int pl = 0; pl < PL; ++pl)
#pragma unroll
        for (int hz = 0; hz < HZc; ++hz) {
          float f[8]; unpack8(tile[pl * HV + (hz * HYc + ly + dy) * HXc + lx + dx], f);
#pragma unroll
          for (int dz = 0; dz < 3; ++dz) {
            const int j = hz - dz;
            if (j < 0 || j >= TZ) continue;
            const float* wr = wsm + ((dz * 3 + dy) * 3 + dx) * CI + pl * 8;
            const float4 w0 = *reinterpret_cast<const float4*>(wr), w1 = *reinterpret_cast<const float4*>(wr + 4);
            acc[j] = fmaf(f[0], w0.x, fmaf(f[1], w0.y, fmaf(f[2], w0.z, fmaf(f[3], w0.w,
                     fmaf(f[4], w1.x, fmaf(f[5], w1.y, fmaf(f[6], w1.z, fmaf(f[7], w1.w, acc[j]))))))));
          }
        }
  const int oy_ = y0 + ly, ox_ = x0 + lx;
  if (oy_ >= a.L[1] || ox_ >= a.L[2]) return;
  float* out = reinterpret_cast<float*>(a.out);
#pragma unroll
  for (int j = 0; j < TZ; ++j) {
    const int oz_ = z0 + j;
    if (oz_ >= a.L[0]) break;
    float v = acc[j];
    if (a.slope != 1.f) v = v > 0.f ? v : v * a.slope;
    if (a.st_out) {            // tiled inference: straight into the stitched uint8 volume (utils.py:109-121)
      const int cz = oz_ - a.st_tpad, cy = oy_ - a.st_tpad, cx = ox_ - a.st_tpad;
      if (cz < 0 || cz >= a.st_od || cy < 0 || cy >= a.st_od || cx < 0 || cx >= a.st_od) continue;
      const long long gx = a.st_index[b * 3 + 0] + cx, gy = a.st_index[b * 3 + 1] + cy, gz = a.st_index[b * 3 + 2] + cz;
      if (gx < a.st_OX && gy < a.st_OY && gz < a.st_OZ) a.st_out[(gz * a.st_OY + gy) * a.st_OX + gx] = tem_to_u8_round(v, a.st_mean, a.st_std);
      continue;
    }
    float* op = out + ((((long long)b * a.OZ + oz_ + a.out_off[0]) * a.OY + oy_ + a.out_off[1]) * a.OX + ox_ + a.out_off[2]) * a.out_C + a.out_coff;
    *op = a.accumulate ? *op + v : v;
  }
}

// ------------------------------------------------------------------------------------------------
// Second-generation 1 -> 8 forward (g0 / d0, incl. the uint8 LUT entry and per-tile origins of tiled inference):
//   * 128 threads own a 24 x 16 x 4 output tile, every thread 3 x-positions (x, x+8, x+16: the 8 lanes of a row still
//     write 128 B contiguous) x 4 z = 96 accumulators, so one broadcast weight load (2 x LDS.128) feeds 12 FFMA x 8 and
//     the index arithmetic is amortised over 12 outputs instead of 4;
//   * CTAs are persistent over tiles: LUT and weights are staged once, the halo of the next tile is fetched into
//     registers while the current one is computed.
// ------------------------------------------------------------------------------------------------
constexpr int W2_TX = 24, W2_TY = 16, W2_TZ = 4, W2_NT = 128, W2_V = 3;
constexpr int W2_HX = W2_TX + 2, W2_HY = W2_TY + 2, W2_HZ = W2_TZ + 2, W2_HALO = W2_HZ * W2_HY * W2_HX;
constexpr int W2_NS = (W2_HALO + W2_NT - 1) / W2_NT;

template <int SDT>
__device__ __forceinline__ void w2_fetch(const ConvArgs& a, long long tl, int ntx, int nty, int ntz, int tid, float* raw) {
  long long t = tl;
  const int tx = (int)(t % ntx); t /= ntx;
  const int ty = (int)(t % nty); t /= nty;
  const int tz = (int)(t % ntz); t /= ntz;
  const int b = (int)t;
  const SrcView& S = a.s0;
  int oz = 0, oy = 0, ox = 0; long long sbase = (long long)b * S.bstride;
  if (S.origins) { oz = S.origins[b * 3]; oy = S.origins[b * 3 + 1]; ox = S.origins[b * 3 + 2]; sbase = 0; }
  const int zb = tz * W2_TZ + S.shift[0] + oz, yb = ty * W2_TY + S.shift[1] + oy, xb = tx * W2_TX + S.shift[2] + ox;
#pragma unroll
  for (int sI = 0; sI < W2_NS; ++sI) {
    const int i = tid + sI * W2_NT;
    const int hx = i % W2_HX, r = i / W2_HX, hy = r % W2_HY, hz = r / W2_HY;
    const int z = zb + hz, y = yb + hy, x = xb + hx;
    const bool ok = i < W2_HALO && z >= 0 && z < S.Z && y >= 0 && y < S.Y && x >= 0 && x < S.X;
    const long long off = sbase + (((long long)z * S.Y + y) * S.X + x) * S.C + S.coff;
    float v = (SDT == DT_U8) ? -1.f : 0.f;           // u8: < 0 marks "outside"
    if (ok) {
      if (SDT == DT_U8) v = (float)reinterpret_cast<const uint8_t*>(S.p)[off];
      else if (SDT == DT_BF16) v = bf2f(reinterpret_cast<const bf16*>(S.p)[off]);
      else v = reinterpret_cast<const float*>(S.p)[off];
    }
    raw[sI] = v;
  }
}

template <int SDT>
__global__ void __launch_bounds__(W2_NT, 3) conv_c1in_v2_kernel(const ConvArgs a, const int ntx, const int nty, const int ntz,
                                                                const long long ntiles, const long long tiles_per_cta) {
  __shared__ float lut[256];
  __shared__ float tile[W2_HALO];
  __shared__ __align__(16) float wsm[27 * 8];
  const int tid = threadIdx.x;
  if (SDT == DT_U8) for (int i = tid; i < 256; i += W2_NT) lut[i] = tem_standardize((float)i, a.lut_mean, a.lut_std);
  for (int i = tid; i < 27 * 8; i += W2_NT) {
    const int co = i & 7, tap = i >> 3;
    wsm[i] = (co < a.Cout) ? bf2f(__float2bfloat16_rn(a.w[tap * a.ws_tap + (long long)co * a.ws_out])) : 0.f;
  }
  const float fill = (SDT == DT_U8 && a.s0.origins) ? tem_standardize(0.f, a.lut_mean, a.lut_std) : 0.f;   // as the first-generation kernel
  const int lx = tid & 7, ly = tid >> 3;
  const long long t0 = (long long)blockIdx.x * tiles_per_cta, t1 = min(t0 + tiles_per_cta, ntiles);
  float raw[W2_NS];
  if (t0 < t1) w2_fetch<SDT>(a, t0, ntx, nty, ntz, tid, raw);
  for (long long tl = t0; tl < t1; ++tl) {
    __syncthreads();
#pragma unroll
    for (int sI = 0; sI < W2_NS; ++sI) {
      const int i = tid + sI * W2_NT;
      if (i < W2_HALO) tile[i] = (SDT == DT_U8) ? (raw[sI] < 0.f ? fill : lut[(int)raw[sI]]) : raw[sI];
    }
    __syncthreads();
    if (tl + 1 < t1) w2_fetch<SDT>(a, tl + 1, ntx, nty, ntz, tid, raw);
    unsigned long long acc[W2_V][W2_TZ][4];          // channel pairs (FFMA2)
#pragma unroll
    for (int v = 0; v < W2_V; ++v)
#pragma unroll
      for (int j = 0; j < W2_TZ; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[v][j][c] = 0ull;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        unsigned long long col[W2_V][W2_HZ];         // the input value in both halves
#pragma unroll
        for (int v = 0; v < W2_V; ++v)
#pragma unroll
          for (int hz = 0; hz < W2_HZ; ++hz) { const float t = tile[(hz * W2_HY + ly + dy) * W2_HX + lx + 8 * v + dx]; col[v][hz] = tem_pk2(t, t); }
#pragma unroll
        for (int dz = 0; dz < 3; ++dz) {
          const ulonglong2 w0 = *reinterpret_cast<const ulonglong2*>(wsm + ((dz * 3 + dy) * 3 + dx) * 8);
          const ulonglong2 w1 = *reinterpret_cast<const ulonglong2*>(wsm + ((dz * 3 + dy) * 3 + dx) * 8 + 4);
          const unsigned long long wv[4] = {w0.x, w0.y, w1.x, w1.y};
#pragma unroll
          for (int v = 0; v < W2_V; ++v)
#pragma unroll
            for (int j = 0; j < W2_TZ; ++j)
#pragma unroll
              for (int c = 0; c < 4; ++c) tem_ffma2(acc[v][j][c], col[v][j + dz], wv[c]);
        }
      }
    long long t = tl;
    const int tx = (int)(t % ntx); t /= ntx;
    const int ty = (int)(t % nty); t /= nty;
    const int tz = (int)(t % ntz); t /= ntz;
    const int b = (int)t;
    const int oy_ = ty * W2_TY + ly;
    if (oy_ < a.L[1]) {
#pragma unroll
      for (int v = 0; v < W2_V; ++v) {
        const int ox_ = tx * W2_TX + lx + 8 * v;
        if (ox_ >= a.L[2]) continue;
#pragma unroll
        for (int j = 0; j < W2_TZ; ++j) {
          const int oz_ = tz * W2_TZ + j;
          if (oz_ >= a.L[0]) break;
          float o[8];
#pragma unroll
          for (int u = 0; u < 4; ++u) tem_upk2(acc[v][j][u], o[2 * u], o[2 * u + 1]);
#pragma unroll
          for (int u = 0; u < 8; ++u) { if (a.slope != 1.f) o[u] = o[u] > 0.f ? o[u] : o[u] * a.slope; }
          uint4 pk; pk.x = pack2(o[0], o[1]); pk.y = pack2(o[2], o[3]); pk.z = pack2(o[4], o[5]); pk.w = pack2(o[6], o[7]);
          bf16* op = reinterpret_cast<bf16*>(a.out) + ((((long long)b * a.OZ + oz_ + a.out_off[0]) * a.OY + oy_ + a.out_off[1]) * a.OX + ox_ + a.out_off[2]) * a.out_C + a.out_coff;
          *reinterpret_cast<uint4*>(op) = pk;
        }
      }
    }
  }
}

static bool c1in_v2_ok(const ConvArgs& a) {
  static const bool off = getenv("TEM_CONV_C1_V1") != nullptr;      // debug knob: first-generation kernel
  return !off && a.form == 0 && a.C0 == 1 && a.Cout == 8 && !a.ref && !a.drop_key && !a.accumulate && !a.bias &&
         (a.use_lut ? a.s0.dtype == DT_U8 : a.s0.dtype != DT_U8);
}
static cudaError_t launch_c1in_v2(const ConvArgs& a, cudaStream_t st) {
  const int ntx = (a.L[2] + W2_TX - 1) / W2_TX, nty = (a.L[1] + W2_TY - 1) / W2_TY, ntz = (a.L[0] + W2_TZ - 1) / W2_TZ;
  const long long ntiles = (long long)a.B * ntx * nty * ntz;
  if (ntiles == 0) return cudaSuccess;
  long long gx = 148 * 3;
  if (gx > ntiles) gx = ntiles;
  const long long per = (ntiles + gx - 1) / gx;
  gx = (ntiles + per - 1) / per;
  if (a.s0.dtype == DT_U8) conv_c1in_v2_kernel<DT_U8><<<(unsigned)gx, W2_NT, 0, st>>>(a, ntx, nty, ntz, ntiles, per);
  else if (a.s0.dtype == DT_BF16) conv_c1in_v2_kernel<DT_BF16><<<(unsigned)gx, W2_NT, 0, st>>>(a, ntx, nty, ntz, ntiles, per);
  else conv_c1in_v2_kernel<DT_F32><<<(unsigned)gx, W2_NT, 0, st>>>(a, ntx, nty, ntz, ntiles, per);
  ++g_tem_launches;
  return cudaGetLastError();
}

}  // namespace

bool conv_c1_supported(const ConvArgs& a) {
  if (a.C1 != 0 || a.bias) return false;
  for (int i = 0; i < 3; ++i) if (a.k[i] != 3 || a.stride[i] != 1) return false;
  const bool plain = !a.accumulate && !a.conv_off[0] && !a.conv_off[1] && !a.conv_off[2];
  // wide models (wf <= 4): the channel dimension is cut into slices of 16 (1 -> C) / 32 (C -> 1) handled by one launch each
  const bool co_ok = a.Cout == 8 || a.Cout == 16 || (a.Cout % 16 == 0 && !a.drop_key);
  const bool ci_ok = a.C0 == 8 || a.C0 == 16 || a.C0 == 32 || (a.C0 % 32 == 0 && a.slope == 1.f);
  if (plain && a.C0 == 1 && co_ok && a.out_dtype == DT_BF16 && a.out_C % 8 == 0 && a.out_coff % 8 == 0 &&
      (!a.ref || (a.ref_C % 8 == 0 && a.ref_coff % 8 == 0)))
    return true;                                                     // 1 -> C (form 0 forward, form 1 = flipped + pad 2)
  if ((a.form == 0 ? plain : true) && a.Cout == 1 && a.out_dtype == DT_F32 && a.s0.dtype == DT_BF16 && ci_ok &&
      a.s0.C % 8 == 0 && a.s0.coff % 8 == 0 && !a.s0.origins && !a.ref && !a.drop_key)
    return true;                                                     // C -> 1 forward; form 1 = data gradient of a 1 -> C layer
  return false;
}

cudaError_t launch_conv_c1(const ConvArgs& a, cudaStream_t st) {
  const int ntx = (a.L[2] + TXT - 1) / TXT, nty = (a.L[1] + TY - 1) / TY, ntz = (a.L[0] + TZ - 1) / TZ;
  const long long grid = (long long)a.B * ntx * nty * ntz;
  if (grid == 0) return cudaSuccess;
  if (a.C0 == 1 && a.Cout > 16) {                  // 1 -> C, wide: one launch per 16 output channels (a 32 B sector per voxel)
    for (int co0 = 0; co0 < a.Cout; co0 += 16) {
      ConvArgs s = a;
      s.Cout = 16; s.w = a.w + (long long)co0 * a.ws_out; s.out_coff = a.out_coff + co0; s.ref_coff = a.ref_coff + co0;
      const cudaError_t e = launch_conv_c1(s, st);
      if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
  }
  if (a.Cout == 1 && a.C0 > 32) {                  // C -> 1, wide (linear output): 32 input channels per launch, accumulated in fp32
    for (int ci0 = 0; ci0 < a.C0; ci0 += 32) {
      ConvArgs s = a;
      s.C0 = 32; s.w = a.w + (long long)ci0 * a.ws_in; s.s0.coff = a.s0.coff + ci0; s.accumulate = a.accumulate || ci0 > 0;
      const cudaError_t e = launch_conv_c1(s, st);
      if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
  }
  if (c1in_v2_ok(a)) return launch_c1in_v2(a, st);
  if (a.C0 == 1) {
    if (a.Cout == 8) conv_c1in_kernel<8><<<(unsigned)grid, 256, 0, st>>>(a, ntx, nty, ntz, a.form == 1);
    else conv_c1in_kernel<16><<<(unsigned)grid, 256, 0, st>>>(a, ntx, nty, ntz, a.form == 1);
  } else {
    const size_t smem = (size_t)HZc * HYc * HXc * (a.C0 / 8) * 16;
    static bool attr = false;
    if (!attr) {
      cudaError_t e = cudaFuncSetAttribute(conv_c1out_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); if (e) return e;
      e = cudaFuncSetAttribute(conv_c1out_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); if (e) return e;
      attr = true;
    }
    const int flip = a.form == 1;
    if (a.C0 == 8) conv_c1out_kernel<8><<<(unsigned)grid, 256, smem, st>>>(a, ntx, nty, ntz, flip);
    else if (a.C0 == 16) conv_c1out_kernel<16><<<(unsigned)grid, 256, smem, st>>>(a, ntx, nty, ntz, flip);
    else conv_c1out_kernel<32><<<(unsigned)grid, 256, smem, st>>>(a, ntx, nty, ntz, flip);
  }
  ++g_tem_launches;
  return cudaGetLastError();
}
