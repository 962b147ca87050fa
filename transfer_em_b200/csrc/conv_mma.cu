// Warp-level tensor-core convolution (mma.sync m16n8k16, bf16 -> fp32) for the layers the tcgen05 kernel does not
// take: stride-2 4x4x4 convolutions, transposed convolutions and their data gradients, 1x1 convolutions
// (models/utils.py:80,129-130; discriminator.py:78-80), any shape with Cin, Cout multiples of 8 (Cout <= 32).
//   form 0:  out[o] = sum_k in[s*o + k - pad] w[k]                      (strided conv fwd, convT dgrad)
//   form 1:  out[j] = sum_{k: (j+pad-k)%s==0} in[(j+pad-k)/s] w[k]      (conv dgrad, convT fwd): one CTA per output
//            parity class, so every class is a dense 2x2x2 (k=4,s=2) gather with warp-uniform taps.
// A rows are output positions: each ldmatrix row is the 16 B (8-channel) chunk of one input voxel in a cp.async-staged
// halo, so stride, taps and parity classes are only address arithmetic.  The 4x4x4 stride-2 layers whose weights fit
// beside the TMA ring now run on the tcgen05 kernels of conv_tc_s2.cu (TMA element strides do the space-to-depth
// split); this kernel keeps the 32 -> 32 stride-2 forward layers (d4, d6), the 1x1 layers and the 2-D shapes.
#include <string.h>
#include "tem_kernels.cuh"

extern unsigned long long g_tem_launches;

namespace {

constexpr int kWarpsC = 8;
constexpr int TXC = 16;
constexpr int kMaxTaps = 64;

struct CmArgs {
  const bf16* in; int IZ, IY, IX, Cin; int shift[3]; long long in_bstride;
  const float* w; long long ws_tap, ws_in, ws_out;
  int k[3], stride[3], pad[3], form;
  int B, L[3];
  bf16* out; int OZ, OY, OX, out_C, out_coff, out_off[3]; int Cout;
  float slope; const bf16* ref; int RZ, RY, RX, ref_C, ref_coff, ref_off[3]; float ref_slope;
  uint32_t drop_key; int accumulate;
  int TZ, TY;                 // tile: TZ x TY rows of 16 positions
  int HZ, HY, HX;             // staged halo extents
  int ntz, nty, ntx;          // tiles per sample (in class coordinates for form 1)
  int NB;                     // Cout / 8
  int kchunks;                // input-channel chunks (16 channels each; 1 when Cin == 8)
  int s_bytes, w_bytes;
};

__device__ __forceinline__ void cp16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int NBT>
__global__ void __launch_bounds__(kWarpsC * 32) conv_mma_kernel(const CmArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ int toff[kMaxTaps + 1];     // halo voxel offset of each local tap
  __shared__ int wtap[kMaxTaps + 1];     // global tap index of each local tap (-1 = zero weights)
  __shared__ int ntap_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbuf = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t wbuf = sbuf + a.s_bytes;

  // tile / parity class
  int t = blockIdx.x;
  const int tx = t % a.ntx; t /= a.ntx;
  const int ty = t % a.nty; t /= a.nty;
  const int tz = t % a.ntz; t /= a.ntz;
  const int b = t;
  int cls = blockIdx.z, par[3] = {0, 0, 0};
  if (a.form == 1) { par[2] = cls % a.stride[2]; cls /= a.stride[2]; par[1] = cls % a.stride[1]; cls /= a.stride[1]; par[0] = cls; }
  const int p0[3] = {tz * a.TZ, ty * a.TY, tx * TXC};     // first position of the tile (class coordinates for form 1)

  // per-axis tap lists and halo origin (input coordinates)
  int d0[3], cnt[3], org[3], amul[3];
  for (int ax = 0; ax < 3; ++ax) {
    if (a.form == 0) { d0[ax] = 0; cnt[ax] = a.k[ax]; amul[ax] = a.stride[ax]; org[ax] = p0[ax] * a.stride[ax] - a.pad[ax]; }
    else {
      const int s = a.stride[ax];
      d0[ax] = (par[ax] + a.pad[ax]) % s;
      cnt[ax] = (a.k[ax] > d0[ax]) ? (a.k[ax] - d0[ax] + s - 1) / s : 0;
      amul[ax] = 1;
      // position q (class coords) -> l = s*q + par, input i0 = (l + pad - d0)/s = q + (par + pad - d0)/s ; taps go i0 - m
      org[ax] = p0[ax] + (par[ax] + a.pad[ax] - d0[ax]) / s - (cnt[ax] - 1);
    }
  }
  {
    // one thread per local tap
    const int n = cnt[0] * cnt[1] * cnt[2];
    if (tid < n) {
      const int mx = tid % cnt[2], my = (tid / cnt[2]) % cnt[1], mz = tid / (cnt[2] * cnt[1]);
      int bz, by, bx, dz, dy, dx;
      if (a.form == 0) { bz = dz = mz; by = dy = my; bx = dx = mx; }
      else { dz = d0[0] + a.stride[0] * mz; dy = d0[1] + a.stride[1] * my; dx = d0[2] + a.stride[2] * mx;
             bz = cnt[0] - 1 - mz; by = cnt[1] - 1 - my; bx = cnt[2] - 1 - mx; }
      toff[tid] = (bz * a.HY + by) * a.HX + bx;
      wtap[tid] = (dz * a.k[1] + dy) * a.k[2] + dx;
      if (tid == n - 1) { toff[n] = toff[tid]; wtap[n] = -1; }     // dummy partner for an odd tap count (Cin == 8)
    }
    if (tid == 0) ntap_s = n;
  }
  __syncthreads();
  const int ntap = ntap_s;
  const bool cin8 = a.Cin == 8;
  const int ksteps = cin8 ? (ntap + 1) / 2 : ntap;        // K=16 steps per channel chunk
  const int hvox = a.HZ * a.HY * a.HX;
  const int npad = NBT * 8;

  constexpr int R = 4;                                    // rows of 16 positions per warp
  const int nrows = a.TZ * a.TY;
  float acc[R][NBT][4];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int n = 0; n < NBT; ++n) { acc[r][n][0] = acc[r][n][1] = acc[r][n][2] = acc[r][n][3] = 0.f; }

  const int a_px = (lane & 7) + 8 * ((lane >> 3) & 1), a_kh = lane >> 4;
  const int b_n = lane & 7, b_kh = (lane >> 3) & 1;
  const bf16* inb = a.in + (long long)b * a.in_bstride;

  for (int cc = 0; cc < a.kchunks; ++cc) {
    if (cc) __syncthreads();
    // ---- stage the input halo for this channel chunk as 8-channel planes
    const int planes = cin8 ? 1 : 2;
    {
      // row-wise: a halo row (fixed z,y) is HX voxels x Cin channels of contiguous global memory
      const int row_chunks = a.HX * planes;
      for (int row = warp; row < a.HZ * a.HY; row += kWarpsC) {
        const int hz = row / a.HY, hy = row - hz * a.HY;
        const int z = org[0] + hz + a.shift[0], y = org[1] + hy + a.shift[1], xb = org[2] + a.shift[2];
        const bool rowok = z >= 0 && z < a.IZ && y >= 0 && y < a.IY;
        const bf16* rp = inb + (((long long)z * a.IY + y) * a.IX + xb) * a.Cin + cc * 16;
        for (int j = lane; j < row_chunks; j += 32) {
          const int hx = cin8 ? j : (j >> 1), pl = cin8 ? 0 : (j & 1);
          const int x = xb + hx;
          const bool ok = rowok && x >= 0 && x < a.IX;
          cp16(sbuf + (uint32_t)(pl * hvox + row * a.HX + hx) * 16u, ok ? (const void*)(rp + (long long)hx * a.Cin + pl * 8) : (const void*)a.in, ok);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // ---- stage the weights of this chunk: [kstep][n][16 k] bf16 (k = 16 channels, or 2 taps x 8 channels)
#pragma unroll 4
    for (int i = tid; i < ksteps * npad * 16; i += kWarpsC * 32) {
      const int kk = i & 15; const int n = (i >> 4) % npad; const int ks = (i >> 4) / npad;
      int tap, ci;
      if (cin8) { const int tl = 2 * ks + (kk >> 3); tap = (tl <= ntap) ? wtap[tl] : -1; ci = kk & 7; }
      else { tap = wtap[ks]; ci = cc * 16 + kk; }
      float wv = 0.f;
      if (tap >= 0 && n < a.Cout) wv = a.w[tap * a.ws_tap + (long long)ci * a.ws_in + (long long)n * a.ws_out];
      reinterpret_cast<bf16*>(smem + a.s_bytes)[i] = __float2bfloat16_rn(wv);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // ---- MMAs
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = warp + r * kWarpsC;
      if (row >= nrows) break;
      const int pz = row / a.TY, py = row % a.TY;
      const int rowbase = (amul[0] * pz * a.HY + amul[1] * py) * a.HX + amul[2] * a_px;
      for (int ks = 0; ks < ksteps; ++ks) {
        const int tl = cin8 ? 2 * ks + a_kh : ks;
        const int pl = cin8 ? 0 : a_kh;
        uint32_t a0, a1, a2, a3;
        ldsm_x4(sbuf + (uint32_t)(pl * hvox + rowbase + toff[tl]) * 16u, a0, a1, a2, a3);
#pragma unroll
        for (int n = 0; n < NBT; ++n) {
          uint32_t b0, b1;
          ldsm_x2(wbuf + (uint32_t)(((ks * npad + n * 8 + b_n) * 16 + b_kh * 8) * 2), b0, b1);
          mma16816(acc[r][n], a0, a1, a2, a3, b0, b1);
        }
      }
    }
  }

  // ---- epilogue: c0,c1 = (px = g, channels 2t,2t+1), c2,c3 = (px = g+8, same channels)
  const int g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int row = warp + r * kWarpsC;
    if (row >= nrows) break;
    const int pz = row / a.TY, py = row % a.TY;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int px = g + 8 * hh;
      int l[3] = {p0[0] + pz, p0[1] + py, p0[2] + px};
      if (a.form == 1) for (int ax = 0; ax < 3; ++ax) l[ax] = l[ax] * a.stride[ax] + par[ax];
      if (l[0] >= a.L[0] || l[1] >= a.L[1] || l[2] >= a.L[2]) continue;
      const long long ovox = (((long long)b * a.OZ + l[0] + a.out_off[0]) * a.OY + l[1] + a.out_off[1]) * a.OX + l[2] + a.out_off[2];
      const long long rvox = (((long long)b * a.RZ + l[0] + a.ref_off[0]) * a.RY + l[1] + a.ref_off[1]) * a.RX + l[2] + a.ref_off[2];
      const uint32_t di = (uint32_t)(((((long long)b * a.L[0] + l[0]) * a.L[1] + l[1]) * a.L[2] + l[2]) * a.Cout);
#pragma unroll
      for (int n = 0; n < NBT; ++n) {
        const int co = n * 8 + 2 * tq;
        if (co >= a.Cout) continue;
        float v0 = acc[r][n][2 * hh], v1 = acc[r][n][2 * hh + 1];
        if (a.ref) {
          const __nv_bfloat162 rf = *reinterpret_cast<const __nv_bfloat162*>(a.ref + rvox * a.ref_C + a.ref_coff + co);
          v0 *= (__low2float(rf) > 0.f) ? 1.f : a.ref_slope;
          v1 *= (__high2float(rf) > 0.f) ? 1.f : a.ref_slope;
        }
        if (a.drop_key) { v0 *= 2.f * tem_keep(a.drop_key, di + co); v1 *= 2.f * tem_keep(a.drop_key, di + co + 1); }
        __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(a.out + ovox * a.out_C + a.out_coff + co);
        if (a.accumulate) { const __nv_bfloat162 o = *op; v0 += __low2float(o); v1 += __high2float(o); }
        if (a.slope != 1.f) { v0 = v0 > 0.f ? v0 : v0 * a.slope; v1 = v1 > 0.f ? v1 : v1 * a.slope; }
        *op = __floats2bfloat162_rn(v0, v1);
      }
    }
  }
}

}  // namespace

bool conv_mma_supported(const ConvArgs& a) {
  if (a.C1 != 0 || a.s0.dtype != DT_BF16 || a.out_dtype != DT_BF16) return false;
  if (a.s0.origins || a.use_lut || a.bias) return false;
  const int cin = a.C0;
  if (!(cin == 8 || (cin % 16 == 0 && cin >= 16))) return false;
  if (a.s0.C != cin || a.s0.coff != 0) return false;
  if (a.Cout % 8 || a.Cout > 32 || a.out_C % 2 || a.out_coff % 2) return false;
  if (a.ref && (a.ref_C % 2 || a.ref_coff % 2)) return false;
  if (a.conv_off[0] || a.conv_off[1] || a.conv_off[2]) return false;
  for (int i = 0; i < 3; ++i) if (a.stride[i] > 2 || a.k[i] > 4) return false;
  return true;
}

cudaError_t launch_conv_mma(const ConvArgs& c, cudaStream_t st) {
  CmArgs a; memset(&a, 0, sizeof(a));
  a.in = (const bf16*)c.s0.p; a.IZ = c.s0.Z; a.IY = c.s0.Y; a.IX = c.s0.X; a.Cin = c.C0; a.in_bstride = c.s0.bstride;
  a.w = c.w; a.ws_tap = c.ws_tap; a.ws_in = c.ws_in; a.ws_out = c.ws_out; a.form = c.form; a.B = c.B;
  for (int i = 0; i < 3; ++i) { a.shift[i] = c.s0.shift[i]; a.k[i] = c.k[i]; a.stride[i] = c.stride[i]; a.pad[i] = c.pad[i]; a.L[i] = c.L[i];
                                a.out_off[i] = c.out_off[i]; a.ref_off[i] = c.ref_off[i]; }
  a.out = (bf16*)c.out; a.OZ = c.OZ; a.OY = c.OY; a.OX = c.OX; a.out_C = c.out_C; a.out_coff = c.out_coff; a.Cout = c.Cout;
  a.slope = c.slope; a.ref = c.ref; a.RZ = c.RZ; a.RY = c.RY; a.RX = c.RX; a.ref_C = c.ref_C; a.ref_coff = c.ref_coff; a.ref_slope = c.ref_slope;
  a.drop_key = c.drop_key; a.accumulate = c.accumulate;
  a.NB = c.Cout / 8; a.kchunks = (c.C0 == 8) ? 1 : c.C0 / 16;
  // positions per axis (class coordinates for form 1)
  int Q[3], cnt[3];
  for (int i = 0; i < 3; ++i) {
    Q[i] = (c.form == 1) ? (c.L[i] + c.stride[i] - 1) / c.stride[i] : c.L[i];
    cnt[i] = (c.form == 1) ? (c.k[i] + c.stride[i] - 1) / c.stride[i] : c.k[i];
    if (Q[i] <= 0) return cudaSuccess;
  }
  const int planes = (c.C0 == 8) ? 1 : 2;
  auto sizes = [&](int tz, int ty, int& hz, int& hy, int& hx, int& sb, int& wb) {
    if (c.form == 0) { hz = (tz - 1) * c.stride[0] + c.k[0]; hy = (ty - 1) * c.stride[1] + c.k[1]; hx = (TXC - 1) * c.stride[2] + c.k[2]; }
    else { hz = tz + cnt[0] - 1; hy = ty + cnt[1] - 1; hx = TXC + cnt[2] - 1; }
    sb = ((hz * hy * hx * planes * 16) + 127) & ~127;
    const int ntap = cnt[0] * cnt[1] * cnt[2];
    const int ksteps = (c.C0 == 8) ? (ntap + 1) / 2 : ntap;
    wb = ((ksteps * a.NB * 8 * 32) + 127) & ~127;
    return sb + wb;
  };
  int TZ = Q[0] >= 4 ? 4 : (Q[0] >= 2 ? 2 : 1), TY = Q[1] >= 8 ? 8 : (Q[1] >= 4 ? 4 : (Q[1] >= 2 ? 2 : 1));
  int hz, hy, hx, sb, wb;
  while (sizes(TZ, TY, hz, hy, hx, sb, wb) > 100 * 1024 && (TZ > 1 || TY > 1)) { if (TZ > 1) TZ >>= 1; else TY >>= 1; }
  // small problems: prefer more CTAs over taller tiles
  auto ntiles = [&](int tz, int ty) { return (long long)c.B * ((Q[0] + tz - 1) / tz) * ((Q[1] + ty - 1) / ty) * ((Q[2] + TXC - 1) / TXC); };
  while (ntiles(TZ, TY) * ((c.form == 1) ? c.stride[0] * c.stride[1] * c.stride[2] : 1) < 2 * 148 && (TZ > 1 || TY > 1)) { if (TZ > 1) TZ >>= 1; else TY >>= 1; }
  const int smem = sizes(TZ, TY, hz, hy, hx, sb, wb);
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  a.TZ = TZ; a.TY = TY; a.HZ = hz; a.HY = hy; a.HX = hx; a.s_bytes = sb; a.w_bytes = wb;
  a.ntz = (Q[0] + TZ - 1) / TZ; a.nty = (Q[1] + TY - 1) / TY; a.ntx = (Q[2] + TXC - 1) / TXC;
  const int ncls = (c.form == 1) ? c.stride[0] * c.stride[1] * c.stride[2] : 1;
  dim3 grid((unsigned)((long long)c.B * a.ntz * a.nty * a.ntx), 1, ncls);
  static bool attr[5] = {false, false, false, false, false};
#define LAUNCH_CM(NBT)                                                                                               \
  {                                                                                                                  \
    if (!attr[NBT]) { cudaError_t e = cudaFuncSetAttribute(conv_mma_kernel<NBT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); if (e) return e; attr[NBT] = true; } \
    conv_mma_kernel<NBT><<<grid, kWarpsC * 32, smem, st>>>(a);                                                       \
  }
  if (a.NB == 1) LAUNCH_CM(1) else if (a.NB == 2) LAUNCH_CM(2) else if (a.NB == 3) LAUNCH_CM(3) else LAUNCH_CM(4)
#undef LAUNCH_CM
  ++g_tem_launches;
  return cudaGetLastError();
}
