// Direct (CUDA-core, fp32 accumulate) convolution kernels: the shape-generic path.
//
// One gather-form kernel covers every convolution of the reference graph and its gradients
// (transfer_em/models/utils.py:73,80,129-130):
//   form 0:  out[o] = sum_k in[s*o + k - pad] * w[k]                      conv forward, convT data-gradient
//   form 1:  out[j] = sum_{k: (j+pad-k) % s == 0} in[(j+pad-k)/s] * w[k]  conv data-gradient, convT forward
// with: two-source input (fused crop-and-concat skip, generator.py:74-86), virtual zero padding
// (ZeroPadding3D of the fakes, cgan.py:161,170), uint8 input standardised through a 256-entry LUT
// (datasets.py:157-163,193-202), per-sample tile origins into one shared volume (utils.py:77-89),
// fused LeakyReLU / dropout / bias epilogue, and for backward the fused LeakyReLU' / dropout-mask
// multiply and accumulate-into-window.  It is also the low-channel kernel (Cin = 1 or Cout = 1 layers are
// memory-bound and never go to tensor cores).  The tcgen05 implicit-GEMM kernels (conv_tc.cu) take over
// the channel-heavy 3x3x3 layers when the shape allows.
#include "tem_kernels.cuh"
extern unsigned long long g_tem_launches;

namespace {

constexpr int kThreads = 128;
constexpr int kVPT = 4;   // voxels per thread (weights are reused across them)

__device__ __forceinline__ float load_scalar(const SrcView& S, long long off, const float* lut) {
  if (S.dtype == DT_U8) return lut[reinterpret_cast<const uint8_t*>(S.p)[off]];
  if (S.dtype == DT_BF16) return bf2f(reinterpret_cast<const bf16*>(S.p)[off]);
  return reinterpret_cast<const float*>(S.p)[off];
}

template <int CO_T, int CI_V>
__global__ void __launch_bounds__(kThreads) conv_direct_kernel(const ConvArgs a) {
  extern __shared__ float wsm[];     // [ntap][nci][CO_T]
  __shared__ float lut[256];
  const int tid = threadIdx.x;
  const int co0 = blockIdx.y * CO_T;
  const int ntap = a.k[0] * a.k[1] * a.k[2];

  if (a.use_lut) {
    for (int i = tid; i < 256; i += kThreads) lut[i] = tem_standardize((float)i, a.lut_mean, a.lut_std);
  }

  // parity class (form 1 with stride 2): all voxels of a block share the tap subset
  int cls = blockIdx.z;
  int par[3];
  par[2] = cls % a.stride[2]; cls /= a.stride[2];
  par[1] = cls % a.stride[1]; cls /= a.stride[1];
  par[0] = cls;
  if (a.form == 0) { par[0] = par[1] = par[2] = 0; }

  // per-axis tap iteration: d = d0 + step*m, input position = i0 + sgn*m
  int d0[3], step[3], cnt[3], sgn;
  if (a.form == 0) {
    sgn = 1;
    for (int ax = 0; ax < 3; ++ax) { d0[ax] = 0; step[ax] = 1; cnt[ax] = a.k[ax]; }
  } else {
    sgn = -1;
    for (int ax = 0; ax < 3; ++ax) {
      int s = a.stride[ax];
      int c0 = par[ax] + a.conv_off[ax] + a.pad[ax];
      d0[ax] = ((c0 % s) + s) % s;
      step[ax] = s;
      cnt[ax] = (a.k[ax] > d0[ax]) ? (a.k[ax] - d0[ax] + s - 1) / s : 0;
    }
  }

  bool valid[kVPT];
  int vb[kVPT], l[kVPT][3], i0[kVPT][3];
  const long long vbase = (long long)blockIdx.x * (kThreads * kVPT) + tid;
#pragma unroll
  for (int j = 0; j < kVPT; ++j) {
    long long v = vbase + (long long)j * kThreads;
    valid[j] = v < a.nvox;
    long long t = valid[j] ? v : 0;
    int qx = (int)(t % a.H[2]); t /= a.H[2];
    int qy = (int)(t % a.H[1]); t /= a.H[1];
    int qz = (int)(t % a.H[0]); t /= a.H[0];
    vb[j] = (int)t;
    int q[3] = {qz, qy, qx};
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
      int st = (a.form == 1) ? a.stride[ax] : 1;
      l[j][ax] = q[ax] * st + par[ax];
      if (l[j][ax] >= a.L[ax]) valid[j] = false;
      int c = l[j][ax] + a.conv_off[ax];
      if (a.form == 0) i0[j][ax] = c * a.stride[ax] - a.pad[ax];
      else i0[j][ax] = (c + a.pad[ax] - d0[ax]) / a.stride[ax];   // exact by construction of d0
    }
  }

  float acc[kVPT][CO_T];
#pragma unroll
  for (int j = 0; j < kVPT; ++j)
#pragma unroll
    for (int c = 0; c < CO_T; ++c) acc[j][c] = 0.f;

  const int Ctot = a.C0 + a.C1;
  for (int cc = 0; cc < Ctot; cc += a.ci_chunk) {
    const bool second = cc >= a.C0;
    const SrcView& S = second ? a.s1 : a.s0;
    const int cbase = second ? cc - a.C0 : cc;
    const int climit = second ? Ctot : a.C0;
    const int nci = min(a.ci_chunk, climit - cc);
    __syncthreads();
    for (int i = tid; i < ntap * nci * CO_T; i += kThreads) {
      int co = i % CO_T; int r = i / CO_T; int ci = r % nci; int tap = r / nci;
      float wv = 0.f;
      if (co0 + co < a.Cout) wv = a.w[tap * a.ws_tap + (long long)(cc + ci) * a.ws_in + (long long)(co0 + co) * a.ws_out];
      wsm[i] = bf2f(__float2bfloat16_rn(wv));   // every conv kernel multiplies with the bf16 weight shadow
    }
    __syncthreads();

    for (int mz = 0; mz < cnt[0]; ++mz) {
      const int dz = d0[0] + step[0] * mz;
      for (int my = 0; my < cnt[1]; ++my) {
        const int dy = d0[1] + step[1] * my;
        for (int mx = 0; mx < cnt[2]; ++mx) {
          const int dx = d0[2] + step[2] * mx;
          const int tap = (dz * a.k[1] + dy) * a.k[2] + dx;
          const float* wt = wsm + (size_t)tap * nci * CO_T;
          long long off[kVPT]; bool inb[kVPT];
#pragma unroll
          for (int j = 0; j < kVPT; ++j) {
            int tz = i0[j][0] + sgn * mz + S.shift[0];
            int ty = i0[j][1] + sgn * my + S.shift[1];
            int tx = i0[j][2] + sgn * mx + S.shift[2];
            long long base;
            if (S.origins) {
              tz += S.origins[vb[j] * 3 + 0]; ty += S.origins[vb[j] * 3 + 1]; tx += S.origins[vb[j] * 3 + 2];
              base = 0;
            } else {
              base = (long long)vb[j] * S.bstride;
            }
            inb[j] = valid[j] && tz >= 0 && tz < S.Z && ty >= 0 && ty < S.Y && tx >= 0 && tx < S.X;
            off[j] = base + (((long long)tz * S.Y + ty) * S.X + tx) * S.C + S.coff + cbase;
          }
          if (CI_V == 8) {
            for (int ci = 0; ci < nci; ci += 8) {
              float xv[kVPT][8];
#pragma unroll
              for (int j = 0; j < kVPT; ++j) {
                uint4 q = make_uint4(0, 0, 0, 0);
                if (inb[j]) q = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(S.p) + off[j] + ci));
                unpack8(q, xv[j]);
              }
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                float wv[CO_T];
                const float* wr = wt + (ci + u) * CO_T;
                if (CO_T % 4 == 0) {
#pragma unroll
                  for (int c = 0; c < CO_T; c += 4) {
                    float4 w4 = *reinterpret_cast<const float4*>(wr + c);
                    wv[c] = w4.x; wv[c + 1] = w4.y; wv[c + 2] = w4.z; wv[c + 3] = w4.w;
                  }
                } else {
#pragma unroll
                  for (int c = 0; c < CO_T; ++c) wv[c] = wr[c];
                }
#pragma unroll
                for (int j = 0; j < kVPT; ++j)
#pragma unroll
                  for (int c = 0; c < CO_T; ++c) acc[j][c] = fmaf(xv[j][u], wv[c], acc[j][c]);
              }
            }
          } else {
            for (int ci = 0; ci < nci; ++ci) {
              float xs[kVPT];
#pragma unroll
              for (int j = 0; j < kVPT; ++j) {
                if (inb[j]) xs[j] = load_scalar(S, off[j] + ci, lut);
                else xs[j] = (a.use_lut && S.origins && valid[j]) ? lut[0] : 0.f;   // u8 tiles read 0 outside the volume
              }
              const float* wr = wt + ci * CO_T;
#pragma unroll
              for (int c = 0; c < CO_T; ++c) {
                float wv = wr[c];
#pragma unroll
                for (int j = 0; j < kVPT; ++j) acc[j][c] = fmaf(xs[j], wv, acc[j][c]);
              }
            }
          }
        }
      }
    }
  }

  // ---- epilogue ----
#pragma unroll
  for (int j = 0; j < kVPT; ++j) {
    if (!valid[j]) continue;
    const int lz = l[j][0], ly = l[j][1], lx = l[j][2];
    float v[CO_T];
#pragma unroll
    for (int c = 0; c < CO_T; ++c) {
      v[c] = acc[j][c];
      if (a.bias && co0 + c < a.Cout) v[c] += a.bias[co0 + c];
    }
    if (a.ref) {
      long long ro = ((((long long)vb[j] * a.RZ + lz + a.ref_off[0]) * a.RY + ly + a.ref_off[1]) * a.RX + lx + a.ref_off[2]) * a.ref_C + a.ref_coff + co0;
#pragma unroll
      for (int c = 0; c < CO_T; ++c)
        if (co0 + c < a.Cout) v[c] *= (bf2f(a.ref[ro + c]) > 0.f) ? 1.f : a.ref_slope;
    }
    if (a.drop_key) {
      uint32_t di = (uint32_t)(((((long long)vb[j] * a.L[0] + lz) * a.L[1] + ly) * a.L[2] + lx) * a.Cout + co0);
#pragma unroll
      for (int c = 0; c < CO_T; ++c) v[c] *= 2.f * tem_keep(a.drop_key, di + c);
    }
    long long oo = ((((long long)vb[j] * a.OZ + lz + a.out_off[0]) * a.OY + ly + a.out_off[1]) * a.OX + lx + a.out_off[2]) * a.out_C + a.out_coff + co0;
    if (a.out_dtype == DT_F32) {
      float* op = reinterpret_cast<float*>(a.out) + oo;
#pragma unroll
      for (int c = 0; c < CO_T; ++c) {
        if (co0 + c >= a.Cout) break;
        float r = v[c];
        if (a.accumulate) r += op[c];
        if (a.slope != 1.f) r = r > 0.f ? r : r * a.slope;
        op[c] = r;
      }
    } else {
      bf16* op = reinterpret_cast<bf16*>(a.out) + oo;
      const bool vec = (CO_T % 8 == 0) && ((oo & 7) == 0) && (co0 + CO_T <= a.Cout);
      if (vec) {
#pragma unroll
        for (int c = 0; c < CO_T; c += 8) {
          float r[8];
          if (a.accumulate) {
            uint4 q = *reinterpret_cast<const uint4*>(op + c);
            unpack8(q, r);
          } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) r[u] = 0.f;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            r[u] += v[c + u];
            if (a.slope != 1.f) r[u] = r[u] > 0.f ? r[u] : r[u] * a.slope;
          }
          uint4 o;
          o.x = pack2(r[0], r[1]); o.y = pack2(r[2], r[3]); o.z = pack2(r[4], r[5]); o.w = pack2(r[6], r[7]);
          *reinterpret_cast<uint4*>(op + c) = o;
        }
      } else {
#pragma unroll
        for (int c = 0; c < CO_T; ++c) {
          if (co0 + c >= a.Cout) break;
          float r = v[c];
          if (a.accumulate) r += bf2f(op[c]);
          if (a.slope != 1.f) r = r > 0.f ? r : r * a.slope;
          op[c] = __float2bfloat16_rn(r);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// weight gradient:  dw[tap][ca][cb] += sum_{b,p} S[b, s*p + tap - pad][ca] * P[b,p][cb]
// One warp owns one (tap, ca-block, cb-block) combination over the block's voxel range; lanes
// stride over voxels (coalesced), partial sums are shuffled down and added with one atomic per
// element per block.
// ---------------------------------------------------------------------------------------------
template <int CA_T, int CB_T>
__global__ void __launch_bounds__(256) wgrad_direct_kernel(const WgradArgs a) {
  __shared__ float lut[256];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (a.use_lut) {
    for (int i = tid; i < 256; i += 256) lut[i] = tem_standardize((float)i, a.lut_mean, a.lut_std);
    __syncthreads();
  }
  const int combo = blockIdx.y * 8 + warp;
  if (combo >= a.ncombo) return;
  const int nb = (a.Cb + CB_T - 1) / CB_T, na = (a.Ca + CA_T - 1) / CA_T;
  int r = combo;
  const int bblk = r % nb; r /= nb;
  const int ablk = r % na; r /= na;
  const int tap = r;
  const int dx = tap % a.k[2], dy = (tap / a.k[2]) % a.k[1], dz = tap / (a.k[2] * a.k[1]);

  float acc[CA_T][CB_T];
#pragma unroll
  for (int i = 0; i < CA_T; ++i)
#pragma unroll
    for (int j = 0; j < CB_T; ++j) acc[i][j] = 0.f;

  const long long v0 = (long long)blockIdx.x * a.vox_per_cta;
  const long long v1 = min(v0 + a.vox_per_cta, a.nvox);
  const SrcView& S = a.S;
  for (long long v = v0 + lane; v < v1; v += 32) {
    long long t = v;
    const int lx = (int)(t % a.L[2]); t /= a.L[2];
    const int ly = (int)(t % a.L[1]); t /= a.L[1];
    const int lz = (int)(t % a.L[0]); t /= a.L[0];
    const int b = (int)t;
    int tz = lz * a.stride[0] + dz - a.pad[0] + S.shift[0];
    int ty = ly * a.stride[1] + dy - a.pad[1] + S.shift[1];
    int tx = lx * a.stride[2] + dx - a.pad[2] + S.shift[2];
    long long sbase;
    if (S.origins) { tz += S.origins[b * 3]; ty += S.origins[b * 3 + 1]; tx += S.origins[b * 3 + 2]; sbase = 0; }
    else sbase = (long long)b * S.bstride;
    const bool inb = tz >= 0 && tz < S.Z && ty >= 0 && ty < S.Y && tx >= 0 && tx < S.X;
    float sa[CA_T];
    if (!inb) {
      const float fill = (a.use_lut && S.origins) ? lut[0] : 0.f;
#pragma unroll
      for (int i = 0; i < CA_T; ++i) sa[i] = fill;
      if (fill == 0.f) continue;
    } else {
      const long long so = sbase + (((long long)tz * S.Y + ty) * S.X + tx) * S.C + S.coff + ablk * CA_T;
      if (CA_T == 8) {
        uint4 q = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(S.p) + so));
        unpack8(q, sa);
      } else {
        sa[0] = load_scalar(S, so, lut);
      }
    }
    const long long po = (long long)b * a.p_bstride +
        ((((long long)lz + a.p_off[0]) * a.PY + ly + a.p_off[1]) * a.PX + lx + a.p_off[2]) * a.p_C + a.p_coff + bblk * CB_T;
    float pb[CB_T];
    if (CB_T == 8) {
      if (a.p_dtype == DT_BF16) {
        uint4 q = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(a.P) + po));
        unpack8(q, pb);
      } else {
        const float4* fp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.P) + po);
        float4 x0 = fp[0], x1 = fp[1];
        pb[0] = x0.x; pb[1] = x0.y; pb[2] = x0.z; pb[3] = x0.w; pb[4] = x1.x; pb[5] = x1.y; pb[6] = x1.z; pb[7] = x1.w;
      }
    } else {
      pb[0] = (a.p_dtype == DT_BF16) ? bf2f(reinterpret_cast<const bf16*>(a.P)[po]) : reinterpret_cast<const float*>(a.P)[po];
    }
#pragma unroll
    for (int i = 0; i < CA_T; ++i)
#pragma unroll
      for (int j = 0; j < CB_T; ++j) acc[i][j] = fmaf(sa[i], pb[j], acc[i][j]);
  }

#pragma unroll
  for (int i = 0; i < CA_T; ++i)
#pragma unroll
    for (int j = 0; j < CB_T; ++j) {
      float s = acc[i][j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == ((i * CB_T + j) & 31)) {
        const int ca = ablk * CA_T + i, cb = bblk * CB_T + j;
        if (ca < a.Ca && cb < a.Cb && s != 0.f)
          atomicAdd(a.dw + tap * a.ws_tap + (long long)ca * a.ws_a + (long long)cb * a.ws_b, s);
      }
    }
}

__global__ void bias_grad_kernel(const void* P, int p_dtype, long long nvox, int C, float* db) {
  // one block per channel group; tiny tensors only (d8 bias: discriminator.py:97-99)
  const int c = blockIdx.x;
  float s = 0.f;
  for (long long v = threadIdx.x; v < nvox; v += blockDim.x)
    s += (p_dtype == DT_BF16) ? bf2f(reinterpret_cast<const bf16*>(P)[v * C + c]) : reinterpret_cast<const float*>(P)[v * C + c];
  __shared__ float red[32];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    atomicAdd(db + c, t);
  }
}

template <int CO_T, int CI_V>
cudaError_t launch_conv_t(const ConvArgs& a, cudaStream_t st) {
  const int ntap = a.k[0] * a.k[1] * a.k[2];
  const size_t smem = (size_t)ntap * a.ci_chunk * CO_T * sizeof(float);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_direct_kernel<CO_T, CI_V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    configured = 200 * 1024;
  }
  int ncls = 1;
  if (a.form == 1) ncls = a.stride[0] * a.stride[1] * a.stride[2];
  dim3 grid((unsigned)((a.nvox + kThreads * kVPT - 1) / (kThreads * kVPT)), (a.Cout + CO_T - 1) / CO_T, ncls);
  conv_direct_kernel<CO_T, CI_V><<<grid, kThreads, smem, st>>>(a); ++g_tem_launches;
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_conv_direct(const ConvArgs& a_in, cudaStream_t st) {
  ConvArgs a = a_in;
  const int Ctot = a.C0 + a.C1;
  const bool vec_in = (a.s0.dtype == DT_BF16) && (a.C0 % 8 == 0) && (a.s0.C % 8 == 0) && (a.s0.coff % 8 == 0) &&
                      (a.C1 == 0 || ((a.s1.dtype == DT_BF16) && a.C1 % 8 == 0 && a.s1.C % 8 == 0 && a.s1.coff % 8 == 0));
  const int civ = vec_in ? 8 : 1;
  int cot = (a.Cout >= 16 && a.Cout % 16 == 0) ? 16 : (a.Cout > 1 ? 8 : 1);
  const int ntap = a.k[0] * a.k[1] * a.k[2];
  // input-channel chunk: weights for one chunk stay in shared memory (<= 64 KB)
  int chunk = civ;
  const int unit = civ;
  int cmax = (a.C1 > 0) ? a.C0 : Ctot;     // a chunk never straddles the two sources
  for (int c = unit; c <= cmax; c += unit) {
    if (cmax % c != 0) continue;
    if ((size_t)ntap * c * cot * sizeof(float) <= 64 * 1024) chunk = c;
  }
  if (a.C1 > 0 && a.C1 % chunk != 0) chunk = unit;
  a.ci_chunk = chunk;
  for (int ax = 0; ax < 3; ++ax) a.H[ax] = (a.form == 1) ? (a.L[ax] + a.stride[ax] - 1) / a.stride[ax] : a.L[ax];
  a.nvox = (long long)a.B * a.H[0] * a.H[1] * a.H[2];
  if (a.nvox == 0) return cudaSuccess;
  if (civ == 8) {
    if (cot == 16) return launch_conv_t<16, 8>(a, st);
    if (cot == 8) return launch_conv_t<8, 8>(a, st);
    return launch_conv_t<1, 8>(a, st);
  }
  if (cot == 16) return launch_conv_t<16, 1>(a, st);
  if (cot == 8) return launch_conv_t<8, 1>(a, st);
  return launch_conv_t<1, 1>(a, st);
}

cudaError_t launch_wgrad_direct(const WgradArgs& a_in, cudaStream_t st) {
  WgradArgs a = a_in;
  const bool va = (a.S.dtype == DT_BF16) && a.Ca % 8 == 0 && a.S.C % 8 == 0 && a.S.coff % 8 == 0;
  const bool vb = a.Cb % 8 == 0 && a.p_C % 8 == 0 && a.p_coff % 8 == 0;
  const int cat = va ? 8 : 1, cbt = vb ? 8 : 1;
  const int ntap = a.k[0] * a.k[1] * a.k[2];
  a.ncombo = ntap * ((a.Ca + cat - 1) / cat) * ((a.Cb + cbt - 1) / cbt);
  a.nvox = (long long)a.B * a.L[0] * a.L[1] * a.L[2];
  if (a.nvox == 0) return cudaSuccess;
  const int gy = (a.ncombo + 7) / 8;
  long long gx = (2 * 148 + gy - 1) / gy;
  if (gx < 1) gx = 1;
  long long per = (a.nvox + gx - 1) / gx;
  if (per < 256) per = 256;
  per = (per + 31) / 32 * 32;
  gx = (a.nvox + per - 1) / per;
  a.vox_per_cta = per;
  dim3 grid((unsigned)gx, gy);
  if (va && vb) wgrad_direct_kernel<8, 8><<<grid, 256, 0, st>>>(a);
  else if (va) wgrad_direct_kernel<8, 1><<<grid, 256, 0, st>>>(a);
  else if (vb) wgrad_direct_kernel<1, 8><<<grid, 256, 0, st>>>(a);
  else wgrad_direct_kernel<1, 1><<<grid, 256, 0, st>>>(a);
  ++g_tem_launches;
  return cudaGetLastError();
}

cudaError_t launch_bias_grad(const void* P, int p_dtype, long long nvox, int C, float* db, cudaStream_t st) {
  bias_grad_kernel<<<C, 256, 0, st>>>(P, p_dtype, nvox, C, db); ++g_tem_launches;
  return cudaGetLastError();
}
