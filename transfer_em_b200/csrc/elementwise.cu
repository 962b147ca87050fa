// Memory-bound helpers: fused loss kernels (cgan.py:110-142), multi-tensor Keras-Adam (cgan.py:69-73),
// uint8 <-> standardised-float conventions (datasets.py:157-171,193-202; utils.py:109-125), stitching.
#include "tem_kernels.cuh"
extern unsigned long long g_tem_launches;
#include <math.h>

namespace {

constexpr float kKerasEps = 1e-7f;

__device__ __forceinline__ float block_sum(float s) {
  __shared__ float red[32];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) red[w] = s;
  __syncthreads();
  float t = 0.f;
  if (w == 0) {
    t = (lane < (int)((blockDim.x + 31) >> 5)) ? red[lane] : 0.f;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;   // valid in thread 0
}

// tfa.losses.SigmoidFocalCrossEntropy(from_logits=True, alpha=.5, gamma) against a constant target
// (cgan.py:78-79,110-120); mode 1 = least-squares GAN.
__global__ void focal_logits_kernel(const float* x, long long n, float target, float gamma, float scale, int mode,
                                    float* loss_out, float* grad) {
  float s = 0.f;
  const float inv_n = 1.0f / (float)n;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float xv = x[i];
    float l, g;
    if (mode == 0) {
      const float z = target;
      const float ce = fmaxf(xv, 0.f) - xv * z + log1pf(expf(-fabsf(xv)));
      const float p = 1.f / (1.f + expf(-xv));
      const float pt = z * p + (1.f - z) * (1.f - p);
      const float omp = 1.f - pt;
      const float alpha_t = z * 0.5f + (1.f - z) * 0.5f;
      const float mod = powf(omp, gamma);
      l = alpha_t * mod * ce;
      const float dpt = (2.f * z - 1.f) * p * (1.f - p);
      const float dmod = (gamma == 2.f) ? 2.f * omp : gamma * powf(omp, gamma - 1.f);
      g = alpha_t * (-dmod * dpt * ce + mod * (p - z));
    } else {
      const float d = xv - target;
      l = d * d; g = 2.f * d;
    }
    s += l;
    if (grad) grad[i] = scale * g * inv_n;
  }
  s = block_sum(s);
  if (threadIdx.x == 0 && loss_out) atomicAdd(loss_out, scale * s * inv_n);
}

// identity / cycle loss (cgan.py:122-142): t = 1-|a-b|/2, focal CE of t against 1 through Keras
// binary_crossentropy (clip to [eps,1-eps], -log(. + eps)); mode 1 = L1 (docstring variants).
__global__ void pair_loss_kernel(const PairLossArgs a) {
  __shared__ float lut[256];
  if (a.use_lut) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = tem_standardize((float)i, a.lut_mean, a.lut_std);
    __syncthreads();
  }
  const long long per = (long long)a.N[0] * a.N[1] * a.N[2];
  const long long total = per * a.B;
  const long long cnt = (long long)a.B * (a.N[0] - 2 * a.crop[0]) * (a.N[1] - 2 * a.crop[1]) * (a.N[2] - 2 * a.crop[2]);
  const float inv = 1.0f / (float)cnt;
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int x = (int)(t % a.N[2]); t /= a.N[2];
    const int y = (int)(t % a.N[1]); t /= a.N[1];
    const int z = (int)(t % a.N[0]); t /= a.N[0];
    const int b = (int)t;
    const bool in = z >= a.crop[0] && z < a.N[0] - a.crop[0] && y >= a.crop[1] && y < a.N[1] - a.crop[1] &&
                    x >= a.crop[2] && x < a.N[2] - a.crop[2];
    float g = 0.f;
    if (in) {
      const long long ao = (long long)b * a.a.bstride +
          (((long long)(z + a.a.shift[0]) * a.a.Y + y + a.a.shift[1]) * a.a.X + x + a.a.shift[2]) * a.a.C + a.a.coff;
      const float av = (a.a.dtype == DT_U8) ? lut[reinterpret_cast<const uint8_t*>(a.a.p)[ao]]
                                            : reinterpret_cast<const float*>(a.a.p)[ao];
      const float d = av - a.b[i];
      const float sg = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
      if (a.mode == 0) {
        const float tc = 1.f - fabsf(d) * 0.5f;           // tconf
        const float omt = 1.f - tc;
        const float clipped = fminf(fmaxf(tc, kKerasEps), 1.f - kKerasEps);
        const float ce = -logf(clipped + kKerasEps);
        const float mod = powf(omt, a.gamma);
        s += 0.5f * mod * ce;
        const float dmod = (a.gamma == 2.f) ? 2.f * omt : a.gamma * powf(omt, a.gamma - 1.f);
        const float dce = (tc > kKerasEps && tc < 1.f - kKerasEps) ? -1.f / (clipped + kKerasEps) : 0.f;
        const float dl_dt = 0.5f * (-dmod * ce + mod * dce);
        g = a.scale * inv * dl_dt * sg * 0.5f;
      } else {
        s += fabsf(d);
        g = -a.scale * inv * sg;
      }
    }
    if (a.grad) a.grad[i] = g;
  }
  s = block_sum(s);
  if (threadIdx.x == 0 && a.loss_out) atomicAdd(a.loss_out, a.scale * s * inv);
}

// Keras Adam: var -= lr_t * m / (sqrt(v) + eps), lr_t = lr*sqrt(1-b2^t)/(1-b1^t) [upstream]
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr_t, float b1, float b2, float eps, float gscale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gv = g[i] * gscale;
    const float mv = b1 * m[i] + (1.f - b1) * gv;
    const float vv = b2 * v[i] + (1.f - b2) * gv * gv;
    m[i] = mv; v[i] = vv;
    p[i] = p[i] - lr_t * mv / (sqrtf(vv) + eps);
  }
}

__global__ void standardize_u8_kernel(const uint8_t* in, float* out, long long n, float mean, float stdv) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = tem_standardize((float)in[i], mean, stdv);
}

__global__ void unstandardize_u8_kernel(const float* in, uint8_t* out, long long n, float mean, float stdv) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = tem_to_u8_round(in[i], mean, stdv);
}

__global__ void dropout_mask_kernel(uint32_t key, float* out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = tem_keep(key, (uint32_t)i);
}

__global__ void cast_bf16_f32_kernel(const bf16* in, float* out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = bf2f(in[i]);
}
__global__ void cast_f32_bf16_kernel(const float* in, bf16* out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__global__ void init_normal_kernel(float* p, long long n, uint64_t seed, float stdv) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const uint64_t r = splitmix64(seed * 0x2545F4914F6CDD1Dull + (uint64_t)i);
    const float u1 = ((float)(uint32_t)(r >> 40) + 1.0f) * (1.0f / 16777217.0f);   // (0,1)
    const float u2 = (float)(uint32_t)((r >> 8) & 0xFFFFFF) * (1.0f / 16777216.0f);
    p[i] = stdv * sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
  }
}

__global__ void stitch_u8_kernel(const StitchArgs a) {
  const long long per = (long long)a.od * a.od * a.od;
  const long long total = per * a.T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int x = (int)(t % a.od); t /= a.od;
    const int y = (int)(t % a.od); t /= a.od;
    const int z = (int)(t % a.od); t /= a.od;
    const int tile = (int)t;
    const long long ox = a.index[tile * 3 + 0] + x, oy = a.index[tile * 3 + 1] + y, oz = a.index[tile * 3 + 2] + z;
    if (ox >= a.OX || oy >= a.OY || oz >= a.OZ) continue;
    const long long yo = (((long long)tile * a.ydim + z + a.tpad) * a.ydim + y + a.tpad) * a.ydim + x + a.tpad;
    a.out[(oz * a.OY + oy) * a.OX + ox] = tem_to_u8_round(a.y[yo], a.mean, a.stdv);
  }
}

__global__ void fetch_input_u8_kernel(const FetchInArgs a) {
  const long long per = (long long)a.od * a.od * a.od;
  const long long total = per * a.T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int x = (int)(t % a.od); t /= a.od;
    const int y = (int)(t % a.od); t /= a.od;
    const int z = (int)(t % a.od); t /= a.od;
    const int tile = (int)t;
    const long long ox = a.index[tile * 3 + 0] + x, oy = a.index[tile * 3 + 1] + y, oz = a.index[tile * 3 + 2] + z;
    if (ox >= a.OX || oy >= a.OY || oz >= a.OZ) continue;
    const long long vz = a.origins[tile * 3 + 0] + a.buf + z, vy = a.origins[tile * 3 + 1] + a.buf + y, vx = a.origins[tile * 3 + 2] + a.buf + x;
    uint8_t u = 0;
    if (vz >= 0 && vz < a.VZ && vy >= 0 && vy < a.VY && vx >= 0 && vx < a.VX) u = a.vol[(vz * a.VY + vy) * a.VX + vx];
    // standardise, un-standardise, (+1)*127.5, truncating cast (utils.py:122-125)
    float v = tem_standardize((float)u, a.mean, a.stdv);
    v = __fmul_rn(v, a.stdv); v = __fadd_rn(v, a.mean); v = __fadd_rn(v, 1.0f); v = __fmul_rn(v, 127.5f);
    a.out[(oz * a.OY + oy) * a.OX + ox] = (uint8_t)(((long long)v) & 0xFF);
  }
}

inline unsigned grid_for(long long n, int threads = 256, long long cap = 148LL * 16) {
  long long g = (n + threads - 1) / threads;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

}  // namespace

cudaError_t launch_focal_logits(const float* x, long long n, float target, float gamma, float scale, int mode,
                                float* loss_out, float* grad, cudaStream_t st) {
  focal_logits_kernel<<<1, 256, 0, st>>>(x, n, target, gamma, scale, mode, loss_out, grad); ++g_tem_launches;
  return cudaGetLastError();
}
cudaError_t launch_pair_loss(const PairLossArgs& a, cudaStream_t st) {
  const long long total = (long long)a.B * a.N[0] * a.N[1] * a.N[2];
  pair_loss_kernel<<<grid_for(total, 256, 148 * 4), 256, 0, st>>>(a); ++g_tem_launches;
  return cudaGetLastError();
}
cudaError_t launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr_t, float b1, float b2,
                        float eps, float gscale, cudaStream_t st) {
  adam_kernel<<<grid_for(n), 256, 0, st>>>(p, g, m, v, n, lr_t, b1, b2, eps, gscale); ++g_tem_launches;
  return cudaGetLastError();
}
// ---- chunked output (model_cloudrun/transferem.py:171-184): the zyx result volume re-tiled into c^3 blocks, every block
// contiguous (C order, clipped at the volume edge), blocks concatenated in the reference's z-outer / y / x-inner order ----
__global__ void chunk_volume_kernel(const uint8_t* vol, long long Z, long long Y, long long X, int c, uint8_t* out) {
  const long long nbx = (X + c - 1) / c, nby = (Y + c - 1) / c;
  long long blk = blockIdx.x;
  const long long bx = blk % nbx; blk /= nbx; const long long by = blk % nby; const long long bz = blk / nby;
  const long long z0 = bz * c, y0 = by * c, x0 = bx * c;
  const long long cz = min((long long)c, Z - z0), cy = min((long long)c, Y - y0), cx = min((long long)c, X - x0);
  // bytes before this block: full z-layers of blocks, full y-rows of blocks in this layer, blocks before it in its row
  const long long before = z0 * Y * X + cz * (y0 * X + cy * x0);
  uint8_t* dst = out + before;
  const long long n = cz * cy * cx;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const long long x = i % cx, r = i / cx, y = r % cy, z = r / cy;
    dst[i] = vol[((z0 + z) * Y + y0 + y) * X + x0 + x];
  }
}
cudaError_t launch_chunk_volume(const uint8_t* vol, long long Z, long long Y, long long X, int c, uint8_t* out, cudaStream_t st) {
  const long long nb = ((Z + c - 1) / c) * ((Y + c - 1) / c) * ((X + c - 1) / c);
  if (nb == 0) return cudaSuccess;
  chunk_volume_kernel<<<(unsigned)nb, 256, 0, st>>>(vol, Z, Y, X, c, out); ++g_tem_launches;
  return cudaGetLastError();
}

// ---- input conditioning on device (datasets.py:123-155 augment, :173-190 get_meanstd) ----
// augment: tf.transpose(perm) -> tf.reverse on the flipped axes -> *= var_adj -> += mean_adj, per sample; with a uint8
// source the scale_tensor + standardize_population steps that precede it in the reference pipeline are fused in front.
__global__ void augment_kernel(const AugmentArgs a) {
  const long long per = (long long)a.n[0] * a.n[1] * a.n[2];
  const long long total = per * a.B;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per); long long r = i % per;
    int o[3];
    o[2] = (int)(r % a.n[2]); r /= a.n[2]; o[1] = (int)(r % a.n[1]); o[0] = (int)(r / a.n[1]);
    // output axis k walks input axis perm[k]; a flipped output axis is read backwards
    int src[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int pk = a.perm[b * 3 + k];
      src[pk] = a.flip[b * 3 + k] ? a.n[k] - 1 - o[k] : o[k];
    }
    // input dims: in_n[perm[k]] = n[k]
    int in_n[3] = {1, 1, 1};
#pragma unroll
    for (int k = 0; k < 3; ++k) in_n[a.perm[b * 3 + k]] = a.n[k];
    const long long si = (long long)b * per + ((long long)src[0] * in_n[1] + src[1]) * in_n[2] + src[2];
    float v;
    if (a.in_dtype == DT_U8) v = tem_standardize((float)reinterpret_cast<const uint8_t*>(a.in)[si], a.mean, a.stdv);
    else v = reinterpret_cast<const float*>(a.in)[si];
    v = __fmul_rn(v, a.var_adj[b]);
    v = __fadd_rn(v, a.mean_adj[b]);
    a.out[i] = v;
  }
}
cudaError_t launch_augment(const AugmentArgs& a, cudaStream_t st) {
  const long long total = (long long)a.B * a.n[0] * a.n[1] * a.n[2];
  if (total == 0) return cudaSuccess;
  augment_kernel<<<grid_for(total), 256, 0, st>>>(a); ++g_tem_launches;
  return cudaGetLastError();
}

// tf.math.reduce_mean / reduce_variance of one fp32 tensor: fp64 sum and sum of squares, one atomic pair per block,
// finalised by the last block (out = {mean, population variance} as fp32)
__global__ void mean_var_kernel(const float* x, long long n, double* acc, unsigned int* done, float* out) {
  double s = 0.0, q = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = (double)x[i]; s += v; q += v * v;
  }
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  __shared__ double ss[8], sq[8];
  if ((threadIdx.x & 31) == 0) { ss[threadIdx.x >> 5] = s; sq[threadIdx.x >> 5] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { ts += ss[w]; tq += sq[w]; }
    atomicAdd(&acc[0], ts); atomicAdd(&acc[1], tq);
    __threadfence();
    if (atomicAdd(done, 1u) == gridDim.x - 1) {
      const double S = atomicAdd(&acc[0], 0.0), Q = atomicAdd(&acc[1], 0.0);
      const double m = S / (double)n;
      out[0] = (float)m; out[1] = (float)fmax(Q / (double)n - m * m, 0.0);
      acc[0] = 0.0; acc[1] = 0.0; *done = 0u;      // scratch is reusable by the next call on the stream
    }
  }
}
cudaError_t launch_mean_var(const float* x, long long n, double* scratch, float* out, cudaStream_t st) {
  unsigned int* done = reinterpret_cast<unsigned int*>(scratch + 2);
  int grid = (int)((n + 256 * 8 - 1) / (256 * 8)); if (grid < 1) grid = 1; if (grid > 148 * 8) grid = 148 * 8;
  mean_var_kernel<<<grid, 256, 0, st>>>(x, n, scratch, done, out); ++g_tem_launches;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// debug.py:7-63 warp_tensor(): 3^d box blur (conv 'SAME', zero padding, filter 1/27 or 1/9), then "holes": every voxel whose
// 4^d 'SAME' window (offsets -1..+2) holds a seed (uniform < rate) is set to the mean of the blurred tensor.
// Kernel 1 writes the blurred tensor and its fp64 sum, kernel 2 applies the holes in place.
// ------------------------------------------------------------------------------------------------
__global__ void warp_blur_kernel(const float* __restrict__ x, float* __restrict__ out, int Z, int Y, int X, int nd, double* acc) {
  const long long n = (long long)Z * Y * X;
  const float w = nd == 3 ? (1.0f / 27.0f) : (1.0f / 9.0f);
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % X); const long long r = i / X; const int yy = (int)(r % Y); const int zz = (int)(r / Y);
    float a = 0.f;
    for (int dz = (nd == 3 ? -1 : 0); dz <= (nd == 3 ? 1 : 0); ++dz) {
      const int z = zz + dz; if (z < 0 || z >= Z) continue;
      for (int dy = -1; dy <= 1; ++dy) {
        const int y = yy + dy; if (y < 0 || y >= Y) continue;
        for (int dx = -1; dx <= 1; ++dx) {
          const int q = xx + dx; if (q < 0 || q >= X) continue;
          a = fmaf(__ldg(x + ((long long)z * Y + y) * X + q), w, a);
        }
      }
    }
    out[i] = a; s += (double)a;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double ss[8];
  if ((threadIdx.x & 31) == 0) ss[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0.0; for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += ss[k]; atomicAdd(acc, t); }
}
__global__ void warp_holes_kernel(float* __restrict__ out, const float* __restrict__ u, int Z, int Y, int X, int nd, float rate, const double* acc) {
  const long long n = (long long)Z * Y * X;
  const float mean = (float)(acc[0] / (double)n);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % X); const long long r = i / X; const int yy = (int)(r % Y); const int zz = (int)(r / Y);
    bool hole = false;
    for (int dz = (nd == 3 ? -1 : 0); dz <= (nd == 3 ? 2 : 0) && !hole; ++dz) {
      const int z = zz + dz; if (z < 0 || z >= Z) continue;
      for (int dy = -1; dy <= 2 && !hole; ++dy) {
        const int y = yy + dy; if (y < 0 || y >= Y) continue;
        for (int dx = -1; dx <= 2; ++dx) {
          const int q = xx + dx; if (q < 0 || q >= X) continue;
          if (__ldg(u + ((long long)z * Y + y) * X + q) < rate) { hole = true; break; }
        }
      }
    }
    if (hole) out[i] = mean;
  }
}
cudaError_t launch_warp_tensor(const float* in, const float* uniform, float* out, int Z, int Y, int X, int nd, float rate, double* scratch, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(double), st); if (e) return e;
  const long long n = (long long)Z * Y * X;
  warp_blur_kernel<<<grid_for(n), 256, 0, st>>>(in, out, Z, Y, X, nd, scratch); ++g_tem_launches;
  warp_holes_kernel<<<grid_for(n), 256, 0, st>>>(out, uniform, Z, Y, X, nd, rate, scratch); ++g_tem_launches;
  return cudaGetLastError();
}

cudaError_t launch_standardize_u8(const uint8_t* in, float* out, long long n, float mean, float stdv, cudaStream_t st) {
  standardize_u8_kernel<<<grid_for(n), 256, 0, st>>>(in, out, n, mean, stdv); ++g_tem_launches;
  return cudaGetLastError();
}
cudaError_t launch_unstandardize_u8(const float* in, uint8_t* out, long long n, float mean, float stdv, cudaStream_t st) {
  unstandardize_u8_kernel<<<grid_for(n), 256, 0, st>>>(in, out, n, mean, stdv); ++g_tem_launches;
  return cudaGetLastError();
}
cudaError_t launch_dropout_mask(uint32_t key, float* out, long long n, cudaStream_t st) {
  dropout_mask_kernel<<<grid_for(n), 256, 0, st>>>(key, out, n); ++g_tem_launches;
  return cudaGetLastError();
}
cudaError_t launch_cast_bf16_f32(const bf16* in, float* out, long long n, cudaStream_t st) {
  cast_bf16_f32_kernel<<<grid_for(n), 256, 0, st>>>(in, out, n); ++g_tem_launches;
  return cudaGetLastError();
}
cudaError_t launch_cast_f32_bf16(const float* in, bf16* out, long long n, cudaStream_t st) {
  cast_f32_bf16_kernel<<<grid_for(n), 256, 0, st>>>(in, out, n); ++g_tem_launches;
  return cudaGetLastError();
}
cudaError_t launch_init_normal(float* p, long long n, uint64_t seed, float stdv, cudaStream_t st) {
  init_normal_kernel<<<grid_for(n), 256, 0, st>>>(p, n, seed, stdv); ++g_tem_launches;
  return cudaGetLastError();
}
cudaError_t launch_stitch_u8(const StitchArgs& a, cudaStream_t st) {
  stitch_u8_kernel<<<grid_for((long long)a.T * a.od * a.od * a.od), 256, 0, st>>>(a); ++g_tem_launches;
  return cudaGetLastError();
}
cudaError_t launch_fetch_input_u8(const FetchInArgs& a, cudaStream_t st) {
  fetch_input_u8_kernel<<<grid_for((long long)a.T * a.od * a.od * a.od), 256, 0, st>>>(a); ++g_tem_launches;
  return cudaGetLastError();
}
