// Shared device/host definitions for the transfer_em_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>

typedef __nv_bfloat16 bf16;

enum { DT_U8 = 0, DT_BF16 = 1, DT_F32 = 2 };

// A view of a dense channels-last tensor [B,Z,Y,X,C] (C = channel pitch).  `shift` maps the logical
// coordinate system of an op to tensor coordinates (negative = virtual zero padding, positive = crop
// window).  `origins` (optional) gives a per-sample (z,y,x) origin into one shared volume
// (tiled inference: transfer_em/utils.py:77-89 without the gather copy).
struct SrcView {
  const void* p;
  const int* origins;
  long long bstride;   // elements between samples (ignored when origins != nullptr)
  int dtype;
  int Z, Y, X;
  int C, coff;
  int shift[3];
};

struct ConvArgs {
  SrcView s0, s1;      // s1: second source of a fused crop-and-concat (generator.py:74-86)
  int C0, C1;          // input channels taken from s0 / s1
  const float* w;      // fp32 master weights
  long long ws_tap, ws_in, ws_out;   // element strides: tap, op-input channel, op-output channel
  const float* bias;
  int k[3], stride[3], pad[3];
  int form;            // 0: out[o] = sum_k in[s*o + k - pad] w[k]   (conv fwd, convT dgrad)
                       // 1: out[j] = sum_{k: (j+pad-k)%s==0} in[(j+pad-k)/s] w[k]  (conv dgrad, convT fwd)
  int B;
  int L[3];            // logical output extent computed by this launch
  int conv_off[3];     // conv coordinate = logical + conv_off
  void* out; int out_dtype;
  int OZ, OY, OX, out_C, out_coff, out_off[3];
  int Cout;
  float slope;         // forward LeakyReLU slope (1 = linear)
  const bf16* ref;     // backward: multiply by (ref > 0 ? 1 : ref_slope)
  int RZ, RY, RX, ref_C, ref_coff, ref_off[3];
  float ref_slope;
  uint32_t drop_key;   // != 0: multiply by 2*keep(hash(idx ^ key)), idx = dense [B,L,Cout] index
  int accumulate;      // out += result
  // merged data gradient of a crop-and-concat layer (generator.py:74-86 backwards): op-output channels [split, Cout) leave
  // to a second tensor with its own window and LeakyReLU' reference; dropout (dense [B,L,split] index) covers [0, split)
  // only.  split == 0: single destination.  Only the tcgen05 3x3x3 kernel implements it (tc_conv_supported)
  int split; void* out2; int O2Z, O2Y, O2X, out2_C, out2_off[3];
  const bf16* ref2; int R2Z, R2Y, R2X, ref2_C, ref2_off[3]; float ref2_slope;
  int use_lut; float lut_mean, lut_std;   // u8 input: (u/127.5 - 1 - mean)/std
  // fused inference epilogue of the Cout = 1 last generator layer (transfer_em/utils.py:109-121): un-standardise, crop tpad,
  // round-half-even, uint8 wrap, scatter tile b to its place in the stitched volume.  st_out == nullptr: plain output
  uint8_t* st_out; const int* st_index; int st_tpad, st_od; float st_mean, st_std; long long st_OZ, st_OY, st_OX;
  int ci_chunk;
  long long nvox;      // voxels per parity class
  int H[3];            // per-class extents (ceil(L/stride) for form 1, L for form 0)
};

struct WgradArgs {
  SrcView S;           // tensor read at s*p + k - pad
  int Ca;
  SrcView S1; int Ca1; // Ca1 != 0: the last Ca1 of the Ca channels come from a second tensor (input of a crop-and-concat
                       // layer, generator.py:74-86); only wgrad_tc.cu implements it
  const void* P; int p_dtype;
  int PZ, PY, PX, p_C, p_coff, p_off[3];
  long long p_bstride;
  int Cb;
  int B, L[3];
  int k[3], stride[3], pad[3];
  float* dw; long long ws_tap, ws_a, ws_b;
  int use_lut; float lut_mean, lut_std;
  long long nvox;
  long long vox_per_cta;
  int ncombo;
};

__host__ __device__ __forceinline__ uint32_t tem_hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ float tem_keep(uint32_t key, uint32_t idx) {
  return (tem_hash32(idx ^ key) >> 31) ? 1.0f : 0.0f;
}

#ifdef __CUDACC__
__device__ __forceinline__ float tem_standardize(float u, float mean, float stdv) {
  // scale_tensor + standardize_population, op order preserved (datasets.py:157-163,193-202)
  float t = __fdiv_rn(u, 127.5f);
  t = __fsub_rn(t, 1.0f);
  t = __fsub_rn(t, mean);
  return __fdiv_rn(t, stdv);
}
__device__ __forceinline__ float bf2f(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ uint8_t tem_to_u8_round(float y, float mean, float stdv) {
  // (y*std + mean + 1) * 127.5 -> np.around -> astype(uint8) wrap   (utils.py:109,118; datasets.py:165-171), op order kept
  float v = __fmul_rn(y, stdv);
  v = __fadd_rn(v, mean);
  v = __fadd_rn(v, 1.0f);
  v = __fmul_rn(v, 127.5f);
  const float r = rintf(v);
  const long long q = (long long)r;
  return (uint8_t)(q & 0xFF);
}
__device__ __forceinline__ void unpack8(const uint4& q, float* f) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
// Packed fp32 FMA (sm_100: FFMA2).  A three-register FFMA issues every second cycle per scheduler on this part, the packed
// form retires two FMAs in the same slot; each lane is an IEEE fma.rn, i.e. bit-identical to fmaf().
__device__ __forceinline__ unsigned long long tem_pk2(float lo, float hi) {
  unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void tem_upk2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void tem_ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
#endif

// Stage-ablation knobs (TEM_S2_DBG bits: skip epilogue work / weight loads / input loads / stores; TEM_DEBUG_GEN_BWD:
// truncated backward) produce WRONG results by design.  They exist only in builds made with -DTEM_ABLATION
// (python -m transfer_em_b200.build --ablation); the shipped library ignores the environment variables.
static inline const char* tem_ablation_env(const char* name) {
#ifdef TEM_ABLATION
  return getenv(name);
#else
  (void)name; return nullptr;
#endif
}
static inline int tem_ablation_bits() { const char* e = tem_ablation_env("TEM_S2_DBG"); return e ? atoi(e) : 0; }

// launchers (conv_direct.cu)
cudaError_t launch_conv_direct(const ConvArgs& a, cudaStream_t st);
cudaError_t launch_wgrad_direct(const WgradArgs& a, cudaStream_t st);
cudaError_t launch_bias_grad(const void* P, int p_dtype, long long nvox, int C, float* db, cudaStream_t st);

// conv_tc3.cu: tcgen05 implicit-GEMM path for 3x3x3 stride-1 convolutions, forward and data gradient (three kz taps
// per MMA, TMEM-resident accumulator strip)
bool tc_conv_supported(const ConvArgs& a);
size_t tc3_packed_bytes(int cin, int cout);
cudaError_t tc3_pack_weights(const ConvArgs& a, bf16* dst, cudaStream_t st);
cudaError_t launch_conv_tc3(const ConvArgs& a, const bf16* wpacked, cudaStream_t st);

// conv_tcw.cu: tcgen05 kernel for wide 3x3x3 stride-1 layers (Cin >= 64, 32-channel output groups, Cin swept in 32-channel passes)
bool tcw_conv_supported(const ConvArgs& a);
size_t tcw_packed_bytes(int cin, int cout);
cudaError_t tcw_pack_weights(const ConvArgs& a, bf16* dst, cudaStream_t st);
cudaError_t launch_conv_tcw(const ConvArgs& a, const bf16* wpacked, cudaStream_t st);

// conv_tc_s2.cu: tcgen05 kernels for the 4x4x4 stride-2 layers (strided conv / transposed conv, forward and data gradient)
bool tc_s2_supported(const ConvArgs& a);
size_t tc_s2_packed_bytes(const ConvArgs& a);
cudaError_t tc_s2_pack_weights(const ConvArgs& a, bf16* dst, cudaStream_t st);
cudaError_t launch_conv_tc_s2(const ConvArgs& a, const bf16* wpacked, cudaStream_t st);
const char* tc_s2_kernel_name(const ConvArgs& a);   // resident-weight or wide (streamed-weight) variant chosen for the shape

// conv_c1tc.cu: tcgen05 kernel for the single-input-channel 3x3x3 layers (g0 / d0 forward, g11 data gradient): dx taps
// folded into Toeplitz weight matrices
bool c1tc_supported(const ConvArgs& a);
size_t c1tc_packed_bytes(const ConvArgs& a);
cudaError_t c1tc_pack_weights(const ConvArgs& a, bf16* dst, cudaStream_t st);
cudaError_t launch_conv_c1tc(const ConvArgs& a, const bf16* wimg, cudaStream_t st);

// conv_small.cu: strided 4x4x4 forward for small output volumes with K = 64*Cin (d4, d6)
bool conv_small_supported(const ConvArgs& a);
cudaError_t launch_conv_small(const ConvArgs& a, cudaStream_t st);

// conv_mma.cu: mma.sync convolution for stride-2 / transposed / 1x1 layers with channel counts that are multiples of 8
bool conv_mma_supported(const ConvArgs& a);
cudaError_t launch_conv_mma(const ConvArgs& a, cudaStream_t st);

// conv_c1.cu: dedicated kernels for the single-channel first / last layers (1 -> C and C -> 1, 3x3x3 stride 1)
bool conv_c1_supported(const ConvArgs& a);
cudaError_t launch_conv_c1(const ConvArgs& a, cudaStream_t st);

// wgrad_mma.cu: tensor-core (mma.sync) weight gradient for channel counts that are multiples of 8
bool wgrad_mma_supported(const WgradArgs& a);
cudaError_t launch_wgrad_mma(const WgradArgs& a, cudaStream_t st);

// wgrad_tc.cu: tcgen05 weight gradient of the 3x3x3 stride-1 convolutions (MN-major operands, rows of the tile as dy taps)
bool wgrad_tc_supported(const WgradArgs& a);
cudaError_t launch_wgrad_tc(const WgradArgs& a, cudaStream_t st);

// wgrad_tc_s2.cu: tcgen05 weight gradient of the 4x4x4 stride-2 layers (parity classes = 2x2x2-tap stride-1 correlations)
bool wgrad_tc_s2_supported(const WgradArgs& a);
cudaError_t launch_wgrad_tc_s2(const WgradArgs& a, cudaStream_t st);

// wgrad_tcw.cu: tcgen05 weight gradient of the wide 3x3x3 stride-1 layers (channel planes fill M, one dz per CTA)
bool wgrad_tcw_supported(const WgradArgs& a);
cudaError_t launch_wgrad_tcw(const WgradArgs& a, cudaStream_t st);

// wgrad_c1.cu: weight gradient of the single-channel first / last layers
bool wgrad_c1_supported(const WgradArgs& a);
cudaError_t launch_wgrad_c1(const WgradArgs& a, cudaStream_t st);
const char* wgrad_c1_last_name();     // "wgrad_c1_kernel" (CUDA cores) or "wgrad_c1tc_kernel" (tcgen05) for the last launch_wgrad_c1 call
// wgrad_c1tc.cu: tcgen05 weight gradient of the single-channel layers (reached through launch_wgrad_c1)
bool wgrad_c1tc_supported(const WgradArgs& w);
cudaError_t launch_wgrad_c1tc(const WgradArgs& w, cudaStream_t st);

// disc_tail.cu: fused tail of the 3-D discriminator (d5 .. d8: 64 -> 1 voxels per sample), one CTA per sample
struct DiscTailArgs {
  int B, C4, e4, e5;                       // a4: [B, e4^3, C4];  a5: [B, e5^3, 32], e5 = e4 - 2 in {4, 5}
  const bf16* a4; bf16 *a5, *a6, *a7;      // stored activations (forward writes a5..a7, backward reads a4..a7)
  float* logits;                           // [B] fp32
  const float *w5, *w6, *w7, *w8, *b8;     // fp32 master weights, Keras layouts
  float slope4, slope5, slope6, slope7;    // LeakyReLU slopes of d4 .. d7 (d6: 0.3^2, discriminator.py:73-74)
  const float* dlogits;                    // backward: [B] fp32
  bf16* d_a4;                              // backward: gradient w.r.t. a4, LeakyReLU'(a4) applied
  float *dw5, *dw6, *dw7, *dw8, *db8;      // backward: weight-gradient accumulators (nullptr: data gradient only)
};
bool disc_tail_supported(const DiscTailArgs& a);
cudaError_t launch_disc_tail_fwd(const DiscTailArgs& a, cudaStream_t st);
cudaError_t launch_disc_tail_bwd(const DiscTailArgs& a, cudaStream_t st);

// elementwise.cu
cudaError_t launch_focal_logits(const float* x, long long n, float target, float gamma, float scale, int mode,
                                float* loss_out, float* grad, cudaStream_t st);
struct PairLossArgs {
  SrcView a;           // reference image window (u8 w/ LUT, or f32); shift = crop
  const float* b;      // generated fp32 dense [B, n, n, n, 1]
  int B, N[3];         // dims of b
  int crop[3];         // loss window = b[crop : N-crop]
  float gamma, scale;  // grad = scale * dl/db / count ; loss_out += scale * mean(l)
  int mode;            // 0 focal-probs (cgan.py:122-142), 1 L1
  int use_lut; float lut_mean, lut_std;
  float* loss_out; float* grad;   // grad dense like b (zeros outside the window)
};
cudaError_t launch_pair_loss(const PairLossArgs& a, cudaStream_t st);
cudaError_t launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr_t, float b1, float b2,
                        float eps, float gscale, cudaStream_t st);
cudaError_t launch_standardize_u8(const uint8_t* in, float* out, long long n, float mean, float stdv, cudaStream_t st);
cudaError_t launch_unstandardize_u8(const float* in, uint8_t* out, long long n, float mean, float stdv, cudaStream_t st);
cudaError_t launch_dropout_mask(uint32_t key, float* out, long long n, cudaStream_t st);
cudaError_t launch_cast_bf16_f32(const bf16* in, float* out, long long n, cudaStream_t st);
cudaError_t launch_cast_f32_bf16(const float* in, bf16* out, long long n, cudaStream_t st);
cudaError_t launch_init_normal(float* p, long long n, uint64_t seed, float stdv, cudaStream_t st);
struct AugmentArgs {     // datasets.py:123-155
  const void* in; int in_dtype;     // DT_U8 (scale + standardise fused in front) or DT_F32
  float* out;
  int B, n[3];                      // OUTPUT dims (z,y,x); 2-D data has n[0] = 1 and perm[0] = 0
  const int* perm;                  // device [B][3]: output axis k <- input axis perm[k]
  const int* flip;                  // device [B][3]
  const float* var_adj; const float* mean_adj;   // device [B]
  float mean, stdv;                 // standardisation of a uint8 source
};
cudaError_t launch_augment(const AugmentArgs& a, cudaStream_t st);
cudaError_t launch_mean_var(const float* x, long long n, double* scratch /* 2 doubles + 1 uint, zeroed */, float* out, cudaStream_t st);
cudaError_t launch_warp_tensor(const float* in, const float* uniform, float* out, int Z, int Y, int X, int nd, float rate, double* scratch, cudaStream_t st);
cudaError_t launch_chunk_volume(const uint8_t* vol, long long Z, long long Y, long long X, int c, uint8_t* out, cudaStream_t st);
struct StitchArgs {
  const float* y;      // generator output fp32 [T, od+2*tpad, .., 1]
  const int* index;    // [T][3] (x,y,z) output index of each tile (utils.py:84)
  int T, ydim, tpad, od;
  float mean, stdv;
  uint8_t* out; long long OZ, OY, OX;   // out[z,y,x], writes clipped to the request (utils.py:128-130)
};
cudaError_t launch_stitch_u8(const StitchArgs& a, cudaStream_t st);
struct FetchInArgs {
  const uint8_t* vol; long long VZ, VY, VX;
  const int* origins;  // [T][3] (z,y,x) tile origin in the volume
  const int* index;    // [T][3] (x,y,z)
  int T, buf, od; float mean, stdv;
  uint8_t* out; long long OZ, OY, OX;
};
cudaError_t launch_fetch_input_u8(const FetchInArgs& a, cudaStream_t st);
