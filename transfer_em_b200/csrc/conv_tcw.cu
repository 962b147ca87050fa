// tcgen05 implicit-GEMM 3x3x3 stride-1 convolution for WIDE layers (Cin, Cout multiples of 32: the wf <= 2 models of
// BASELINE config 4, generator.py:54-110 with 64/128/256 channels), forward and data gradient.
//
// conv_tc3.cu keeps the whole bf16 weight image of a layer in shared memory, which stops fitting at 64 x 64 channels
// (27*Cin*Cout*2 B).  This kernel keeps its structure - z-marching CTA over a 16 x 8 voxel column, TMA halo planes,
// three kz taps per MMA (N = 3 x 32 columns), TMEM-resident accumulator strip of 11 output slices - and adds two splits:
//   * output channels: blockIdx.y owns 32 output channels (N = 96 keeps the MMA above the shared-memory A-read bound:
//     48 tensor cycles against 4 KB + 3 KB of operand reads);
//   * input channels: the z-chunk is swept once per 32-channel chunk of Cin ("pass"); all passes accumulate into the
//     same TMEM strip, so nothing is spilled between passes.  A pass needs 55 KB of weights (9 taps x 2 k-steps x 96
//     columns); the weight buffer is double-buffered and the next pass's image is fetched during the current pass.
// Tensor-pipe time per (input slice, chunk) = 18 MMAs x 48 cycles; input tiles are re-read Cout/32 times (L2 hits:
// the 72^3 x 64-channel activation of one sample is 48 MB).  Epilogue identical to conv_tc3.cu, on a 32-channel window.
#include <cuda.h>
#include <string.h>
#include <stdlib.h>
#include "tem_kernels.cuh"
#include "ptx_sm100.cuh"

extern unsigned long long g_tem_launches;

namespace {

constexpr int TX = 8, TY = 16, HX = TX + 2, HY = TY + 2;
constexpr int PLANE_BYTES = HY * HX * 16;
constexpr int PLANE_STRIDE = 2944;
constexpr int WCH = 32;                                  // input channels per pass
constexpr int WPL = WCH / 8;                             // planes per ring slot
constexpr int SLOT_BYTES = WPL * PLANE_STRIDE;           // 11776
constexpr int CPW = 32, NPW = 3 * CPW;                   // output channels per CTA, MMA N
constexpr int WSTEPS = 9 * (WCH / 16);                   // k-steps per pass
constexpr int WBYTES = WSTEPS * NPW * 32;                // 55296: packed weights of one (cout group, cin chunk)
constexpr int RINGW = 8;
constexpr int ZCAP = 512 / CPW - 5;                      // 11 output slices per z-chunk
constexpr int kThreads = 192;

struct TcwArgs {
  int B, L[3];
  int nchunks, chunks0;        // Cin/32 passes; the first chunks0 come from map0, the rest from map1
  int shift0[3], shift1[3];
  const bf16* wpacked;         // [cout group][cin chunk][WBYTES]
  int ntx, nty, nzc;
  bf16* out; int OZ, OY, OX, out_C, out_coff, out_off[3];
  int Cout;
  float slope;
  const bf16* ref; int RZ, RY, RX, ref_C, ref_coff, ref_off[3]; float ref_slope;
  uint32_t drop_key;
  int accumulate;
  int swz;                     // input tiles as [voxel][32 ch] rows in the 64B-swizzle layout (one TMA request per voxel)
  int dbg;                     // experiment bits (TEM_S2_DBG): 1 no epilogue memory traffic, 2 no weight loads, 4 no input loads
};

__device__ __forceinline__ void tmem_st8_zero(uint32_t taddr) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(0u) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
conv3_tcw_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1, const TcwArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[RINGW], empty_bar[RINGW], wfull_bar[2], wempty_bar[2], tzero_bar, tfull_bar[ZCAP];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* wsm = smem;                                   // two weight buffers
  uint8_t* ring = smem + 2 * WBYTES;                     // 110592: multiple of 1024

  int w = blockIdx.x;
  const int zc_i = w % a.nzc; w /= a.nzc;
  const int tx_i = w % a.ntx; w /= a.ntx;
  const int ty_i = w % a.nty; w /= a.nty;
  const int b = w;
  const int cg = blockIdx.y;                             // group of 32 output channels
  const int x0 = tx_i * TX, y0 = ty_i * TY, z0 = zc_i * ZCAP;
  const int nz = min(ZCAP, a.L[0] - z0);
  const int nslices = nz + 2;

  if (threadIdx.x == 0) {
    for (int i = 0; i < RINGW; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&wfull_bar[i], 1); mbar_init(&wempty_bar[i], 1); }
    mbar_init(&tzero_bar, 128);
    for (int i = 0; i < nz; ++i) mbar_init(&tfull_bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      int slot = 0; uint32_t ph = 0;
      for (int c = 0; c < a.nchunks; ++c) {
        const int wb = c & 1;
        mbar_wait(&wempty_bar[wb], (((uint32_t)(c >> 1)) & 1u) ^ 1u);
        if (a.dbg & 2) mbar_arrive(&wfull_bar[wb]);
        else {
          mbar_arrive_expect_tx(&wfull_bar[wb], (uint32_t)WBYTES);
          bulk_load(wsm + wb * WBYTES, reinterpret_cast<const uint8_t*>(a.wpacked) + ((size_t)cg * a.nchunks + c) * WBYTES, (uint32_t)WBYTES, &wfull_bar[wb]);
        }
        const bool src1 = c >= a.chunks0;
        const CUtensorMap* mp = src1 ? &map1 : &map0;
        const int pl0 = (src1 ? c - a.chunks0 : c) * WPL;
        const int sx = x0 + (src1 ? a.shift1[2] : a.shift0[2]), sy = y0 + (src1 ? a.shift1[1] : a.shift0[1]), sz = z0 + (src1 ? a.shift1[0] : a.shift0[0]);
        for (int s = 0; s < nslices; ++s) {
          mbar_wait(&empty_bar[slot], ph ^ 1u);
          if (a.dbg & 4) mbar_arrive(&full_bar[slot]);
          else {
            mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)WPL * PLANE_BYTES);
            uint8_t* dst = ring + (size_t)slot * SLOT_BYTES;
            if (a.swz) tma_load_5d(dst, mp, &full_bar[slot], pl0 * 8, sx, sy, sz + s, b);
            else {
#pragma unroll
              for (int p = 0; p < WPL; ++p) tma_load_5d(dst + p * PLANE_STRIDE, mp, &full_bar[slot], (pl0 + p) * 8, sx, sy, sz + s, b);
            }
          }
          if (++slot == RINGW) { slot = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NPW >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    mbar_wait(&tzero_bar, 0);                       // accumulator strip has been zeroed
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t wb16 = smem_u32(wsm) >> 4;
    const uint32_t rbase = smem_u32(ring);
    // A operand: 8-channel planes (SBO = one halo row, LBO = next plane), or rows of 32 channels (64 B, SWIZZLE_64B as TMA
    // writes them; SBO = one halo row of 10 voxels, K-step = 32 B further in the row, tap = start shifted by whole rows)
    const uint32_t a_hi = a.swz ? (((uint32_t)(HX * 64) >> 4) | (1u << 14) | (4u << 29)) : (((uint32_t)(HX * 16) >> 4) | (1u << 14));
    const uint32_t b_hi = (128u >> 4) | (1u << 14);                         // SBO = 128 B between n-groups
    const uint32_t b_lbo = ((uint32_t)(NPW * 16) >> 4) << 16;
    const uint32_t a_lbo = a.swz ? (1u << 16) : (((uint32_t)PLANE_STRIDE >> 4) << 16);   // K halves = two consecutive planes
    int slot = 0; uint32_t ph = 0;
    for (int c = 0; c < a.nchunks; ++c) {
      const int wb = c & 1;
      mbar_wait(&wfull_bar[wb], ((uint32_t)(c >> 1)) & 1u);
      const bool last = c == a.nchunks - 1;
      for (int s = 0; s < nslices; ++s) {
        mbar_wait(&full_bar[slot], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)((ZCAP + 1 - s) * CPW);
        const uint32_t sb16 = (rbase + (uint32_t)slot * SLOT_BYTES) >> 4;
        if (elect_one()) {
          uint32_t blo = (wb16 + (uint32_t)(wb * WBYTES >> 4)) | b_lbo;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            uint32_t alo = (sb16 + (uint32_t)((t / 3) * HX + (t % 3)) * (a.swz ? 4u : 1u)) | a_lbo;
#pragma unroll
            for (int kc = 0; kc < WCH / 16; ++kc) {
              umma_bf16(d_tmem, ((uint64_t)a_hi << 32) | alo, ((uint64_t)b_hi << 32) | blo, idesc, 1u);
              alo += a.swz ? 2u : ((uint32_t)(2 * PLANE_STRIDE) >> 4);
              blo += (uint32_t)(NPW * 32) >> 4;
            }
          }
          umma_commit(&empty_bar[slot]);
          if (last && s >= 2) umma_commit(&tfull_bar[s - 2]);      // output slice s-2 is complete after the last pass
          if (s == nslices - 1) umma_commit(&wempty_bar[wb]);      // this pass's weights are no longer read
        }
        __syncwarp();
        if (++slot == RINGW) { slot = 0; ph ^= 1u; }
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int yl = row >> 3, xl = row & 7;
    const int oy = y0 + yl, ox = x0 + xl;
    const bool inside = oy < a.L[1] && ox < a.L[2] && !(a.dbg & 1);
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int c = 0; c < 512; c += 8) tmem_st8_zero(lane_base + (uint32_t)c);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    mbar_arrive(&tzero_bar);
    const int co0 = cg * CPW;
    const int ncol = min(CPW, a.Cout - co0);
    // the LeakyReLU' operand of the data gradient is fetched PF output slices ahead (see conv_tc3.cu)
    constexpr int PF = 3;
    uint4 refq[PF][CPW / 8];
    const long long ref_zstride = (long long)a.RY * a.RX * a.ref_C;
    const long long ref_base = ((((long long)b * a.RZ + z0 + a.ref_off[0]) * a.RY + oy + a.ref_off[1]) * a.RX + ox + a.ref_off[2]) * a.ref_C + a.ref_coff + co0;
    auto fetch_ref = [&](int zo, uint4* q) {
      if (a.ref && inside && zo < nz) {
#pragma unroll
        for (int c = 0; c < CPW / 8; ++c) if (c * 8 < ncol) q[c] = __ldg(reinterpret_cast<const uint4*>(a.ref + ref_base + (long long)zo * ref_zstride + c * 8));
      }
    };
#pragma unroll
    for (int u = 0; u < PF; ++u) fetch_ref(u, refq[u]);
    for (int zb = 0; zb < nz; zb += PF) {
#pragma unroll
    for (int pu = 0; pu < PF; ++pu) {
      const int zo = zb + pu;
      if (zo >= nz) break;
      const int oz = z0 + zo;
      mbar_wait(&tfull_bar[zo], 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r[CPW];
      const uint32_t taddr = lane_base + (uint32_t)((ZCAP + 1 - zo) * CPW);
#pragma unroll
      for (int c = 0; c < CPW; c += 8) tmem_ld8(taddr + c, r + c);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (!inside) continue;
      bf16* op = a.out + ((((long long)b * a.OZ + oz + a.out_off[0]) * a.OY + oy + a.out_off[1]) * a.OX + ox + a.out_off[2]) * a.out_C + a.out_coff + co0;
      const uint32_t di = (uint32_t)(((((long long)b * a.L[0] + oz) * a.L[1] + oy) * a.L[2] + ox) * a.Cout) + (uint32_t)co0;
#pragma unroll
      for (int c = 0; c < CPW; c += 8) {
        if (c < ncol) {
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = __uint_as_float(r[c + u]);
          if (a.ref) {
            float f[8];
            unpack8(refq[pu][c / 8], f);
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] *= (f[u] > 0.f) ? 1.f : a.ref_slope;
          }
          if (a.drop_key) {
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] *= 2.f * tem_keep(a.drop_key, di + c + u);
          }
          float o[8];
          if (a.accumulate) unpack8(*reinterpret_cast<const uint4*>(op + c), o);
          else {
#pragma unroll
            for (int u = 0; u < 8; ++u) o[u] = 0.f;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            o[u] += v[u];
            if (a.slope != 1.f) o[u] = o[u] > 0.f ? o[u] : o[u] * a.slope;
          }
          uint4 pk;
          pk.x = pack2(o[0], o[1]); pk.y = pack2(o[2], o[3]); pk.z = pack2(o[4], o[5]); pk.w = pack2(o[6], o[7]);
          *reinterpret_cast<uint4*>(op + c) = pk;
        }
      }
      fetch_ref(zo + PF, refq[pu]);
    }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// bf16 UMMA B image per (cout group g, cin chunk c): [step = tap9*2 + kc][k-half][n-group (12)][8 rows][8 elems],
// n = kz*32 + co, ci = c*32 + (2*kc + k-half)*8 + e
struct PackWArgs {
  const float* w; long long ws_tap, ws_in, ws_out;
  int flip, nchunks, cout;
  bf16* dst; long long total;
};
__global__ void pack_weights_w_kernel(const PackWArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.total) return;
  long long t = i;
  const int e = (int)(t & 7); t >>= 3;
  const int r = (int)(t & 7); t >>= 3;
  const int g12 = (int)(t % 12); t /= 12;
  const int j = (int)(t & 1); t >>= 1;
  const int step = (int)(t % WSTEPS); t /= WSTEPS;
  const int c = (int)(t % a.nchunks); t /= a.nchunks;
  const int cg = (int)t;
  const int tap9 = step >> 1, kc = step & 1;
  const int n = g12 * 8 + r, dz = n / CPW, col = n % CPW;
  const int ci = c * WCH + (2 * kc + j) * 8 + e, co = cg * CPW + col;
  float v = 0.f;
  if (co < a.cout) {
    int tap = dz * 9 + tap9;
    if (a.flip) tap = 26 - tap;
    v = a.w[(long long)tap * a.ws_tap + (long long)ci * a.ws_in + (long long)co * a.ws_out];
  }
  a.dst[i] = __float2bfloat16_rn(v);
}

}  // namespace

bool tcw_conv_supported(const ConvArgs& a) {
  if (a.form != 0 && a.form != 1) return false;
  for (int i = 0; i < 3; ++i) if (a.k[i] != 3 || a.stride[i] != 1 || a.conv_off[i]) return false;
  if (a.s0.dtype != DT_BF16 || a.out_dtype != DT_BF16 || a.s0.origins || a.use_lut || a.bias) return false;
  const int cin = a.C0 + a.C1;
  if (cin < WCH || a.C0 % WCH || a.C1 % WCH) return false;
  if (a.s0.C != a.C0 || a.s0.coff != 0) return false;
  if (a.C1 && (a.s1.dtype != DT_BF16 || a.s1.C != a.C1 || a.s1.coff != 0)) return false;
  if (a.Cout % 8 || a.Cout < 32 || a.out_C % 8 || a.out_coff % 8) return false;
  if (a.ref && (a.ref_C % 8 || a.ref_coff % 8)) return false;
  return tem_get_encode() != nullptr;
}

size_t tcw_packed_bytes(int cin, int cout) { return (size_t)((cout + CPW - 1) / CPW) * (cin / WCH) * WBYTES; }

cudaError_t tcw_pack_weights(const ConvArgs& a, bf16* dst, cudaStream_t st) {
  PackWArgs p;
  const int cin = a.C0 + a.C1;
  p.w = a.w; p.ws_tap = a.ws_tap; p.ws_in = a.ws_in; p.ws_out = a.ws_out;
  p.flip = (a.form == 1) ? 1 : 0; p.nchunks = cin / WCH; p.cout = a.Cout;
  p.dst = dst; p.total = (long long)(tcw_packed_bytes(cin, a.Cout) / 2);
  pack_weights_w_kernel<<<(unsigned)((p.total + 255) / 256), 256, 0, st>>>(p); ++g_tem_launches;
  return cudaGetLastError();
}

// [voxel][32 channels] halo tiles (64 B rows) in the 64B-swizzle layout
static bool make_map_sw64(CUtensorMap* m, const void* base, int B, int Z, int Y, int X, int C) {
  EncodeTiledFn enc = tem_get_encode();
  if (!enc) return false;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)Z, (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)X * C * 2, (cuuint64_t)Y * X * C * 2, (cuuint64_t)Z * Y * X * C * 2};
  cuuint32_t box[5] = {(cuuint32_t)WCH, (cuuint32_t)HX, (cuuint32_t)HY, 1, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

cudaError_t launch_conv_tcw(const ConvArgs& a, const bf16* wpacked, cudaStream_t st) {
  TcwArgs t; memset(&t, 0, sizeof(t));
  t.B = a.B; for (int i = 0; i < 3; ++i) t.L[i] = a.L[i];
  t.nchunks = (a.C0 + a.C1) / WCH; t.chunks0 = a.C0 / WCH;
  const int pad = (a.form == 1) ? 2 : 0;
  for (int i = 0; i < 3; ++i) { t.shift0[i] = a.s0.shift[i] - pad; t.shift1[i] = a.s1.shift[i] - pad; t.out_off[i] = a.out_off[i]; t.ref_off[i] = a.ref_off[i]; }
  t.wpacked = wpacked;
  t.ntx = (a.L[2] + TX - 1) / TX; t.nty = (a.L[1] + TY - 1) / TY; t.nzc = (a.L[0] + ZCAP - 1) / ZCAP;
  t.out = (bf16*)a.out; t.OZ = a.OZ; t.OY = a.OY; t.OX = a.OX; t.out_C = a.out_C; t.out_coff = a.out_coff;
  t.Cout = a.Cout; t.slope = a.slope;
  t.ref = a.ref; t.RZ = a.RZ; t.RY = a.RY; t.RX = a.RX; t.ref_C = a.ref_C; t.ref_coff = a.ref_coff; t.ref_slope = a.ref_slope;
  t.drop_key = a.drop_key; t.accumulate = a.accumulate;
  t.dbg = tem_ablation_bits();
  static const int swz = getenv("TEM_TCW_NO_SWIZZLE") ? 0 : 1;             // debug knob: 8-channel plane tiles
  t.swz = swz;
  CUtensorMap m0, m1;
  if (swz) {
    if (!make_map_sw64(&m0, a.s0.p, a.B, a.s0.Z, a.s0.Y, a.s0.X, a.s0.C)) return cudaErrorInvalidValue;
    if (a.C1) { if (!make_map_sw64(&m1, a.s1.p, a.B, a.s1.Z, a.s1.Y, a.s1.X, a.s1.C)) return cudaErrorInvalidValue; }
    else m1 = m0;
  } else {
    if (!tem_make_map_5d(&m0, a.s0.p, a.B, a.s0.Z, a.s0.Y, a.s0.X, a.s0.C, HX, HY)) return cudaErrorInvalidValue;
    if (a.C1) { if (!tem_make_map_5d(&m1, a.s1.p, a.B, a.s1.Z, a.s1.Y, a.s1.X, a.s1.C, HX, HY)) return cudaErrorInvalidValue; }
    else m1 = m0;
  }
  const size_t smem = (size_t)2 * WBYTES + (size_t)RINGW * SLOT_BYTES + 1024;
  static bool attr = false;
  if (!attr) { cudaError_t e = cudaFuncSetAttribute(conv3_tcw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e) return e; attr = true; }
  const dim3 grid((unsigned)((long long)a.B * t.ntx * t.nty * t.nzc), (unsigned)((a.Cout + CPW - 1) / CPW));
  conv3_tcw_kernel<<<grid, kThreads, smem, st>>>(m0, m1, t); ++g_tem_launches;
  return cudaGetLastError();
}
