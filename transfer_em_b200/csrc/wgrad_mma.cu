// Tensor-core weight gradient for the channel-narrow layers of the reference graph (Ca, Cb in {8,16,32,...}):
//   dw[tap][ca][cb] += sum_{b,p} S[b, s*p + tap - pad][ca] * P[b,p][cb]          
// (conv: S = layer input, P = dy; transposed conv: S = dy, P = layer input - models/utils.py:73,80,129-130).
//
// Used for the stride-2 / transposed / 1x1 layers and as the fall-back of wgrad_tc.cu (the tcgen05 kernel of the 3x3x3
// stride-1 layers).  Why warp-level mma.sync (m16n8k16, bf16 -> fp32) here: the GEMM is M = (tap, ca) x N = cb x
// K = voxels with ca, cb as small as 8.  tcgen05.mma needs M >= 64 with ONE uniform stride between its 8-row groups,
// but the 8-channel groups of this M dimension are the taps, whose shared-memory offsets (dz*HY*HX + dy*HX + dx) are
// not uniformly strided (wgrad_tc.cu side-steps that by letting tile ROWS carry the dy taps).  m16n8k16 fits
// (2 taps x 8 ca) x 8 cb exactly, and both operands come straight out of the natural [voxel][8 channels] layout with
// ldmatrix.trans, for any stride.
//
// One CTA stages a (TZ x TY x 16) tile of output positions plus its input halo with cp.async (double buffered,
// 8-channel planes so every ldmatrix row is a 16 B contiguous chunk), each warp owns up to 16 (M-tile, N-block)
// accumulator tiles in registers, and partial sums leave through fp32 atomics once per CTA.
#include <string.h>
#include "tem_kernels.cuh"

extern unsigned long long g_tem_launches;

namespace {

constexpr int kWarps = 8;
constexpr int kTPW = 16;          // accumulator tiles per warp
constexpr int TXW = 16;           // x positions per K-step

struct WgArgs {
  const bf16* S; int SZ, SY, SX; int shift[3]; long long s_bstride;
  const bf16* P; int PZ, PY, PX; int p_off[3]; long long p_bstride;
  int Ca, Cb, B, L[3], k[3], stride[3], pad[3];
  float* dw; long long ws_tap, ws_a, ws_b;
  int TZ, TY, ty_sh;                // tile (TZ x TY x 16 positions), TY a power of two
  int HZ, HY, HX;                   // halo extents
  int ntz, nty, ntx; long long ntiles; long long tiles_per_cta;
  int Mtiles, NB, ntap, ntiles_out; // output tiling
  int G, tpg, rsplit;               // tile groups, tiles per group, row splitters per group
  int s_bytes, p_bytes;             // per-buffer shared memory sizes
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Row-wise staging: a halo row (fixed z,y) is HX voxels x Ca channels of contiguous global memory, copied as 16 B
// chunks into the 8-channel planes; one warp per row, so the only per-chunk index math is j / planes, j % planes.
__device__ __forceinline__ void stage_tile(const WgArgs& a, long long tile, uint32_t sbuf, uint32_t pbuf, int warp, int lane) {
  long long t = tile;
  const int tx = (int)(t % a.ntx); t /= a.ntx;
  const int ty = (int)(t % a.nty); t /= a.nty;
  const int tz = (int)(t % a.ntz); t /= a.ntz;
  const int b = (int)t;
  const int px0 = tx * TXW, py0 = ty * a.TY, pz0 = tz * a.TZ;
  const int hvox = a.HZ * a.HY * a.HX;
  const int pa = a.Ca >> 3, pa_sh = 31 - __clz(pa);          // planes are powers of two: no integer division per chunk
  const int sz0 = pz0 * a.stride[0] - a.pad[0] + a.shift[0], sy0 = py0 * a.stride[1] - a.pad[1] + a.shift[1], sx0 = px0 * a.stride[2] - a.pad[2] + a.shift[2];
  const bf16* Sb = a.S + (long long)b * a.s_bstride;
  const int srow_chunks = a.HX * pa;
  for (int row = warp; row < a.HZ * a.HY; row += kWarps) {
    const int hz = row / a.HY, hy = row % a.HY;
    const int z = sz0 + hz, y = sy0 + hy;
    const bool rowok = z >= 0 && z < a.SZ && y >= 0 && y < a.SY;
    const bf16* rp = Sb + (((long long)z * a.SY + y) * a.SX + sx0) * a.Ca;
    for (int j = lane; j < srow_chunks; j += 32) {
      const int hx = j >> pa_sh, plane = j & (pa - 1);
      const int x = sx0 + hx;
      const bool ok = rowok && x >= 0 && x < a.SX;
      cp_async16(sbuf + (uint32_t)(plane * hvox + row * a.HX + hx) * 16u, ok ? (const void*)(rp + (long long)j * 8) : (const void*)a.S, ok);
    }
  }
  const int tvox = a.TZ * a.TY * TXW;
  const int pb = a.Cb >> 3, pb_sh = 31 - __clz(pb);
  const bf16* Pb = a.P + (long long)b * a.p_bstride;
  const int prow_chunks = TXW * pb;
  for (int row = warp; row < a.TZ * a.TY; row += kWarps) {
    const int pz = row >> a.ty_sh, py = row & (a.TY - 1);
    const int z = pz0 + pz, y = py0 + py;
    const bool rowok = z < a.L[0] && y < a.L[1];
    const bf16* rp = Pb + ((((long long)z + a.p_off[0]) * a.PY + y + a.p_off[1]) * a.PX + px0 + a.p_off[2]) * a.Cb;
    for (int j = lane; j < prow_chunks; j += 32) {
      const int px = j >> pb_sh, plane = j & (pb - 1);
      const bool ok = rowok && (px0 + px) < a.L[2];
      cp_async16(pbuf + (uint32_t)(plane * tvox + row * TXW + px) * 16u, ok ? (const void*)(rp + (long long)j * 8) : (const void*)a.P, ok);
    }
  }
}

__global__ void __launch_bounds__(kWarps * 32, 2) wgrad_mma_kernel(const WgArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t buf_bytes = (uint32_t)(a.s_bytes + a.p_bytes);
  const int hvox = a.HZ * a.HY * a.HX, tvox = a.TZ * a.TY * TXW;

  // this warp's output tiles: tile id = nblk * Mtiles + mtile (consecutive ids share the P fragment)
  // warp -> (tile group, row part): with few output tiles the 8 warps split the K rows of a tile instead
  int group, rpart;
  if (a.G >= kWarps) { group = blockIdx.y * kWarps + warp; rpart = 0; }
  else { group = warp % a.G; rpart = warp / a.G; if (rpart >= a.rsplit) group = a.G; }
  const int tile0 = group * a.tpg;
  const int tile_end = (group < a.G) ? min(tile0 + a.tpg, a.ntiles_out) : 0;
  const int mat = lane >> 3, r8 = lane & 7;
  const int a_half = mat & 1, a_vg = mat >> 1;        // A: matrices (v0-7,h0) (v0-7,h1) (v8-15,h0) (v8-15,h1)
  const int b_v = ((lane >> 3) & 1) * 8 + r8;         // B: matrices (v0-7) (v8-15)
  int aoff[kTPW], nblk[kTPW];                         // this thread's ldmatrix row offset (in 16 B voxels) per tile
#pragma unroll
  for (int i = 0; i < kTPW; ++i) {
    const int id = tile0 + i;
    const int idc = (id < tile_end) ? id : 0;
    nblk[i] = idc / a.Mtiles;
    const int m = idc % a.Mtiles;
    int tap, plane;
    if (a.Ca == 8) { tap = 2 * m + a_half; if (tap >= a.ntap) tap = 2 * m; plane = 0; }
    else { const int cb16 = a.Ca >> 4; tap = m / cb16; plane = 2 * (m % cb16) + a_half; }
    const int dx = tap % a.k[2], dy = (tap / a.k[2]) % a.k[1], dz = tap / (a.k[2] * a.k[1]);
    aoff[i] = plane * hvox + (dz * a.HY + dy) * a.HX + dx + a.stride[2] * (a_vg * 8 + r8);
  }
  float acc[kTPW][4];
#pragma unroll
  for (int i = 0; i < kTPW; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }

  const long long t_begin = (long long)blockIdx.x * a.tiles_per_cta;
  const long long t_end = min(t_begin + a.tiles_per_cta, a.ntiles);
  if (t_begin < t_end) stage_tile(a, t_begin, base, base + a.s_bytes, warp, lane);
  asm volatile("cp.async.commit_group;" ::: "memory");

  for (long long t = t_begin; t < t_end; ++t) {
    const int cur = (int)((t - t_begin) & 1);
    if (t + 1 < t_end) stage_tile(a, t + 1, base + (cur ^ 1) * buf_bytes, base + (cur ^ 1) * buf_bytes + a.s_bytes, warp, lane);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    const uint32_t sbuf = base + cur * buf_bytes, pbuf = sbuf + a.s_bytes;
    for (int row = rpart; row < a.TZ * a.TY; row += a.rsplit) {
      const int pz = row >> a.ty_sh, py = row & (a.TY - 1);
      const int rowbase = (pz * a.stride[0] * a.HY + py * a.stride[1]) * a.HX;
      // branch-free over all kTPW slots (dead slots re-use tile 0's addresses and are discarded at the end): the
      // compiler can issue the 16 ldmatrix pairs back to back instead of serialising ldmatrix -> mma latencies
      uint32_t af[kTPW][4], bfr[kTPW][2];
#pragma unroll
      for (int i = 0; i < kTPW; ++i) {
        ldsm_x2_t(pbuf + (uint32_t)(nblk[i] * tvox + row * TXW + b_v) * 16u, bfr[i][0], bfr[i][1]);
        ldsm_x4_t(sbuf + (uint32_t)(rowbase + aoff[i]) * 16u, af[i][0], af[i][1], af[i][2], af[i][3]);
      }
#pragma unroll
      for (int i = 0; i < kTPW; ++i) mma_bf16(acc[i], af[i][0], af[i][1], af[i][2], af[i][3], bfr[i][0], bfr[i][1]);
    }
    __syncthreads();
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");

  // D fragment: c0,c1 = (row g, cols 2t,2t+1); c2,c3 = (row g+8, cols 2t,2t+1); row = (half, ca), col = cb.
  // When several warps hold partial sums of the same tiles (row split), they are combined in shared memory first:
  // same-address global atomics serialise in L2 and were the tail of this kernel.
  __syncthreads();
  float* red = reinterpret_cast<float*>(smem);
  const bool use_red = a.rsplit > 1 && (size_t)a.ntiles_out * 128 * sizeof(float) <= (size_t)2 * buf_bytes;
  if (use_red) {
    for (int i = tid; i < a.ntiles_out * 128; i += kWarps * 32) red[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kTPW; ++i) {
      const int id = tile0 + i;
      if (id >= tile_end) continue;
#pragma unroll
      for (int r = 0; r < 4; ++r) if (acc[i][r] != 0.f) atomicAdd(&red[id * 128 + lane * 4 + r], acc[i][r]);
    }
    __syncthreads();
  }
  auto emit = [&](int id, int ln, int r, float v) {
    if (v == 0.f) return;
    const int g = ln >> 2, tq = ln & 3, hh = r >> 1, col = r & 1;
    const int m = id % a.Mtiles, nb = id / a.Mtiles;
    int tap, ca;
    if (a.Ca == 8) { tap = 2 * m + hh; ca = g; if (tap >= a.ntap) return; }
    else { const int cb16 = a.Ca >> 4; tap = m / cb16; ca = (2 * (m % cb16) + hh) * 8 + g; }
    atomicAdd(a.dw + tap * a.ws_tap + (long long)ca * a.ws_a + (long long)(nb * 8 + 2 * tq + col) * a.ws_b, v);
  };
  if (use_red) {
    for (int i = tid; i < a.ntiles_out * 128; i += kWarps * 32) emit(i >> 7, (i & 127) >> 2, i & 3, red[i]);
  } else {
#pragma unroll
    for (int i = 0; i < kTPW; ++i) {
      const int id = tile0 + i;
      if (id >= tile_end) continue;
#pragma unroll
      for (int r = 0; r < 4; ++r) emit(id, lane, r, acc[i][r]);
    }
  }
}

}  // namespace

bool wgrad_mma_supported(const WgradArgs& w) {
  if (w.S.dtype != DT_BF16 || w.p_dtype != DT_BF16) return false;
  if (w.S.origins || w.use_lut) return false;
  if (w.Ca % 8 || w.Cb % 8 || w.Ca < 8 || w.Cb < 8) return false;
  if (!(w.Ca == 8 || w.Ca % 16 == 0)) return false;
  if (((w.Ca >> 3) & ((w.Ca >> 3) - 1)) || ((w.Cb >> 3) & ((w.Cb >> 3) - 1))) return false;   // 8-channel plane counts are powers of two
  if (w.S.C != w.Ca || w.S.coff != 0 || w.p_C != w.Cb || w.p_coff != 0) return false;
  for (int i = 0; i < 3; ++i) if (w.stride[i] > 2) return false;
  return true;
}

cudaError_t launch_wgrad_mma(const WgradArgs& w, cudaStream_t st) {
  WgArgs a; memset(&a, 0, sizeof(a));
  a.S = (const bf16*)w.S.p; a.SZ = w.S.Z; a.SY = w.S.Y; a.SX = w.S.X; a.s_bstride = w.S.bstride;
  a.P = (const bf16*)w.P; a.PZ = w.PZ; a.PY = w.PY; a.PX = w.PX; a.p_bstride = w.p_bstride;
  for (int i = 0; i < 3; ++i) { a.shift[i] = w.S.shift[i]; a.p_off[i] = w.p_off[i]; a.L[i] = w.L[i]; a.k[i] = w.k[i]; a.stride[i] = w.stride[i]; a.pad[i] = w.pad[i]; }
  a.Ca = w.Ca; a.Cb = w.Cb; a.B = w.B; a.dw = w.dw; a.ws_tap = w.ws_tap; a.ws_a = w.ws_a; a.ws_b = w.ws_b;
  a.ntap = w.k[0] * w.k[1] * w.k[2];
  a.Mtiles = (w.Ca == 8) ? (a.ntap + 1) / 2 : a.ntap * (w.Ca / 16);
  a.NB = w.Cb / 8;
  a.ntiles_out = a.Mtiles * a.NB;
  if ((long long)w.B * w.L[0] * w.L[1] * w.L[2] == 0) return cudaSuccess;
  // tile: shrink until two buffers fit in ~96 KB
  int TZ = (w.L[0] >= 4) ? 4 : (w.L[0] >= 2 ? 2 : 1), TY = (w.L[1] >= 8) ? 8 : (w.L[1] >= 4 ? 4 : (w.L[1] >= 2 ? 2 : 1));
  auto bytes = [&](int tz, int ty, int& hz, int& hy, int& hx, int& sb, int& pb) {
    hz = (tz - 1) * w.stride[0] + w.k[0]; hy = (ty - 1) * w.stride[1] + w.k[1]; hx = (TXW - 1) * w.stride[2] + w.k[2];
    sb = ((hz * hy * hx * w.Ca * 2) + 127) & ~127; pb = ((tz * ty * TXW * w.Cb * 2) + 127) & ~127;
    return 2 * (sb + pb);
  };
  int hz, hy, hx, sb, pb;
  while (bytes(TZ, TY, hz, hy, hx, sb, pb) > 100 * 1024 && (TZ > 1 || TY > 1)) { if (TZ > 1) TZ >>= 1; else TY >>= 1; }
  // keep enough tiles to occupy the machine
  auto ntl = [&](int tz, int ty) { return (long long)w.B * ((w.L[0] + tz - 1) / tz) * ((w.L[1] + ty - 1) / ty) * ((w.L[2] + TXW - 1) / TXW); };
  while (ntl(TZ, TY) < 2 * 148 && (TZ > 1 || TY > 2)) { if (TZ > 1) TZ >>= 1; else TY >>= 1; }
  const int smem = bytes(TZ, TY, hz, hy, hx, sb, pb);
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  a.TZ = TZ; a.TY = TY; a.ty_sh = (TY == 8) ? 3 : (TY == 4 ? 2 : (TY == 2 ? 1 : 0)); a.HZ = hz; a.HY = hy; a.HX = hx; a.s_bytes = sb; a.p_bytes = pb;
  a.ntz = (w.L[0] + TZ - 1) / TZ; a.nty = (w.L[1] + TY - 1) / TY; a.ntx = (w.L[2] + TXW - 1) / TXW;
  a.ntiles = (long long)w.B * a.ntz * a.nty * a.ntx;
  a.G = (a.ntiles_out + kTPW - 1) / kTPW;
  a.tpg = (a.ntiles_out + a.G - 1) / a.G;
  a.rsplit = (a.G >= kWarps) ? 1 : kWarps / a.G;
  const int gy = (a.G + kWarps - 1) / kWarps;
  long long gx = (2 * 148 + gy - 1) / gy;
  if (gx > a.ntiles) gx = a.ntiles;
  a.tiles_per_cta = (a.ntiles + gx - 1) / gx;
  gx = (a.ntiles + a.tiles_per_cta - 1) / a.tiles_per_cta;
  static bool attr = false;
  if (!attr) { cudaError_t e = cudaFuncSetAttribute(wgrad_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); if (e) return e; attr = true; }
  wgrad_mma_kernel<<<dim3((unsigned)gx, gy), kWarps * 32, smem, st>>>(a); ++g_tem_launches;
  return cudaGetLastError();
}
