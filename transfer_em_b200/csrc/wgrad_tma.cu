// TMA-staged tensor-core weight gradient (second generation of wgrad_mma.cu, same math and tile algebra):
//   dw[tap][ca][cb] += sum_{b,p} S[b, s*p + tap - pad][ca] * P[b,p][cb]
// One CTA owns a (TY x 16)-position column of one sample and marches along z.  A dedicated producer warp streams the
// input z-slices of S (halo box (8ch, HX, HY) per 8-channel plane, OOB zero-fill = padding / ragged edges) and the
// position slices of P through two mbarrier rings with TMA, so staging costs no issue slots in the eight consumer
// warps, each input slice is fetched once per column (z reuse through the ring), and partial sums stay in registers
// for the whole march.  Consumers run ldmatrix.trans + mma.sync.m16n8k16 exactly as in wgrad_mma.cu.
#include <string.h>
#include "tem_kernels.cuh"
#include "ptx_sm100.cuh"

extern unsigned long long g_tem_launches;

namespace {

constexpr int kCons = 8;            // consumer warps
constexpr int kTPWt = 16;           // accumulator tiles per warp
constexpr int RS_MAX = 12, RP_MAX = 4;

struct WtArgs {
  int B, L[3], k[3], stride[3], s_org[3];
  int Ca, Cb, pa, pb;
  int TY, HY, HX;
  int s_plane, s_slot, p_plane, p_slot;   // bytes
  int RS, RP, ZG;                         // ring depths; output z-slices per barrier round
  int nty, ntx, nzc, zc;
  float* dw; long long ws_tap, ws_a, ws_b;
  int Mtiles, NB, ntap, ntiles_out, G, tpg, rsplit;
};

__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm2t(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma_t(float* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__((kCons + 1) * 32, 2)
wgrad_tma_kernel(const __grid_constant__ CUtensorMap mapS, const __grid_constant__ CUtensorMap mapP, const WtArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t sfull[RS_MAX], sempty[RS_MAX], pfull[RP_MAX], pempty[RP_MAX];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* sring = smem;
  uint8_t* pring = smem + (size_t)a.RS * a.s_slot;
  if (tid == 0) {
    for (int i = 0; i < a.RS; ++i) { mbar_init(&sfull[i], 1); mbar_init(&sempty[i], kCons); }
    for (int i = 0; i < a.RP; ++i) { mbar_init(&pfull[i], 1); mbar_init(&pempty[i], kCons); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  int w = blockIdx.x;
  const int zc_i = w % a.nzc; w /= a.nzc;
  const int tx_i = w % a.ntx; w /= a.ntx;
  const int ty_i = w % a.nty; w /= a.nty;
  const int b = w;
  const int px0 = tx_i * 16, py0 = ty_i * a.TY, pz0 = zc_i * a.zc;
  const int nz = min(a.zc, a.L[0] - pz0);
  const int sz = a.stride[0];
  const int nS = sz * (nz - 1) + a.k[0];             // input z-slices this CTA consumes

  if (warp == kCons) {
    // ================= producer =================
    if (lane == 0) {
      const int sx0 = a.stride[2] * px0 + a.s_org[2], sy0 = a.stride[1] * py0 + a.s_org[1], sz0 = sz * pz0 + a.s_org[0];
      int s_issued = 0, sslot = 0; uint32_t sph = 0;
      int pslot = 0; uint32_t pph = 0;
      for (int j0 = 0; j0 < nz; j0 += a.ZG) {
        const int zg = min(a.ZG, nz - j0);
        const int need = sz * (j0 + zg - 1) + a.k[0];  // slices 0 .. need-1 must be in flight for this group of output slices
        while (s_issued < need) {
          mbar_wait(&sempty[sslot], sph ^ 1u);
          mbar_arrive_expect_tx(&sfull[sslot], (uint32_t)(a.pa * a.HY * a.HX * 16));
          for (int p = 0; p < a.pa; ++p)
            tma_load_5d(sring + (size_t)sslot * a.s_slot + p * a.s_plane, &mapS, &sfull[sslot], p * 8, sx0, sy0, sz0 + s_issued, b);
          ++s_issued; if (++sslot == a.RS) { sslot = 0; sph ^= 1u; }
        }
        mbar_wait(&pempty[pslot], pph ^ 1u);
        mbar_arrive_expect_tx(&pfull[pslot], (uint32_t)(a.pb * a.ZG * a.TY * 16 * 16));
        for (int p = 0; p < a.pb; ++p)
          tma_load_5d(pring + (size_t)pslot * a.p_slot + p * a.p_plane, &mapP, &pfull[pslot], p * 8, px0, py0, pz0 + j0, b);
        if (++pslot == a.RP) { pslot = 0; pph ^= 1u; }
      }
      (void)nS;
    }
    return;
  }

  // ================= consumers =================
  int group, rpart;
  if (a.G >= kCons) { group = blockIdx.y * kCons + warp; rpart = 0; }
  else { group = warp % a.G; rpart = warp / a.G; if (rpart >= a.rsplit) group = a.G; }
  const int tile0 = group * a.tpg;
  const int tile_end = (group < a.G) ? min(tile0 + a.tpg, a.ntiles_out) : 0;
  const int mat = lane >> 3, r8 = lane & 7;
  const int a_half = mat & 1, a_vg = mat >> 1;
  const int b_v = ((lane >> 3) & 1) * 8 + r8;
  int aoff[kTPWt], meta[kTPWt];                       // meta = kz tap + 4 * n-block
#pragma unroll
  for (int i = 0; i < kTPWt; ++i) {
    const int id = tile0 + i;
    const int idc = (id < tile_end) ? id : 0;
    const int nb_i = idc / a.Mtiles;
    const int m = idc % a.Mtiles;
    int tap, plane;
    if (a.Ca == 8) { tap = 2 * m + a_half; if (tap >= a.ntap) tap = 2 * m; plane = 0; }
    else { const int cb16 = a.Ca >> 4; tap = m / cb16; plane = 2 * (m % cb16) + a_half; }
    const int dx = tap % a.k[2], dy = (tap / a.k[2]) % a.k[1];
    meta[i] = tap / (a.k[2] * a.k[1]) + 4 * nb_i;
    aoff[i] = plane * a.s_plane + (dy * a.HX + dx + a.stride[2] * (a_vg * 8 + r8)) * 16;
  }
  float acc[kTPWt][4];
#pragma unroll
  for (int i = 0; i < kTPWt; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }

  const uint32_t sbase = smem_u32(sring), pbase = smem_u32(pring);
  int s_waited = 0, wslot = 0; uint32_t wph = 0;      // S slices known to have landed
  int zslot = 0;                                      // ring slot of input slice sz*j
  int rslot = 0;                                      // next S slot to release
  int pslot = 0; uint32_t pph = 0;
  const int row_bytes = a.stride[1] * a.HX * 16;
  const int slice_rows = a.TY * 16 * 16;              // bytes of one z-slice inside a P plane
  for (int j0 = 0; j0 < nz; j0 += a.ZG) {
    const int zg = min(a.ZG, nz - j0);
    const int need = sz * (j0 + zg - 1) + a.k[0];
    while (s_waited < need) { mbar_wait(&sfull[wslot], wph); ++s_waited; if (++wslot == a.RS) { wslot = 0; wph ^= 1u; } }
    mbar_wait(&pfull[pslot], pph);
    const uint32_t pb_addr = pbase + (uint32_t)pslot * a.p_slot;
    if (tile_end > tile0) {
      for (int rr = rpart; rr < zg * a.TY; rr += a.rsplit) {
        const int jj = rr / a.TY, row = rr - jj * a.TY;
        int zs = zslot + sz * jj; if (zs >= a.RS) zs -= a.RS;
        uint32_t b0 = 0, b1 = 0; int bcur = -1;
        const uint32_t rterm = (uint32_t)(row * row_bytes);
#pragma unroll
        for (int i = 0; i < kTPWt; ++i) {
          if (tile0 + i >= tile_end) continue;          // warp-uniform
          const int nb_i = meta[i] >> 2, dz_i = meta[i] & 3;
          if (nb_i != bcur) {
            bcur = nb_i;
            ldsm2t(pb_addr + (uint32_t)(bcur * a.p_plane + jj * slice_rows + (row * 16 + b_v) * 16), b0, b1);
          }
          int sl = zs + dz_i; if (sl >= a.RS) sl -= a.RS;
          uint32_t a0, a1, a2, a3;
          ldsm4t(sbase + (uint32_t)sl * a.s_slot + rterm + (uint32_t)aoff[i], a0, a1, a2, a3);
          mma_t(acc[i], a0, a1, a2, a3, b0, b1);
        }
      }
    }
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(&pempty[pslot]);
      for (int r = 0; r < sz * zg; ++r) { mbar_arrive(&sempty[rslot]); if (++rslot == a.RS) rslot = 0; }
    }
    if (++pslot == a.RP) { pslot = 0; pph ^= 1u; }
    zslot += sz * zg; while (zslot >= a.RS) zslot -= a.RS;
  }

  // ---- combine the row-split partial sums in shared memory, then one atomic per output per CTA
  asm volatile("bar.sync 1, %0;" ::"r"(kCons * 32) : "memory");
  float* red = reinterpret_cast<float*>(smem);
  const size_t ring_bytes = (size_t)a.RS * a.s_slot + (size_t)a.RP * a.p_slot;
  const bool use_red = a.rsplit > 1 && (size_t)a.ntiles_out * 128 * sizeof(float) <= ring_bytes;
  if (use_red) {
    for (int i = tid; i < a.ntiles_out * 128; i += kCons * 32) red[i] = 0.f;
    asm volatile("bar.sync 1, %0;" ::"r"(kCons * 32) : "memory");
#pragma unroll
    for (int i = 0; i < kTPWt; ++i) {
      const int id = tile0 + i;
      if (id >= tile_end) continue;
#pragma unroll
      for (int r = 0; r < 4; ++r) if (acc[i][r] != 0.f) atomicAdd(&red[id * 128 + lane * 4 + r], acc[i][r]);
    }
    asm volatile("bar.sync 1, %0;" ::"r"(kCons * 32) : "memory");
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
    for (int i = tid; i < a.ntiles_out * 128; i += kCons * 32) emit(i >> 7, (i & 127) >> 2, i & 3, red[i]);
  } else {
#pragma unroll
    for (int i = 0; i < kTPWt; ++i) {
      const int id = tile0 + i;
      if (id >= tile_end) continue;
#pragma unroll
      for (int r = 0; r < 4; ++r) emit(id, lane, r, acc[i][r]);
    }
  }
}

}  // namespace

bool wgrad_tma_supported(const WgradArgs& w) {
  if (!wgrad_mma_supported(w)) return false;
  if (tem_get_encode() == nullptr) return false;
  if (w.p_off[0] || w.p_off[1] || w.p_off[2]) return false;
  if (w.PZ != w.L[0] || w.PY != w.L[1] || w.PX != w.L[2]) return false;
  if (w.k[0] > 4) return false;
  if (w.S.bstride != (long long)w.S.Z * w.S.Y * w.S.X * w.S.C) return false;
  if (w.p_bstride != (long long)w.PZ * w.PY * w.PX * w.p_C) return false;
  return true;
}

cudaError_t launch_wgrad_tma(const WgradArgs& w, cudaStream_t st) {
  WtArgs a; memset(&a, 0, sizeof(a));
  a.B = w.B;
  for (int i = 0; i < 3; ++i) { a.L[i] = w.L[i]; a.k[i] = w.k[i]; a.stride[i] = w.stride[i]; a.s_org[i] = w.S.shift[i] - w.pad[i]; }
  if ((long long)w.B * w.L[0] * w.L[1] * w.L[2] == 0) return cudaSuccess;
  a.Ca = w.Ca; a.Cb = w.Cb; a.pa = w.Ca / 8; a.pb = w.Cb / 8;
  a.dw = w.dw; a.ws_tap = w.ws_tap; a.ws_a = w.ws_a; a.ws_b = w.ws_b;
  a.ntap = w.k[0] * w.k[1] * w.k[2];
  a.Mtiles = (w.Ca == 8) ? (a.ntap + 1) / 2 : a.ntap * (w.Ca / 16);
  a.NB = w.Cb / 8; a.ntiles_out = a.Mtiles * a.NB;
  a.G = (a.ntiles_out + kTPWt - 1) / kTPWt;
  a.tpg = (a.ntiles_out + a.G - 1) / a.G;
  a.rsplit = (a.G >= kCons) ? 1 : kCons / a.G;
  const int gy = (a.G + kCons - 1) / kCons;
  int TY = (w.L[1] >= 8) ? 8 : (w.L[1] >= 4 ? 4 : (w.L[1] >= 2 ? 2 : 1));
  a.RP = 2;
  auto layout = [&](int ty, int zg) {
    a.TY = ty; a.ZG = zg; a.HY = (ty - 1) * w.stride[1] + w.k[1]; a.HX = 15 * w.stride[2] + w.k[2];
    a.RS = w.stride[0] * (zg - 1) + w.k[0] + w.stride[0] * zg;       // one group in use + the next group in flight
    if (a.RS > RS_MAX) a.RS = RS_MAX;
    a.s_plane = (a.HY * a.HX * 16 + 127) & ~127; a.s_slot = a.pa * a.s_plane;
    a.p_plane = zg * ty * 16 * 16; a.p_slot = a.pb * a.p_plane;
    return (size_t)a.RS * a.s_slot + (size_t)a.RP * a.p_slot;
  };
  int ZG = (w.L[0] >= 4) ? 4 : (w.L[0] >= 2 ? 2 : 1);
  while (layout(TY, ZG) > 100 * 1024 && (ZG > 1 || TY > 1)) { if (ZG > 1) ZG >>= 1; else TY >>= 1; }
  if (w.stride[0] * (ZG - 1) + w.k[0] > a.RS) return cudaErrorInvalidConfiguration;
  const size_t smem = layout(TY, ZG);
  if (smem > 200 * 1024 || a.HX > 256 || a.HY > 256) return cudaErrorInvalidConfiguration;
  a.nty = (w.L[1] + TY - 1) / TY; a.ntx = (w.L[2] + 15) / 16;
  const long long cols = (long long)w.B * a.nty * a.ntx;
  int nzc = 1;
  while (cols * nzc * gy < 2 * 148 && (w.L[0] + nzc) / (nzc + 1) >= 4) ++nzc;
  a.zc = (w.L[0] + nzc - 1) / nzc; a.nzc = (w.L[0] + a.zc - 1) / a.zc;
  CUtensorMap mS, mP;
  if (!tem_make_map_5d(&mS, w.S.p, w.B, w.S.Z, w.S.Y, w.S.X, w.S.C, a.HX, a.HY)) return cudaErrorInvalidValue;
  if (!tem_make_map_5d(&mP, w.P, w.B, w.PZ, w.PY, w.PX, w.p_C, 16, TY, ZG)) return cudaErrorInvalidValue;
  static bool attr = false;
  if (!attr) { cudaError_t e = cudaFuncSetAttribute(wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); if (e) return e; attr = true; }
  dim3 grid((unsigned)(cols * a.nzc), gy);
  wgrad_tma_kernel<<<grid, (kCons + 1) * 32, smem, st>>>(mS, mP, a); ++g_tem_launches;
  return cudaGetLastError();
}
