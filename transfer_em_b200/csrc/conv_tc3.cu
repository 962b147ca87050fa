// tcgen05 implicit-GEMM 3x3x3 stride-1 convolution: the three kz taps share one MMA.
//
// A per-OUTPUT-slice formulation issues 27*Cin/16 MMAs of shape M128 x N(Cout) x K16 per slice; with Cout = 8..32 every
// one of them re-reads a 4 KB A tile from shared memory for 8-16 cycles of math and the kernel is pinned at the
// shared-memory A-read rate (profiles/README.md, round 1).  Here the MMAs are issued per INPUT z-slice instead: an input slice s
// contributes to the output slices s, s-1, s-2 through the taps kz = 0, 1, 2, so one MMA with
// N' = 3*CP columns (B = [W(kz=0) | W(kz=1) | W(kz=2)]) accumulates all three at once into three ADJACENT column
// groups of a TMEM-resident accumulator strip: output slice zo lives at column (ZCAP + 1 - zo) * CP, hence slices
// s, s-1, s-2 are consecutive.  One third of the MMAs and of the A-operand traffic; every input slice is consumed by
// exactly one MMA batch (the ring slot is released at once); accumulators of a whole z-chunk stay in TMEM
// (<= 512 columns), are zeroed once with tcgen05.st and drained slice by slice by the epilogue warps while later
// slices are still being accumulated.
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
constexpr int RING3_MAX = 16;
constexpr int kTcThreads = 320;     // TMA producer, MMA issuer, 2 x 4 epilogue warps (two per TMEM lane quadrant, alternating output slices)
constexpr int kMaxChunk = 64;

struct Tc3Args {
  int B, L[3];                 // conv output extent (z,y,x)
  int planes0, planes1;        // 8-channel planes taken from map0 / map1
  int merged0, merged1;        // map folds (channel, x): see tem_make_map_c8
  int shift0[3], shift1[3];    // tensor coordinate = conv-input coordinate + shift (z,y,x)
  int spd;                     // k-steps (K=16 MMAs) per dz
  int cin8;                    // 1 when Cin == 8 (tap-pair k-steps)
  const bf16* wpacked; int wbytes;
  int ntx, nty, nzc, zc, zcap, tmem_cols, ring;
  int items, nbmax;            // work items (persistent CTAs loop over them), pipeline steps of the longest item
  int sb;                      // input slices per pipeline step: one full / empty / tfull barrier round trip per sb slices
  bf16* out; int OZ, OY, OX, out_C, out_coff, out_off[3];
  int out_f32;                 // Cout == 1 (last generator layer): channel 0 leaves as fp32, or, with st_out, as uint8 into the stitched volume
  uint8_t* st_out; const int* st_index; int st_tpad, st_od; float st_mean, st_std; long long st_OZ, st_OY, st_OX;
  int Cout;
  float slope;
  const bf16* ref; int RZ, RY, RX, ref_C, ref_coff, ref_off[3]; float ref_slope;
  uint32_t drop_key;
  int accumulate;
  int split;                   // != 0: channels [split, Cout) go to out2 / ref2 (merged dgrad of a concat layer); dropout on [0, split)
  bf16* out2; int O2Z, O2Y, O2X, out2_C, out2_off[3];
  const bf16* ref2; int R2Z, R2Y, R2X, ref2_C, ref2_off[3]; float ref2_slope;
  int dbg;                     // stage-ablation bits (-DTEM_ABLATION builds only): 1 no epilogue work, 4 no input loads, 16 no MMAs
};


__device__ __forceinline__ void tmem_st8_zero(uint32_t taddr) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(0u) : "memory");
}

// MODE selects the epilogue at compile time (a kernel that carries all of them is > 64 KB of code, which eight epilogue warps
// per CTA fetch over and over): 0 forward (LeakyReLU), 1 data gradient (LeakyReLU' mask of `ref`, optional accumulate),
// 2 data gradient with dropout, 3 Cout = 1 forward with fp32 / fused uint8 output (CP = 8)
template <int CP, bool SPLIT, int MODE>
__global__ void __launch_bounds__(kTcThreads, 2)
conv3_tc3_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1, const Tc3Args a) {
  constexpr int NP = (CP == 8) ? 32 : 3 * CP;          // MMA N: three kz column groups (+ one zero group when CP == 8)
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[RING3_MAX], empty_bar[RING3_MAX], w_bar, strip_bar, tfull_bar[kMaxChunk];
  const int RING3 = a.ring;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int planes = a.planes0 + a.planes1;
  const uint32_t wbytes_pad = (uint32_t)((a.wbytes + 1023) & ~1023);
  uint8_t* wsm = smem;
  uint8_t* ring = smem + wbytes_pad;
  const int SB = a.sb;
  const uint32_t slice_bytes = (uint32_t)planes * PLANE_STRIDE;
  const uint32_t slot_bytes = (uint32_t)SB * slice_bytes;

  // Persistent CTAs (two per SM): TMEM allocation, barrier set-up, the weight image and the first zeroing of the
  // accumulator strip are paid once per CTA, not once per work item; the TMA producer runs ahead into the next item while
  // the epilogue still drains the current one.  Work item = (sample, y tile, x tile, z chunk), item k of this CTA is
  // blockIdx.x + k * gridDim.x.  (Stage ablation of the one-item-per-CTA version, profiles/ablation_r2.txt: 25 of 52 us
  // remained on g1 with loads, MMAs and epilogue work all switched off -- four rounds of CTA start-up and drain.)
  auto decode = [&](int it, int& b, int& x0, int& y0, int& z0, int& nz) {
    int w = it;
    const int zc_i = w % a.nzc; w /= a.nzc;
    const int tx_i = w % a.ntx; w /= a.ntx;
    const int ty_i = w % a.nty; w /= a.nty;
    b = w; x0 = tx_i * TX; y0 = ty_i * TY; z0 = zc_i * a.zc;
    nz = min(a.zc, a.L[0] - z0);
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < RING3; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&w_bar, 1); mbar_init(&strip_bar, 8);           // one arrival per epilogue warp
    for (int i = 0; i < a.nbmax; ++i) mbar_init(&tfull_bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)a.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  // Pipeline granularity.  Measured on B200 (tools/ubench/sync_latency.cu): one producer -> MMA -> epilogue hand-over
  // (mbarrier wait + tcgen05.commit x 2 + loop) costs the issuing warp 290-390 cycles however deep the ring is, while the
  // five N = 32 MMAs of a Cin = 8 slice keep the tensor pipe busy for ~180.  A pipeline step therefore covers SB input
  // slices (>= ~24 MMAs): one expect_tx / full wait / empty commit / tfull commit per step.
  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&w_bar, (uint32_t)a.wbytes);
      bulk_load(wsm, a.wpacked, (uint32_t)a.wbytes, &w_bar);
      int slot = 0; uint32_t ph = 0;
      for (int it = blockIdx.x; it < a.items; it += gridDim.x) {
        int b, x0, y0, z0, nz; decode(it, b, x0, y0, z0, nz);
        const int nslices = nz + 2, nbatch = (nslices + SB - 1) / SB;
        for (int bt = 0; bt < nbatch; ++bt) {
          const int s0 = bt * SB, n_in = min(SB, nslices - s0);
          mbar_wait(&empty_bar[slot], ph ^ 1u);
          if (a.dbg & 4) { mbar_arrive(&full_bar[slot]); if (++slot == RING3) { slot = 0; ph ^= 1u; } continue; }
          mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)(n_in * planes) * PLANE_BYTES);
          for (int i = 0; i < n_in; ++i) {
            uint8_t* dst = ring + (size_t)slot * slot_bytes + (size_t)i * slice_bytes;
            const int s = s0 + i;
            for (int p = 0; p < a.planes0; ++p)
              tma_load_plane(dst + p * PLANE_STRIDE, &map0, &full_bar[slot], a.merged0, p, x0 + a.shift0[2], y0 + a.shift0[1], z0 + s + a.shift0[0], b);
            for (int p = 0; p < a.planes1; ++p)
              tma_load_plane(dst + (a.planes0 + p) * PLANE_STRIDE, &map1, &full_bar[slot], a.merged1, p, x0 + a.shift1[2], y0 + a.shift1[1], z0 + s + a.shift1[0], b);
          }
          if (++slot == RING3) { slot = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // The whole warp runs the issue loop convergently, so descriptors and barrier addresses stay in uniform registers;
    // one elected lane issues the MMAs / commits of a step (no per-instruction election loops in the SASS).
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    mbar_wait(&w_bar, 0);
    const uint32_t wbase = smem_u32(wsm);
    const uint32_t rbase = smem_u32(ring);
    // descriptor words: only the start-address field (low 14 bits of the low word, 16 B units) changes per MMA
    const uint32_t a_hi = ((uint32_t)(HX * 16) >> 4) | (1u << 14);          // SBO = one halo row, version 1
    const uint32_t b_hi = (128u >> 4) | (1u << 14);                         // SBO = 128 B between n-groups
    const uint32_t b_lbo = ((uint32_t)(NP * 16) >> 4) << 16;                // LBO = NP*16 B between the K halves
    const uint32_t a_lbo_planes = ((uint32_t)PLANE_STRIDE >> 4) << 16;      // Cin >= 16: K halves are two planes
    const uint32_t wb16 = wbase >> 4;
    const int kcs = planes >> 1;
    int slot = 0; uint32_t ph = 0, item_ph = 0;
    for (int it = blockIdx.x; it < a.items; it += gridDim.x) {
      int b, x0, y0, z0, nz; decode(it, b, x0, y0, z0, nz);
      const int nslices = nz + 2, nbatch = (nslices + SB - 1) / SB;
      mbar_wait(&strip_bar, item_ph); item_ph ^= 1u;        // the strip is clean: zeroed at start / drained + re-zeroed by the epilogue
      for (int bt = 0; bt < nbatch; ++bt) {
        const int s0 = bt * SB, n_in = min(SB, nslices - s0);
        mbar_wait(&full_bar[slot], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t slot16 = (rbase + (uint32_t)slot * slot_bytes) >> 4;
        if (elect_one()) {
          for (int i = 0; i < n_in; ++i) {
            const int s = s0 + i;
            // columns of output slices s (kz=0), s-1 (kz=1), s-2 (kz=2) are adjacent
            const uint32_t d_tmem = tmem_base + (uint32_t)((a.zcap + 1 - s) * CP);
            const uint32_t sb16 = slot16 + (uint32_t)i * (slice_bytes >> 4);
            uint32_t blo = wb16 | b_lbo;
            if (a.dbg & 16) {
            } else if (a.cin8) {
#pragma unroll
              for (int p = 0; p < 5; ++p) {
                const int t0 = 2 * p, t1 = (2 * p + 1 < 9) ? 2 * p + 1 : 2 * p;
                const uint32_t o0 = (uint32_t)((t0 / 3) * HX + (t0 % 3));
                const uint32_t o1 = (uint32_t)((t1 / 3) * HX + (t1 % 3));
                const uint32_t alo = (sb16 + o0) | ((o1 - o0) << 16);       // LBO = distance between the two taps (0: dummy half)
                umma_bf16(d_tmem, ((uint64_t)a_hi << 32) | alo, ((uint64_t)b_hi << 32) | blo, idesc, 1u);
                blo += (uint32_t)(NP * 32) >> 4;
              }
            } else {
#pragma unroll
              for (int t = 0; t < 9; ++t) {
                uint32_t alo = (sb16 + (uint32_t)((t / 3) * HX + (t % 3))) | a_lbo_planes;
                for (int kc = 0; kc < kcs; ++kc) {
                  umma_bf16(d_tmem, ((uint64_t)a_hi << 32) | alo, ((uint64_t)b_hi << 32) | blo, idesc, 1u);
                  alo += (uint32_t)(2 * PLANE_STRIDE) >> 4;
                  blo += (uint32_t)(NP * 32) >> 4;
                }
              }
            }
          }
          umma_commit(&empty_bar[slot]);                       // the input slices of this step are consumed by it only
          umma_commit(&tfull_bar[bt]);                         // output slices <= s0 + n_in - 3 have received their three kz parts
        }
        __syncwarp();
        if (++slot == RING3) { slot = 0; ph ^= 1u; }
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int yl = row >> 3, xl = row & 7;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const int ew = (warp - 2) >> 2;                 // 0 / 1: this warp drains the even / odd output slices of its quadrant
    if (ew == 0) {
      // zero the accumulator strip once (this quadrant's 32 lanes, all allocated columns); afterwards every drained column
      // group is re-zeroed right after it has been read
      for (int c = 0; c < a.tmem_cols; c += 8) tmem_st8_zero(lane_base + (uint32_t)c);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (lane == 0) mbar_arrive(&strip_bar);
    const bool has_c8[4] = {0 < a.Cout, 8 < a.Cout, 16 < a.Cout, 24 < a.Cout};
    // Operands of the fused epilogue (stored activation of the LeakyReLU' factor) are fetched PF output slices ahead of
    // their use: with one 16 B load per thread and slice in flight an SM keeps only ~4 KB of this stream outstanding,
    // far below what the ~2 us round trip needs (Little's law); PF slices ahead restore the bandwidth.
    constexpr int PF = (MODE == 1 || MODE == 2) ? (32 / CP >= 2 ? 32 / CP * 2 : 2) : 1;      // CP = 8: 8 slices, 16: 4, 32: 2  (32 registers); forward: no operand to prefetch, one copy of the loop body
    const long long ref_zstride = (long long)a.RY * a.RX * a.ref_C;
    const long long ref2_zstride = (long long)a.R2Y * a.R2X * a.ref2_C;
    const int split = SPLIT ? a.split : 64;                // 8-channel chunks at or above it belong to the second destination
    const int drop_C = SPLIT ? a.split : a.Cout;
    uint64_t tf_ph = 0;                                      // phase bit of every tfull barrier (an item may use fewer steps)
    for (int it = blockIdx.x; it < a.items; it += gridDim.x) {
      int b, x0, y0, z0, nz; decode(it, b, x0, y0, z0, nz);
      const int nbatch = (nz + 2 + SB - 1) / SB;
      const int oy = y0 + yl, ox = x0 + xl;
      const bool inside = oy < a.L[1] && ox < a.L[2] && !(a.dbg & 1);
      uint4 refq[PF][CP / 8];
      // per-item bases: the per-slice part of every address is one multiply-add (index arithmetic hoisted out of the slice loop)
      const long long out_zstride = (long long)a.OY * a.OX * a.out_C;
      bf16* const out_base = a.out + ((((long long)b * a.OZ + z0 + a.out_off[0]) * a.OY + oy + a.out_off[1]) * a.OX + ox + a.out_off[2]) * a.out_C + a.out_coff;
      const uint32_t di_base = (uint32_t)(((((long long)b * a.L[0] + z0) * a.L[1] + oy) * a.L[2] + ox) * drop_C);
      const uint32_t di_zstride = (uint32_t)(a.L[1] * a.L[2] * drop_C);
      const long long ref_base = ((((long long)b * a.RZ + z0 + a.ref_off[0]) * a.RY + oy + a.ref_off[1]) * a.RX + ox + a.ref_off[2]) * a.ref_C + a.ref_coff;
      // second destination (CP >= 16 only): bases carry -split so that chunk c sits at base + c there too
      const long long out2_zstride = (long long)a.O2Y * a.O2X * a.out2_C;
      bf16* const out2_base = SPLIT ? a.out2 + ((((long long)b * a.O2Z + z0 + a.out2_off[0]) * a.O2Y + oy + a.out2_off[1]) * a.O2X + ox + a.out2_off[2]) * a.out2_C - a.split : nullptr;
      const long long ref2_base = !SPLIT ? 0 : ((((long long)b * a.R2Z + z0 + a.ref2_off[0]) * a.R2Y + oy + a.ref2_off[1]) * a.R2X + ox + a.ref2_off[2]) * a.ref2_C - a.split;
      auto fetch_ref = [&](int zo, uint4* qv) {
        if ((MODE == 1 || MODE == 2) && a.ref && inside && zo < nz) {
#pragma unroll
          for (int c = 0; c < CP / 8; ++c) {
            if (!has_c8[c]) continue;
            const bf16* rp = (SPLIT && c * 8 >= split) ? a.ref2 + ref2_base + (long long)zo * ref2_zstride : a.ref + ref_base + (long long)zo * ref_zstride;
            qv[c] = __ldg(reinterpret_cast<const uint4*>(rp + c * 8));
          }
        }
      };
#pragma unroll
      for (int u = 0; u < PF; ++u) fetch_ref(2 * u + ew, refq[u]);
      for (int zb = 0; zb < nz; zb += 2 * PF) {
#pragma unroll
        for (int pu = 0; pu < PF; ++pu) {
          const int zo = zb + 2 * pu + ew;
          if (zo >= nz) break;
          // output slice zo has its three kz parts once input slice zo + 2 has been multiplied.  CP == 8: the MMA of input
          // slice zo + 3 still touches this column group (the fourth, zero-weight group of N = 32 adds 0 to it) and would
          // write a stale value back over the re-zeroed columns, so that step must have retired as well
          const int tb = min(zo + (CP == 8 ? 3 : 2), nz + 1) / SB;
          mbar_wait(&tfull_bar[tb], (uint32_t)(tf_ph >> tb) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          uint32_t r[CP];
          const uint32_t taddr = lane_base + (uint32_t)((a.zcap + 1 - zo) * CP);
#pragma unroll
          for (int c = 0; c < CP; c += 8) tmem_ld8(taddr + c, r + c);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int c = 0; c < CP; c += 8) tmem_st8_zero(taddr + c);          // clean for the next item of this CTA
          if (!inside) continue;
          float v[CP];
#pragma unroll
          for (int c = 0; c < CP; ++c) v[c] = __uint_as_float(r[c]);
          if ((MODE == 1 || MODE == 2) && a.ref) {
#pragma unroll
            for (int c = 0; c < CP; c += 8) {
              if (c < a.Cout) {
                float f[8];
                unpack8(refq[pu][c / 8], f);
                const float rs = (SPLIT && c >= split) ? a.ref2_slope : a.ref_slope;
#pragma unroll
                for (int u = 0; u < 8; ++u) v[c + u] *= (f[u] > 0.f) ? 1.f : rs;
              }
            }
          }
          fetch_ref(zo + 2 * PF, refq[pu]);
          if (MODE == 2 && a.drop_key) {
            const uint32_t di = di_base + (uint32_t)zo * di_zstride;
#pragma unroll
            for (int c = 0; c < CP; ++c) if (!SPLIT || c < split) v[c] *= 2.f * tem_keep(a.drop_key, di + c);
          }
          if (CP == 8 && MODE == 3) {           // g11 (16 -> 1, linear): one fp32 value per voxel, or the fused inference epilogue of utils.py:109-121
            const float y = a.slope != 1.f ? (v[0] > 0.f ? v[0] : v[0] * a.slope) : v[0];
            const int oz = z0 + zo;
            if (a.st_out) {
              const int cz = oz - a.st_tpad, cy = oy - a.st_tpad, cx = ox - a.st_tpad;
              if (cz >= 0 && cz < a.st_od && cy >= 0 && cy < a.st_od && cx >= 0 && cx < a.st_od) {
                const long long gx = a.st_index[b * 3 + 0] + cx, gy = a.st_index[b * 3 + 1] + cy, gz = a.st_index[b * 3 + 2] + cz;
                if (gx < a.st_OX && gy < a.st_OY && gz < a.st_OZ) a.st_out[(gz * a.st_OY + gy) * a.st_OX + gx] = tem_to_u8_round(y, a.st_mean, a.st_std);
              }
            } else {
              reinterpret_cast<float*>(a.out)[((((long long)b * a.OZ + oz + a.out_off[0]) * a.OY + oy + a.out_off[1]) * a.OX + ox + a.out_off[2]) * a.out_C + a.out_coff] = y;
            }
            continue;
          }
          bf16* const op1 = out_base + (long long)zo * out_zstride;
          bf16* const op2 = out2_base + (long long)zo * out2_zstride;
#pragma unroll
          for (int c = 0; c < CP; c += 8) {
            if (c < a.Cout) {
              bf16* const op = (SPLIT && c >= split) ? op2 : op1;
              float o[8];
              if (MODE >= 1 && a.accumulate) unpack8(*reinterpret_cast<const uint4*>(op + c), o);
              else {
#pragma unroll
                for (int u = 0; u < 8; ++u) o[u] = 0.f;
              }
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                o[u] += v[c + u];
                if (MODE == 0) o[u] = o[u] > 0.f ? o[u] : o[u] * a.slope;      // slope = 1: identity
              }
              uint4 pk;
              pk.x = pack2(o[0], o[1]); pk.y = pack2(o[2], o[3]); pk.z = pack2(o[4], o[5]); pk.w = pack2(o[6], o[7]);
              *reinterpret_cast<uint4*>(op + c) = pk;
            }
          }
        }
      }
      // end of the item: all its MMAs have retired once the last step's barrier has flipped.  The column groups of the
      // virtual output slices -2, -1 (kz = 1, 2 parts of the first two input slices) and nz, nz + 1 (kz = 0, 1 parts of the
      // last two) hold partial sums nobody reads: the even / odd warps wipe them
      mbar_wait(&tfull_bar[nbatch - 1], (uint32_t)(tf_ph >> (nbatch - 1)) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      {
        const uint32_t junk = lane_base + (uint32_t)((ew == 0 ? a.zcap + 2 : a.zcap - nz) * CP);
#pragma unroll
        for (int c = 0; c < 2 * CP; c += 8) tmem_st8_zero(junk + c);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (lane == 0) mbar_arrive(&strip_bar);
      tf_ph ^= (nbatch >= 64) ? ~0ull : ((1ull << nbatch) - 1ull);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols) : "memory");
  }
}

// bf16 UMMA B image [step][k-half][n-group][8 rows][8 elems] with n = (kz, co): N' = 3*CP columns (32 when CP == 8)
struct Pack3Args {
  const float* w; long long ws_tap, ws_in, ws_out;
  int flip, cin, cols, cp, np, spd, cin8;
  bf16* dst; int total;
};
__global__ void pack_weights3_kernel(const Pack3Args a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.total) return;
  int t = i;
  const int e = t & 7; t >>= 3;
  const int r = t & 7; t >>= 3;
  const int ng = a.np >> 3;
  const int g = t % ng; t /= ng;
  const int j = t & 1; t >>= 1;
  const int s = t;
  int tap9, ci;
  if (a.cin8) { tap9 = 2 * s + j; ci = e; }
  else { const int kcs = a.cin >> 4; tap9 = s / kcs; ci = (2 * (s % kcs) + j) * 8 + e; }
  const int n = g * 8 + r;
  const int dz = n / a.cp, co = n % a.cp;
  float v = 0.f;
  if (tap9 < 9 && dz < 3 && co < a.cols) {
    int tap = dz * 9 + tap9;
    if (a.flip) tap = 26 - tap;
    v = a.w[tap * a.ws_tap + (long long)ci * a.ws_in + (long long)co * a.ws_out];
  }
  a.dst[i] = __float2bfloat16_rn(v);
}

int sb_of(int spd) { const int sb = 30 / spd; return sb < 1 ? 1 : (sb > 8 ? 8 : sb); }   // input slices per pipeline step
int cp_of(int cout) { return cout <= 8 ? 8 : (cout <= 16 ? 16 : 32); }
int np_of(int cp) { return cp == 8 ? 32 : 3 * cp; }

}  // namespace

bool tc_conv_supported(const ConvArgs& a) {
  if (a.form != 0 && !(a.form == 1 && a.stride[0] == 1 && a.stride[1] == 1 && a.stride[2] == 1)) return false;
  if (a.k[0] != 3 || a.k[1] != 3 || a.k[2] != 3) return false;
  if (a.stride[0] != 1 || a.stride[1] != 1 || a.stride[2] != 1) return false;
  const bool last = a.Cout == 1 && a.out_dtype == DT_F32 && a.form == 0 && !a.ref && !a.drop_key && !a.accumulate && a.C0 + a.C1 <= 32;   // g11: 16 -> 1, fp32 / fused uint8 output
  if (a.s0.dtype != DT_BF16 || (a.out_dtype != DT_BF16 && !last)) return false;
  if (a.s0.origins || a.use_lut || a.bias) return false;
  const int cin = a.C0 + a.C1;
  if (!(cin == 8 || cin % 16 == 0)) return false;
  if (a.C0 % 8 || a.C1 % 8 || a.s0.C % 8 || a.s0.coff != 0 || a.s0.C != a.C0) return false;
  if (a.C1 && (a.s1.dtype != DT_BF16 || a.s1.C % 8 || a.s1.coff != 0 || a.s1.C != a.C1)) return false;
  if (!last && (a.Cout % 8 || a.Cout > 32 || a.out_C % 8 || a.out_coff % 8)) return false;
  if (a.st_out && !last) return false;
  if ((a.ref || a.drop_key || a.accumulate) && a.slope != 1.f) return false;     // the gradient epilogues carry no activation
  if (a.ref && (a.ref_C % 8 || a.ref_coff % 8)) return false;
  if (a.split) {     // two destinations: whole 8-channel chunks each, both bf16, both with a LeakyReLU' reference
    if (last || a.split % 8 || a.split >= a.Cout || a.out_coff || a.ref_coff || !a.out2 || !a.ref || !a.ref2) return false;
    if (a.out2_C % 8 || a.ref2_C % 8 || a.out_C % 8) return false;
  }
  const size_t smem = ((tc3_packed_bytes(cin, a.Cout) + 1023) & ~(size_t)1023) + (size_t)2 * sb_of(cin == 8 ? 5 : 9 * (cin / 16)) * (cin / 8) * PLANE_STRIDE + 1024;   // two pipeline steps at least
  if (smem > 200 * 1024) return false;
  if (a.conv_off[0] || a.conv_off[1] || a.conv_off[2]) return false;
  return tem_get_encode() != nullptr;
}

size_t tc3_packed_bytes(int cin, int cout) {
  const int spd = (cin == 8) ? 5 : 9 * (cin / 16);
  return (size_t)spd * np_of(cp_of(cout)) * 32;
}

cudaError_t tc3_pack_weights(const ConvArgs& a, bf16* dst, cudaStream_t st) {
  Pack3Args p;
  const int cin = a.C0 + a.C1;
  p.w = a.w; p.ws_tap = a.ws_tap; p.ws_in = a.ws_in; p.ws_out = a.ws_out;
  p.flip = (a.form == 1) ? 1 : 0;
  p.cin = cin; p.cols = a.Cout; p.cp = cp_of(a.Cout); p.np = np_of(p.cp);
  p.cin8 = cin == 8; p.spd = p.cin8 ? 5 : 9 * (cin / 16);
  p.dst = dst; p.total = (int)(tc3_packed_bytes(cin, a.Cout) / 2);
  pack_weights3_kernel<<<(p.total + 255) / 256, 256, 0, st>>>(p); ++g_tem_launches;
  return cudaGetLastError();
}

cudaError_t launch_conv_tc3(const ConvArgs& a, const bf16* wpacked, cudaStream_t st) {
  Tc3Args t; memset(&t, 0, sizeof(t));
  const int cin = a.C0 + a.C1;
  t.B = a.B; for (int i = 0; i < 3; ++i) t.L[i] = a.L[i];
  t.planes0 = a.C0 / 8; t.planes1 = a.C1 / 8;
  const int pad = (a.form == 1) ? 2 : 0;
  for (int i = 0; i < 3; ++i) { t.shift0[i] = a.s0.shift[i] - pad; t.shift1[i] = a.s1.shift[i] - pad; }
  t.cin8 = cin == 8; t.spd = t.cin8 ? 5 : 9 * (cin / 16);
  t.wpacked = wpacked; t.wbytes = (int)tc3_packed_bytes(cin, a.Cout);
  t.ntx = (a.L[2] + TX - 1) / TX; t.nty = (a.L[1] + TY - 1) / TY;
  const int cp = cp_of(a.Cout);
  // accumulator strip: (zcap + 5) column groups of CP columns; 256 columns keep two CTAs per SM when that leaves a
  // useful chunk, otherwise the whole TMEM (one CTA per SM)
  static const char* c16 = getenv("TEM_TC3_COLS16");
  static const char* c8 = getenv("TEM_TC3_COLS8");
  t.tmem_cols = (cp == 8) ? (c8 ? atoi(c8) : 256) : (cp == 16 ? (c16 ? atoi(c16) : 256) : 512);   // measured: 2 CTAs/SM beat a longer z-chunk at CP = 16
  t.zcap = t.tmem_cols / cp - 5;
  if (t.zcap > kMaxChunk) t.zcap = kMaxChunk;
  const long long cols = (long long)a.B * t.ntx * t.nty;
  int nzc = (a.L[0] + t.zcap - 1) / t.zcap;
  static const char* mc_s = getenv("TEM_TC3_MINCHUNK");
  const int min_chunk = mc_s ? atoi(mc_s) : 3;     // measured: small layers are a serial MMA chain per CTA, finer z chunks spread it over more SMs
  static const char* r32_s = getenv("TEM_TC3_RES32");          // debug knob
  const int resident = 148 * ((t.tmem_cols == 512) ? (r32_s ? atoi(r32_s) : 1) : 2);     // CTAs the GPU holds at once: a 512-column strip takes the whole TMEM of an SM
  while (cols * nzc < resident && (a.L[0] + nzc) / (nzc + 1) >= min_chunk) ++nzc;
  t.zc = (a.L[0] + nzc - 1) / nzc; t.nzc = (a.L[0] + t.zc - 1) / t.zc;
  t.out = (bf16*)a.out; t.OZ = a.OZ; t.OY = a.OY; t.OX = a.OX; t.out_C = a.out_C; t.out_coff = a.out_coff;
  for (int i = 0; i < 3; ++i) { t.out_off[i] = a.out_off[i]; t.ref_off[i] = a.ref_off[i]; }
  t.Cout = a.Cout; t.slope = a.slope;
  t.out_f32 = a.out_dtype == DT_F32;
  t.st_out = a.st_out; t.st_index = a.st_index; t.st_tpad = a.st_tpad; t.st_od = a.st_od; t.st_mean = a.st_mean; t.st_std = a.st_std;
  t.st_OZ = a.st_OZ; t.st_OY = a.st_OY; t.st_OX = a.st_OX;
  t.ref = a.ref; t.RZ = a.RZ; t.RY = a.RY; t.RX = a.RX; t.ref_C = a.ref_C; t.ref_coff = a.ref_coff; t.ref_slope = a.ref_slope;
  t.drop_key = a.drop_key; t.accumulate = a.accumulate;
  t.split = a.split; t.out2 = (bf16*)a.out2; t.O2Z = a.O2Z; t.O2Y = a.O2Y; t.O2X = a.O2X; t.out2_C = a.out2_C;
  t.ref2 = a.ref2; t.R2Z = a.R2Z; t.R2Y = a.R2Y; t.R2X = a.R2X; t.ref2_C = a.ref2_C; t.ref2_slope = a.ref2_slope;
  for (int i = 0; i < 3; ++i) { t.out2_off[i] = a.out2_off[i]; t.ref2_off[i] = a.ref2_off[i]; }
  t.dbg = tem_ablation_bits();
  CUtensorMap m0, m1;
  if (!tem_make_map_plane(&m0, &t.merged0, a.s0.p, a.B, a.s0.Z, a.s0.Y, a.s0.X, a.s0.C, HX, HY)) return cudaErrorInvalidValue;
  if (a.C1) { if (!tem_make_map_plane(&m1, &t.merged1, a.s1.p, a.B, a.s1.Z, a.s1.Y, a.s1.X, a.s1.C, HX, HY)) return cudaErrorInvalidValue; }
  else { m1 = m0; t.merged1 = t.merged0; }
  // pipeline step = sb input slices with >= ~24 MMAs between two barrier round trips (see the kernel); the ring holds
  // ~56 KB per CTA in flight (Little's law at the ~2 us TMA round trip), at least two steps
  static const char* sb_s = getenv("TEM_TC3_SB");
  int sb = sb_s ? atoi(sb_s) : sb_of(t.spd);
  if (sb < 1) sb = 1; if (sb > 8) sb = 8;
  if (sb > t.zc + 2) sb = t.zc + 2;
  t.sb = sb;
  const size_t step_bytes = (size_t)sb * (cin / 8) * PLANE_STRIDE;
  static const char* ring_s = getenv("TEM_TC3_RING");
  int ring = ring_s ? atoi(ring_s) : (int)((56 * 1024) / step_bytes);
  if (ring > RING3_MAX) ring = RING3_MAX;
  if (ring < 2) ring = 2;
  t.ring = ring;
  const size_t smem = (((size_t)t.wbytes + 1023) & ~(size_t)1023) + (size_t)ring * step_bytes + 1024;
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  t.items = (int)(cols * t.nzc);
  t.nbmax = (t.zc + 2 + sb - 1) / sb;
  if (t.nbmax > kMaxChunk) return cudaErrorInvalidConfiguration;
  const unsigned grid = (unsigned)(t.items < resident ? t.items : resident);     // persistent: every CTA resident from the start
  static bool attr[14] = {};
#define LAUNCH_TC3(CPV, SPL, MD, IDX)                                                                                   \
  {                                                                                                                     \
    if (!attr[IDX]) { cudaError_t e = cudaFuncSetAttribute(conv3_tc3_kernel<CPV, SPL, MD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); if (e) return e; attr[IDX] = true; } \
    conv3_tc3_kernel<CPV, SPL, MD><<<grid, kTcThreads, smem, st>>>(m0, m1, t);                                          \
  }
#define LAUNCH_TC3_MODE(CPV, SPL, IDX)                                                                                  \
  { if (mode == 2) LAUNCH_TC3(CPV, SPL, 2, IDX) else LAUNCH_TC3(CPV, SPL, 1, IDX + 1) }
  // forward launches (no mask, no dropout, no accumulation, LeakyReLU or linear) take MODE 0; everything else the gradient modes
  const int mode = t.out_f32 ? 3 : (a.drop_key ? 2 : ((a.ref || a.accumulate || a.split) ? 1 : 0));
  if (a.split) { if (cp == 16) LAUNCH_TC3_MODE(16, true, 0) else LAUNCH_TC3_MODE(32, true, 2) }
  else if (mode == 3) LAUNCH_TC3(8, false, 3, 4)
  else if (mode == 0) { if (cp == 8) LAUNCH_TC3(8, false, 0, 5) else if (cp == 16) LAUNCH_TC3(16, false, 0, 6) else LAUNCH_TC3(32, false, 0, 7) }
  else if (cp == 8) LAUNCH_TC3_MODE(8, false, 8) else if (cp == 16) LAUNCH_TC3_MODE(16, false, 10) else LAUNCH_TC3_MODE(32, false, 12)
#undef LAUNCH_TC3_MODE
#undef LAUNCH_TC3
  ++g_tem_launches;
  return cudaGetLastError();
}
