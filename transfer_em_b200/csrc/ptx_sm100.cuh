// PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk[.tensor]), tcgen05 (MMA / commit / TMEM load), tensor maps.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {      // true in exactly one lane of a converged warp
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// one 8-channel plane of a halo tile; `merged` = the tensor has exactly 8 channels and its map folds (channel, x) into
// one contiguous inner dimension (tem_make_map_c8): one request per tile row instead of one per voxel
__device__ __forceinline__ void tma_load_plane(void* dst, const CUtensorMap* map, uint64_t* bar, int merged, int plane, int x, int y, int z, int b) {
  if (merged) tma_load_5d(dst, map, bar, x * 8, y, z, b, 0);
  else tma_load_5d(dst, map, bar, plane * 8, x, y, z, b);
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // SWIZZLE_NONE (interleave) K-major descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48)
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}


// ---- host: 5-D bf16 tensor map over a dense channels-last [B,Z,Y,X,C] tensor, box = 8 channels x bx x by x 1 x 1
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeTiledFn tem_get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// C == 8: a voxel is 16 B and x-neighbours are contiguous, so (channel, x) is ONE dimension of 8*X elements.  TMA issues one
// request per inner-dimension row of the box: with the plain map that is one 16 B request per voxel (measured ~5 cycles
// each per SM, i.e. < 1 TB/s over the chip); merged, a halo row of bx voxels is a single request of bx*16 B.
static inline bool tem_make_map_c8(CUtensorMap* m, const void* base, int B, int Z, int Y, int X, int bx, int by, int bz = 1) {
  EncodeTiledFn enc = tem_get_encode();
  if (!enc || bx * 8 > 256) return false;
  cuuint64_t dims[5] = {(cuuint64_t)X * 8, (cuuint64_t)Y, (cuuint64_t)Z, (cuuint64_t)B, 1};
  cuuint64_t strides[4] = {(cuuint64_t)X * 16, (cuuint64_t)Y * X * 16, (cuuint64_t)Z * Y * X * 16, (cuuint64_t)B * Z * Y * X * 16};
  cuuint32_t box[5] = {(cuuint32_t)bx * 8, (cuuint32_t)by, (cuuint32_t)bz, 1, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}
// plane map of a [B,Z,Y,X,C] tensor: merged when C == 8 (and the box allows it), per-voxel otherwise
static inline bool tem_make_map_5d(CUtensorMap* m, const void* base, int B, int Z, int Y, int X, int C, int bx, int by, int bz = 1);
static inline bool tem_make_map_plane(CUtensorMap* m, int* merged, const void* base, int B, int Z, int Y, int X, int C, int bx, int by) {
  static const bool off = getenv("TEM_NO_TMA_MERGE") != nullptr;     // debug knob
  if (C == 8 && bx * 8 <= 256 && !off) { *merged = 1; return tem_make_map_c8(m, base, B, Z, Y, X, bx, by); }
  *merged = 0; return tem_make_map_5d(m, base, B, Z, Y, X, C, bx, by);
}

static inline bool tem_make_map_5d(CUtensorMap* m, const void* base, int B, int Z, int Y, int X, int C, int bx, int by, int bz) {
  EncodeTiledFn enc = tem_get_encode();
  if (!enc) return false;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)Z, (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)X * C * 2, (cuuint64_t)Y * X * C * 2, (cuuint64_t)Z * Y * X * C * 2};
  cuuint32_t box[5] = {8, (cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

