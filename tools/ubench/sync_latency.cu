// Micro-benchmarks of the synchronisation primitives the tcgen05 pipelines are built from (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sync_latency sync_latency.cu && ./sync_latency
// Prints cycles per iteration for: (1) mbarrier ping-pong between two warps, (2) tcgen05.commit -> mbarrier wait with no MMA
// outstanding (same warp), (3) back-to-back tcgen05.commit issue, (4) producer/consumer ring handshake with commits
// (the skeleton of conv3_tc3 / wgrad_tc), (5) tcgen05.ld + wait::ld, (6) a tiny MMA + commit + wait round trip.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../transfer_em_b200/csrc/ptx_sm100.cuh"

__device__ __forceinline__ long long clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)); return c; }

__global__ void k_pingpong(long long* out, int iters) {
  __shared__ uint64_t a, b;
  if (threadIdx.x == 0) { mbar_init(&a, 1); mbar_init(&b, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = clk();
  if (warp == 0) { uint32_t ph = 0; for (int i = 0; i < iters; ++i) { if (lane == 0) mbar_arrive(&a); mbar_wait(&b, ph); ph ^= 1; } }
  else if (warp == 1) { uint32_t ph = 0; for (int i = 0; i < iters; ++i) { mbar_wait(&a, ph); ph ^= 1; if (lane == 0) mbar_arrive(&b); } }
  long long t1 = clk();
  if (threadIdx.x == 0) out[0] = (t1 - t0) / iters;
}

__global__ void k_commit_wait(long long* out, int iters) {
  __shared__ uint64_t a; __shared__ uint32_t tb;
  if (threadIdx.x == 0) { mbar_init(&a, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tb)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t ph = 0;
    long long t0 = clk();
    for (int i = 0; i < iters; ++i) { if (elect_one()) umma_commit(&a); __syncwarp(); mbar_wait(&a, ph); ph ^= 1; }
    long long t1 = clk();
    if (threadIdx.x == 0) out[1] = (t1 - t0) / iters;
    // back-to-back commits to 8 barriers, waiting only at the end of each group
    __shared__ uint64_t bars[8];
    if (threadIdx.x == 0) { for (int j = 0; j < 8; ++j) mbar_init(&bars[j], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    ph = 0;
    t0 = clk();
    for (int i = 0; i < iters / 8; ++i) {
      if (elect_one()) { for (int j = 0; j < 8; ++j) umma_commit(&bars[j]); }
      __syncwarp();
      for (int j = 0; j < 8; ++j) mbar_wait(&bars[j], ph);
      ph ^= 1;
    }
    t1 = clk();
    if (threadIdx.x == 0) out[2] = (t1 - t0) / (iters / 8 * 8);
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(32u) : "memory");
  }
}

// ring handshake: producer warp (arrive on full), "MMA" warp (wait full, commit empty [+ commit tfull]), epilogue warp (wait tfull, tcgen05.ld)
template <int RING>
__global__ void k_ring(long long* out, int iters, int do_ld, int two_commits, int slot) {
  __shared__ uint64_t full[RING], empty[RING], tfull[RING], tfree[RING]; __shared__ uint32_t tb;
  if (threadIdx.x == 0) { for (int j = 0; j < RING; ++j) { mbar_init(&full[j], 1); mbar_init(&empty[j], 1); mbar_init(&tfull[j], 1); mbar_init(&tfree[j], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tb)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  long long t0 = clk();
  if (warp == 0) {
    if (lane == 0) { int s = 0; uint32_t ph = 0; for (int i = 0; i < iters; ++i) { mbar_wait(&empty[s], ph ^ 1u); mbar_arrive(&full[s]); if (++s == RING) { s = 0; ph ^= 1u; } } }
  } else if (warp == 1) {
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&full[s], ph);
      if (two_commits) mbar_wait(&tfree[s], ph ^ 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) { umma_commit(&empty[s]); if (two_commits) umma_commit(&tfull[s]); }
      __syncwarp();
      if (++s == RING) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 2 && two_commits) {
    int s = 0; uint32_t ph = 0; uint32_t r[8]; uint32_t acc = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&tfull[s], ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (do_ld) { tmem_ld8(tb + ((uint32_t)(64) << 16), r); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += r[0]; }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (lane == 0) mbar_arrive(&tfree[s]);
      if (++s == RING) { s = 0; ph ^= 1u; }
    }
    if (acc == 12345) out[15] = acc;
  }
  long long t1 = clk();
  if (threadIdx.x == 32) out[slot] = (t1 - t0) / iters;
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(64u) : "memory");
}

// one M128 x N32 x K16 MMA per iteration (operands = whatever is in shared memory), commit, wait: full MMA round trip
__global__ void k_mma_roundtrip(long long* out, int iters, int mmas_per_iter) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t a; __shared__ uint32_t tb;
  if (threadIdx.x == 0) { mbar_init(&a, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tb)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t ad = umma_desc(smem_u32(sm), 128, 256), bd = umma_desc(smem_u32(sm) + 8192, 128, 256);
    uint32_t ph = 0;
    long long t0 = clk();
    for (int i = 0; i < iters; ++i) {
      if (elect_one()) { for (int j = 0; j < mmas_per_iter; ++j) umma_bf16(tb, ad, bd, idesc, 1u); umma_commit(&a); }
      __syncwarp(); mbar_wait(&a, ph); ph ^= 1;
    }
    long long t1 = clk();
    if (threadIdx.x == 0) out[mmas_per_iter == 1 ? 8 : (mmas_per_iter == 5 ? 9 : 10)] = (t1 - t0) / iters;
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(64u) : "memory");
  }
}

int main() {
  long long* d; cudaMalloc(&d, 16 * sizeof(long long)); cudaMemset(d, 0, 16 * sizeof(long long));
  const int N = 4096;
  k_pingpong<<<1, 64>>>(d, N);
  k_commit_wait<<<1, 64>>>(d, N);
  k_ring<8><<<1, 96>>>(d, N, 0, 0, 3);       // full/empty only, commit as the "consumer release"
  k_ring<8><<<1, 96>>>(d, N, 0, 1, 4);       // + tfull / tfree with an epilogue warp (no tcgen05.ld)
  k_ring<8><<<1, 96>>>(d, N, 1, 1, 5);       // + tcgen05.ld in the epilogue warp
  k_ring<2><<<1, 96>>>(d, N, 1, 1, 6);       // shallow ring
  cudaFuncSetAttribute(k_mma_roundtrip, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  k_mma_roundtrip<<<1, 64, 32768>>>(d, N, 1);
  k_mma_roundtrip<<<1, 64, 32768>>>(d, N, 5);
  k_mma_roundtrip<<<1, 64, 32768>>>(d, N, 20);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("status %s\n", cudaGetErrorString(e));
  printf("mbarrier ping-pong round trip (2 warps)          : %lld cycles\n", h[0]);
  printf("tcgen05.commit -> mbarrier wait (no MMA pending)  : %lld cycles\n", h[1]);
  printf("tcgen05.commit x8 back to back, per commit        : %lld cycles\n", h[2]);
  printf("ring<8> full/empty, release by commit, per slot   : %lld cycles\n", h[3]);
  printf("ring<8> + tfull/tfree epilogue hop, per slot      : %lld cycles\n", h[4]);
  printf("ring<8> + tcgen05.ld in the epilogue, per slot    : %lld cycles\n", h[5]);
  printf("ring<2> + tcgen05.ld in the epilogue, per slot    : %lld cycles\n", h[6]);
  printf("1 MMA (M128 N32 K16) + commit + wait              : %lld cycles\n", h[8]);
  printf("5 MMAs + commit + wait                            : %lld cycles\n", h[9]);
  printf("20 MMAs + commit + wait                           : %lld cycles\n", h[10]);
  return 0;
}
