// Microbenchmark: sustained tcgen05.mma.cta_group::2 rate (cycles per 256 x N x 16 bf16 MMA, SS operands: each CTA of
// the pair supplies its 128 rows of A and N/2 rows of B from its own shared memory), all SMs busy, against the
// single-CTA 128 x N x 16 rate of umma_rate.cu.  Question it answers: does pairing lower the per-SM operand fetch
// (A 4 KB + B N*16 B per K = 16 step instead of 4 KB + N*32 B) for the N = 64 / N = 96 layers?  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I satellite_computervision_b200/csrc \
//        tools/microbench/umma_rate_2cta.cu -o tools/microbench/build/umma_rate_2cta
#include <cstdio>
#include "ptx.cuh"
using namespace scv;

__device__ __forceinline__ uint32_t cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64, 1) rate2_kernel(long long* out, int rounds) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cta_rank();
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0;
  if (warp == 1) {
    if (lane == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc2(&slot, N < 32 ? 32 : N);
    tmem_relinquish2();
  }
  fence_proxy_async();
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    const uint32_t a0 = smem_u32(base), b0 = a0 + 16384;
    const uint64_t da0 = umma_smem_desc(a0, 128), db0 = umma_smem_desc(b0, 128);
    constexpr uint32_t IDESC = umma_idesc_bf16(256, N);
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      if (rank == 0) {  // the leader issues for the pair; the commit arrives on both CTAs' barriers
        if (elect_one()) {
#pragma unroll
          for (int i = 0; i < 36; ++i)
            umma2_bf16(tm, da0 + (uint64_t)(((i & 3) * 32) >> 4), db0 + (uint64_t)(((i & 3) * 32) >> 4), IDESC, 1u);
          umma2_commit_mc(&bar, 3);
        }
        __syncwarp();
      }
      while (!mbar_try_wait(&bar, r & 1)) {}
    }
    long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[0] = (t1 - t0);
  }
  tc_fence_before();
  cluster_sync();
  if (warp == 1) tmem_dealloc2(tm, N < 32 ? 32 : N);
}

template <int N>
void run(long long* d) {
  const int rounds = 2000;
  size_t smem = 16384 + 32768 + 2048;
  cudaFuncSetAttribute(rate2_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  rate2_kernel<N><<<148, 64, smem>>>(d, 10);
  rate2_kernel<N><<<148, 64, smem>>>(d, rounds);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double cyc = (double)h / (rounds * 36.0);
  printf("2-CTA N=%3d : %.1f cycles per MMA (256xNx16 per pair = 128xNx16 per SM), per-SM operands %.0f B -> %.1f B/clk  [%s]\n", N,
         cyc, 4096.0 + N * 16.0, (4096.0 + N * 16.0) / cyc, cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  run<32>(d);
  run<64>(d);
  run<96>(d);
  run<128>(d);
  run<192>(d);
  run<256>(d);
  return 0;
}
