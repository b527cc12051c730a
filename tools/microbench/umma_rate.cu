// Microbenchmark: sustained tcgen05.mma rate (cycles per 128 x N x 16 bf16 MMA, SS operands) on one SM,
// all SMs busy, for N in {32,64,128,256}; A rows 128 B (SW128), B tile N x 64.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I satellite_computervision_b200/csrc \
//        tools/microbench/umma_rate.cu -o gpurun_out/umma_rate
#include <cstdio>
#include "ptx.cuh"
using namespace scv;

template <int N>
__global__ void __launch_bounds__(64, 1) rate_kernel(long long* out, int rounds, int distinct_a) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (9 * 16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0;
  if (warp == 1) {
    if (lane == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc(&slot, N < 32 ? 32 : N);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    const uint32_t a0 = smem_u32(base), b0 = a0 + 9 * 16384;
    const uint64_t da0 = umma_smem_desc(a0, 128), db0 = umma_smem_desc(b0, 128);
    constexpr uint32_t IDESC = umma_idesc_bf16(128, N);
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < 36; ++i) {
          const int tap = distinct_a ? (i / 4) : 0;
          umma_bf16(tm, da0 + (uint64_t)((tap * 16384 + (i & 3) * 32) >> 4), db0 + (uint64_t)(((i & 3) * 32) >> 4), IDESC, 1u);
        }
        umma_commit(&bar);
      }
      __syncwarp();
      while (!mbar_try_wait(&bar, r & 1)) {}
    }
    long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[0] = (t1 - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tm, N < 32 ? 32 : N);
}

template <int N>
void run(long long* d, int distinct) {
  const int rounds = 2000;
  size_t smem = 9 * 16384 + 32768 + 2048;
  cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  rate_kernel<N><<<148, 64, smem>>>(d, 10, distinct);
  rate_kernel<N><<<148, 64, smem>>>(d, rounds, distinct);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("N=%3d distinct_a=%d : %.1f cycles per MMA (128xNx16), %.1f B/clk operands  [%s]\n", N, distinct,
         (double)h / (rounds * 36.0), (4096.0 + N * 32.0) / ((double)h / (rounds * 36.0)), cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  for (int distinct = 0; distinct < 2; ++distinct) {
    run<32>(d, distinct);
    run<64>(d, distinct);
    run<128>(d, distinct);
    run<256>(d, distinct);
  }
  return 0;
}
