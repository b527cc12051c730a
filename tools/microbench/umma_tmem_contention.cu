// Microbenchmark: does epilogue-style TMEM traffic (tcgen05.ld / tcgen05.st from other warps, other columns)
// slow a stream of 128 x 96 x 16 UMMAs?  Prints cycles per MMA for: no epilogue warps, ld only, ld + st.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I satellite_computervision_b200/csrc \
//             tools/microbench/umma_tmem_contention.cu -o tools/microbench/build/umma_tmem_contention
#include <cstdio>
#include "ptx.cuh"
using namespace scv;

__device__ __forceinline__ void st32_zero(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
      "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(0u) : "memory");
}

template <int N>
__global__ void __launch_bounds__(64 + 512, 1) k(long long* out, int rounds, int mode, int groups, int smem_traffic) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (4 * 34816 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0;
  if (threadIdx.x == 0) stop = 0;
  if (warp == 1) {
    if (lane == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    const uint32_t a0 = smem_u32(base), b0 = a0 + 4 * 34816;
    const uint64_t db0 = umma_smem_desc(b0, 64);
    constexpr uint32_t IDESC = umma_idesc_bf16(128, N);
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      if (elect_one()) {
#pragma unroll
        for (int row = 0; row < 8; ++row) {
          const uint64_t da0 = umma_smem_desc(a0 + (row & 3) * 34816, 64);
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
              umma_bf16(tm + (row & 1) * 32, da0 + (uint64_t)((kx * 64 + kk * 32) >> 4), db0 + (uint64_t)((kx * 3 * 32 * 64 + kk * 32) >> 4), IDESC, 1u);
        }
        umma_commit(&bar);
      }
      __syncwarp();
      while (!mbar_try_wait(&bar, r & 1)) {}
    }
    long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[0] = (t1 - t0);
    if (lane == 0) stop = 1;
  } else if (warp >= 2 && warp < 2 + 4 * groups && mode > 0) {
    // epilogue-like traffic on columns 256.. (never touched by the MMAs)
    const uint32_t t = tm + 256 + ((warp - 2) >> 2) * 64 + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    uint8_t* stage = base + 4 * 34816 + 40960 + (warp - 2) * 1024;
    while (!stop) {
      uint32_t r[32];
      tmem_ld32(t, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc += r[j];
      if (mode > 1) {
        st32_zero(t);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      if (smem_traffic) {
#pragma unroll
        for (int j = 0; j < 4; ++j) sts128(smem_u32(stage) + lane * 32 + (j & 1) * 16, acc, acc, acc, acc);
      }
    }
    if (acc == 0xdeadbeef) out[1] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tm, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  const int rounds = 1000;
  size_t smem = 4 * 34816 + 65536 + 2048;
  cudaFuncSetAttribute(k<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int smt = 0; smt < 2; ++smt)
    for (int mode = 0; mode < 3; ++mode)
      for (int groups = 1; groups <= 4; groups += 3) {
        k<96><<<148, 64 + 512, smem>>>(d, 10, mode, groups, smt);
        k<96><<<148, 64 + 512, smem>>>(d, rounds, mode, groups, smt);
        cudaError_t e = cudaDeviceSynchronize();
        long long h = 0;
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("N=96 mode=%d (0 none, 1 ld, 2 ld+st) epilogue groups=%d smem stores=%d : %6.1f cycles per MMA [%s]\n", mode, groups, smt,
               (double)h / (rounds * 48.0), cudaGetErrorString(e));
      }
  return 0;
}
