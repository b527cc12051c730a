// Microbenchmark for the row-streaming kernel's UMMA pattern: cycles per tcgen05.mma (128 x N x 16, bf16, SS)
// for N in {32,64,96,128,192}, operand rows of 64 B (SWIZZLE_64B) or 128 B (SWIZZLE_128B), with
//   shiftA : A start address moved by kx pixel rows (the three horizontal taps of one 130-pixel halo row)
//   rotD   : the accumulator column address advances by N/3 columns per "row" (ring of row accumulators)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I satellite_computervision_b200/csrc \
//             tools/microbench/umma_rows_rate.cu -o gpurun_out/umma_rows_rate
#include <cstdio>
#include "ptx.cuh"
using namespace scv;

template <int N, int ROWB>
__global__ void __launch_bounds__(64, 1) rate_kernel(long long* out, int rounds, int shiftA, int rotD) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (4 * 34816 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0;
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
  constexpr int KS = ROWB / 32;       // k steps per tap
  constexpr int PER_ROW = 3 * KS;     // UMMAs per input row
  constexpr int ROWS = 8;             // rows per round
  if (warp == 1) {
    const uint32_t a0 = smem_u32(base), b0 = a0 + 4 * 34816;
    const uint64_t db0 = umma_smem_desc(b0, ROWB);
    constexpr uint32_t IDESC = umma_idesc_bf16(128, N);
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      if (elect_one()) {
#pragma unroll
        for (int row = 0; row < ROWS; ++row) {
          const uint64_t da0 = umma_smem_desc(a0 + (row & 3) * 34816, ROWB);
          const uint32_t d = tm + (rotD ? ((r * ROWS + row) % 4) * (N / 3) : 0);
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
#pragma unroll
            for (int k = 0; k < KS; ++k)
              umma_bf16(d, da0 + (uint64_t)(((shiftA ? kx * ROWB : 0) + k * 32) >> 4),
                        db0 + (uint64_t)((kx * 3 * (N / 3) * ROWB + k * 32) >> 4), IDESC, 1u);
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
  if (warp == 1) tmem_dealloc(tm, 512);
  (void)PER_ROW;
}

template <int N, int ROWB>
void run(long long* d, int shiftA, int rotD) {
  const int rounds = 1000;
  size_t smem = 4 * 34816 + 65536 + 2048;
  cudaFuncSetAttribute(rate_kernel<N, ROWB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  rate_kernel<N, ROWB><<<148, 64, smem>>>(d, 10, shiftA, rotD);
  rate_kernel<N, ROWB><<<148, 64, smem>>>(d, rounds, shiftA, rotD);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / (rounds * 8.0 * 3 * (ROWB / 32));
  printf("N=%3d rowB=%3d shiftA=%d rotD=%d : %6.1f cycles per MMA, %6.1f B/clk operands  [%s]\n", N, ROWB, shiftA, rotD, per,
         (4096.0 + N * 32.0) / per, cudaGetErrorString(e));
}

template <int ROWB>
void sweep(long long* d) {
  for (int mode = 0; mode < 4; ++mode) {
    const int shiftA = mode & 1, rotD = mode >> 1;
    run<32, ROWB>(d, shiftA, 0);
    run<64, ROWB>(d, shiftA, 0);
    run<96, ROWB>(d, shiftA, rotD);
    run<128, ROWB>(d, shiftA, 0);
    run<192, ROWB>(d, shiftA, rotD);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  sweep<64>(d);
  sweep<128>(d);
  return 0;
}
