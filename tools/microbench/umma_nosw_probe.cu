// Probe: semantics of the no-swizzle K-major UMMA smem descriptor (LBO / SBO roles).
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include "ptx.cuh"
using namespace scv;

__device__ uint64_t desc_raw(uint32_t saddr, uint32_t f16, uint32_t f32) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((f16 >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((f32 >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// variant 0: field@16 = K-half distance (LBO), field@32 = 8-row-group distance (SBO); variant 1: swapped
__global__ void probe(float* out, int variant, int a_kh, int a_grp, int b_kh, int b_grp) {
  __shared__ __align__(1024) uint8_t smA[16384];
  __shared__ __align__(1024) uint8_t smB[8192];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 16384 / 2; i += blockDim.x) reinterpret_cast<__nv_bfloat16*>(smA)[i] = __float2bfloat16(0.f);
  for (int i = threadIdx.x; i < 8192 / 2; i += blockDim.x) reinterpret_cast<__nv_bfloat16*>(smB)[i] = __float2bfloat16(0.f);
  __syncthreads();
  // A is a window into a pixel array P[px][8 ch] (16 B per pixel): byte b of smA belongs to pixel b/16
  for (int i = threadIdx.x; i < 16384 / 2; i += blockDim.x) {
    const int px = i / 8, c = i % 8;
    reinterpret_cast<__nv_bfloat16*>(smA)[i] = __float2bfloat16((float)((px % 5) + (c % 3)));
  }
  for (int i = threadIdx.x; i < 32 * 16; i += blockDim.x) {
    const int n = i / 16, k = i % 16;
    const int off = (n / 8) * b_grp + (n % 8) * 16 + (k / 8) * b_kh + (k % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(smB + off) = __float2bfloat16((float)((n % 4) + (k % 2)));
  }
  if (warp == 0) {
    if (lane == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc(&slot, 32);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    if (elect_one()) {
      const uint64_t da = variant == 0 ? desc_raw(smem_u32(smA), a_kh, a_grp) : desc_raw(smem_u32(smA), a_grp, a_kh);
      const uint64_t db = variant == 0 ? desc_raw(smem_u32(smB), b_kh, b_grp) : desc_raw(smem_u32(smB), b_grp, b_kh);
      umma_bf16(tm, da, db, umma_idesc_bf16(128, 32), 0u);
      umma_commit(&bar);
    }
    __syncwarp();
  }
  while (!mbar_try_wait(&bar, 0)) {}
  tc_fence_after();
  if (warp < 4) {
    uint32_t raw[32];
    tmem_ld32(tm + ((uint32_t)(warp * 32) << 16), raw);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 32 + j] = __uint_as_float(raw[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 32);
}

int main(int argc, char** argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1;
  float* d; cudaMalloc(&d, 128 * 32 * 4);
  static float h[128 * 32];
  // (a_kh, a_grp, b_kh, b_grp) layouts to try: standard (K halves 128 B apart inside a 256 B group stride), and
  // the conv layout (K half = +16 B, group = 160 B)
  int cfgs[3][4] = {{2048, 128, 512, 128}, {128, 256, 128, 256}, {16, 160, 512, 128}};
  for (int ci = 0; ci < 3; ++ci)
    for (int variant = 0; variant < 1; ++variant) {
      if (only >= 0 && ci != only) continue;
      auto& c = cfgs[ci];
      cudaMemset(d, 0, sizeof h);
      probe<<<1, 128>>>(d, variant, c[0], c[1], c[2], c[3]);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
      int bad = 0; float maxv = 0;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) {
        float ref = 0;
        for (int k = 0; k < 16; ++k) {
          const int byte = (m / 8) * c[1] + (m % 8) * 16 + (k / 8) * c[0] + (k % 8) * 2;
          const int px = byte / 16, ch = (byte % 16) / 2;
          ref += (float)((px % 5) + (ch % 3)) * (float)((n % 4) + (k % 2));
        }
        if (h[m * 32 + n] != ref) ++bad;
        maxv = fmaxf(maxv, fabsf(h[m * 32 + n]));
      }
      printf("layout a_kh=%d a_grp=%d b_kh=%d b_grp=%d variant=%d (0: field16=K-half,field32=group): mismatches=%d max|d|=%.1f d[1][1]=%.1f d[9][2]=%.1f [%s]\n",
             c[0], c[1], c[2], c[3], variant, bad, maxv, h[1 * 32 + 1], h[9 * 32 + 2], cudaGetErrorString(e));
    }
  return 0;
}
