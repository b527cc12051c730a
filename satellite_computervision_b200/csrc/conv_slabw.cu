// Instantiations + dispatch of the weight-streaming halo-slab convolution kernel.
#include "conv_slabw.cuh"

namespace scv {

cudaError_t conv_slabw_launch(const ConvLaunch& L, cudaStream_t s) {
  if (L.p.ntaps != 9 || L.KC != 64 || L.BN != 128) return cudaErrorInvalidValue;
  if (L.EPI == EPI_STORE)
    conv_slabw_kernel<64, 128, EPI_STORE><<<L.grid, kSlabwThreads, L.smem, s>>>(L.tmA, L.tmB, L.tmOut, L.tmPool, L.p);
  else if (L.EPI == EPI_POOL_SKIP)
    conv_slabw_kernel<64, 128, EPI_POOL_SKIP><<<L.grid, kSlabwThreads, L.smem, s>>>(L.tmA, L.tmB, L.tmOut, L.tmPool, L.p);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t conv_slabw_init_attributes() {
  const int kMax = 227 * 1024;
  cudaError_t e = cudaFuncSetAttribute(conv_slabw_kernel<64, 128, EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv_slabw_kernel<64, 128, EPI_POOL_SKIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax);
}

}  // namespace scv
