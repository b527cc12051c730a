// Instantiations + dispatch of the weight-streaming halo-slab convolution kernel.
#include "conv_slabw.cuh"

namespace scv {

cudaError_t conv_slabw_launch(const ConvLaunch& L, cudaStream_t s) {
  if (L.p.ntaps != 9 || L.KC != 64 || (L.BN != 128 && L.BN != 64)) return cudaErrorInvalidValue;
  if (L.BN == 64) {  // Cout = 64 with Cin >= 128 (decoder_1/conv0): resident weights would leave room for two slabs only
    if (L.EPI == EPI_STORE)
      SCV_LAUNCH_CHECK(launch_pdl(conv_slabw_kernel<64, 64, EPI_STORE>, L.grid, kSlabwThreads, L.smem, s, L.tmA, L.tmB, L.tmOut, L.tmPool, L.p));
    else if (L.EPI == EPI_POOL_SKIP)
      SCV_LAUNCH_CHECK(launch_pdl(conv_slabw_kernel<64, 64, EPI_POOL_SKIP>, L.grid, kSlabwThreads, L.smem, s, L.tmA, L.tmB, L.tmOut, L.tmPool, L.p));
    else
      return cudaErrorInvalidValue;
    return cudaGetLastError();
  }
  if (L.EPI == EPI_STORE)
    SCV_LAUNCH_CHECK(launch_pdl(conv_slabw_kernel<64, 128, EPI_STORE>, L.grid, kSlabwThreads, L.smem, s, L.tmA, L.tmB, L.tmOut, L.tmPool, L.p));
  else if (L.EPI == EPI_POOL_SKIP)
    SCV_LAUNCH_CHECK(launch_pdl(conv_slabw_kernel<64, 128, EPI_POOL_SKIP>, L.grid, kSlabwThreads, L.smem, s, L.tmA, L.tmB, L.tmOut, L.tmPool, L.p));
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t conv_slabw_init_attributes() {
  const int kMax = 227 * 1024;
  cudaError_t e = cudaFuncSetAttribute(conv_slabw_kernel<64, 128, EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax);
  if (e != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(conv_slabw_kernel<64, 64, EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(conv_slabw_kernel<64, 64, EPI_POOL_SKIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax)) != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv_slabw_kernel<64, 128, EPI_POOL_SKIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax);
}

}  // namespace scv
