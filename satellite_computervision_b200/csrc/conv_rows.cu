// Instantiations + dispatch of the row-streaming tap-packed convolution kernel.
#include "conv_rows.cuh"

namespace scv {

namespace {
template <int KC, int COUT, int EPI>
cudaError_t rows_one(const ConvLaunch& L, cudaStream_t s) {
  SCV_LAUNCH_CHECK(launch_pdl(conv_rows_kernel<KC, COUT, EPI>, L.grid, rows_threads(COUT, EPI), L.smem, s, L.tmA, L.tmB, L.tmOut, L.tmPool, L.p));
  return cudaGetLastError();
}
template <int KC, int COUT>
cudaError_t rows_epi(const ConvLaunch& L, cudaStream_t s) {
  switch (L.EPI) {
    case EPI_STORE: return rows_one<KC, COUT, EPI_STORE>(L, s);
    case EPI_POOL_SKIP: return rows_one<KC, COUT, EPI_POOL_SKIP>(L, s);
    case EPI_HEAD: return rows_one<KC, COUT, EPI_HEAD>(L, s);
  }
  return cudaErrorInvalidValue;
}
template <int KC, int COUT>
cudaError_t rows_attr() {
  const int kMax = 227 * 1024;
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(conv_rows_kernel<KC, COUT, EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(conv_rows_kernel<KC, COUT, EPI_POOL_SKIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax)) != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv_rows_kernel<KC, COUT, EPI_HEAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax);
}
}  // namespace

cudaError_t conv_rows_launch(const ConvLaunch& L, cudaStream_t s) {
  if (L.p.ntaps != 9 || (L.p.W & 1) || (L.p.H & 1)) return cudaErrorInvalidValue;
  if (L.KC == 16 && L.BN == 32) return rows_epi<16, 32>(L, s);  // first layer: 8 stored channels, zero-filled to 16 by TMA
  if (L.KC == 16 && L.BN == 64) return rows_epi<16, 64>(L, s);
  if (L.KC == 32 && L.BN == 32) return rows_epi<32, 32>(L, s);
  if (L.KC == 32 && L.BN == 64) return rows_epi<32, 64>(L, s);
  if (L.KC == 64 && L.BN == 32) return rows_epi<64, 32>(L, s);
  if (L.KC == 64 && L.BN == 64) return rows_epi<64, 64>(L, s);
  return cudaErrorInvalidValue;
}

cudaError_t conv_rows_init_attributes() {
  cudaError_t e;
  if ((e = rows_attr<16, 32>()) != cudaSuccess) return e;
  if ((e = rows_attr<16, 64>()) != cudaSuccess) return e;
  if ((e = rows_attr<32, 32>()) != cudaSuccess) return e;
  if ((e = rows_attr<32, 64>()) != cudaSuccess) return e;
  if ((e = rows_attr<64, 32>()) != cudaSuccess) return e;
  return rows_attr<64, 64>();
}

}  // namespace scv
