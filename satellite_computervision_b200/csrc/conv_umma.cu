// Instantiations + dispatch of the UMMA implicit-GEMM convolution kernels.
#include "conv_slab2.cuh"
#include "conv_slabw.cuh"
#include "conv_umma.cuh"

namespace scv {

namespace {
template <int KC, int BN, int EPI>
cudaError_t launch_ptile(const ConvLaunch& L, cudaStream_t stream);
template <int KC, int BN, int EPI>
cudaError_t launch_one(const ConvLaunch& L, cudaStream_t stream) {
  if (L.slab == 3) return launch_ptile<KC, BN, EPI>(L, stream);
  if (L.slab) {
    constexpr int NTAPS = EPI == EPI_CONVT ? 1 : 9;
    if (L.p.ntaps != NTAPS) return cudaErrorInvalidValue;
    if (L.nacc == 4) {
      if constexpr (slab_nacc_ok(BN, 4))
        SCV_LAUNCH_CHECK(launch_pdl(conv_slab_kernel<KC, BN, EPI, NTAPS, 4>, L.grid, slab_threads(4), L.smem, stream, L.tmA, L.tmB, L.tmOut, L.tmPool, L.p));
      else
        return cudaErrorInvalidValue;
    } else {
      SCV_LAUNCH_CHECK(launch_pdl(conv_slab_kernel<KC, BN, EPI, NTAPS, 2>, L.grid, slab_threads(2), L.smem, stream, L.tmA, L.tmB, L.tmOut, L.tmPool, L.p));
    }
  } else {
    SCV_LAUNCH_CHECK(launch_pdl(conv_umma_kernel<KC, BN, EPI>, L.grid, kConvThreads, L.smem, stream, L.tmA, L.tmB, L.p));
  }
  return cudaGetLastError();
}
template <int KC, int BN, int EPI>
cudaError_t launch_ptile(const ConvLaunch& L, cudaStream_t stream) {
  SCV_LAUNCH_CHECK(launch_pdl(conv_ptile_kernel<KC, BN, EPI>, L.grid, kPtileThreads, L.smem, stream, L.tmA, L.tmB, L.p));
  return cudaGetLastError();
}
template <int KC, int BN, int EPI>
cudaError_t attr_one() {
  const int kMax = 227 * 1024;
  cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel<KC, BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(conv_ptile_kernel<KC, BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax);
  if (e != cudaSuccess) return e;
  constexpr int NTAPS = EPI == EPI_CONVT ? 1 : 9;
  e = cudaFuncSetAttribute(conv_slab_kernel<KC, BN, EPI, NTAPS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax);
  if (e != cudaSuccess) return e;
  if constexpr (slab_nacc_ok(BN, 4))
    e = cudaFuncSetAttribute(conv_slab_kernel<KC, BN, EPI, NTAPS, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax);
  return e;
}

template <int KC, int BN>
cudaError_t launch_epi(const ConvLaunch& L, cudaStream_t s) {
  switch (L.EPI) {
    case EPI_STORE: return launch_one<KC, BN, EPI_STORE>(L, s);
    case EPI_POOL_SKIP: return launch_one<KC, BN, EPI_POOL_SKIP>(L, s);
    case EPI_CONVT: return launch_one<KC, BN, EPI_CONVT>(L, s);
    case EPI_HEAD:
      if constexpr (BN <= 128) return launch_one<KC, BN, EPI_HEAD>(L, s);
      return cudaErrorInvalidValue;
  }
  return cudaErrorInvalidValue;
}
template <int KC>
cudaError_t launch_bn(const ConvLaunch& L, cudaStream_t s) {
  switch (L.BN) {
    case 32: return launch_epi<KC, 32>(L, s);
    case 64: return launch_epi<KC, 64>(L, s);
    case 128: return launch_epi<KC, 128>(L, s);
    case 256: return launch_epi<KC, 256>(L, s);
  }
  return cudaErrorInvalidValue;
}

template <int KC, int BN>
cudaError_t attr_epi() {
  cudaError_t e;
  if ((e = attr_one<KC, BN, EPI_STORE>()) != cudaSuccess) return e;
  if ((e = attr_one<KC, BN, EPI_POOL_SKIP>()) != cudaSuccess) return e;
  if ((e = attr_one<KC, BN, EPI_CONVT>()) != cudaSuccess) return e;
  if constexpr (BN <= 128) {
    if ((e = attr_one<KC, BN, EPI_HEAD>()) != cudaSuccess) return e;
  }
  return cudaSuccess;
}
template <int KC>
cudaError_t attr_bn() {
  cudaError_t e;
  if ((e = attr_epi<KC, 32>()) != cudaSuccess) return e;
  if ((e = attr_epi<KC, 64>()) != cudaSuccess) return e;
  if ((e = attr_epi<KC, 128>()) != cudaSuccess) return e;
  if ((e = attr_epi<KC, 256>()) != cudaSuccess) return e;
  return cudaSuccess;
}
}  // namespace

// KC == 8 exists only as the slab kernel of a 3x3 first layer (EPI_STORE / EPI_POOL_SKIP).
template <int BN>
cudaError_t launch_kc8(const ConvLaunch& L, cudaStream_t s) {
  if (!L.slab || L.p.ntaps != 9) return cudaErrorInvalidValue;
  if constexpr (slab_nacc_ok(BN, 4)) {
    if (L.nacc == 4) {
      if (L.EPI == EPI_STORE)
        SCV_LAUNCH_CHECK(launch_pdl(conv_slab_kernel<8, BN, EPI_STORE, 9, 4>, L.grid, slab_threads(4), L.smem, s, L.tmA, L.tmB, L.tmOut, L.tmPool, L.p));
      else if (L.EPI == EPI_POOL_SKIP)
        SCV_LAUNCH_CHECK(launch_pdl(conv_slab_kernel<8, BN, EPI_POOL_SKIP, 9, 4>, L.grid, slab_threads(4), L.smem, s, L.tmA, L.tmB, L.tmOut, L.tmPool, L.p));
      else
        return cudaErrorInvalidValue;
      return cudaGetLastError();
    }
  }
  if (L.EPI == EPI_STORE)
    SCV_LAUNCH_CHECK(launch_pdl(conv_slab_kernel<8, BN, EPI_STORE, 9, 2>, L.grid, slab_threads(2), L.smem, s, L.tmA, L.tmB, L.tmOut, L.tmPool, L.p));
  else if (L.EPI == EPI_POOL_SKIP)
    SCV_LAUNCH_CHECK(launch_pdl(conv_slab_kernel<8, BN, EPI_POOL_SKIP, 9, 2>, L.grid, slab_threads(2), L.smem, s, L.tmA, L.tmB, L.tmOut, L.tmPool, L.p));
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}
template <int BN>
cudaError_t attr_kc8() {
  const int kMax = 227 * 1024;
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(conv_slab_kernel<8, BN, EPI_STORE, 9, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(conv_slab_kernel<8, BN, EPI_POOL_SKIP, 9, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax)) != cudaSuccess) return e;
  if constexpr (slab_nacc_ok(BN, 4)) {
    if ((e = cudaFuncSetAttribute(conv_slab_kernel<8, BN, EPI_STORE, 9, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(conv_slab_kernel<8, BN, EPI_POOL_SKIP, 9, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax)) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t conv_launch(const ConvLaunch& L, cudaStream_t stream) {
  if (L.slab < 0) return cudaSuccess;  // folded into the previous (fused) launch
  if (L.slab == 5) return conv_fused_launch(L, stream);
  if (L.slab == 2) return conv_rows_launch(L, stream);
  if (L.slab == 4) return conv_slabw_launch(L, stream);
  if (L.slab == 6) return conv_slab2_launch(L, stream);
  if (L.KC == 8) {
    switch (L.BN) {
      case 32: return launch_kc8<32>(L, stream);
      case 64: return launch_kc8<64>(L, stream);
      case 128: return launch_kc8<128>(L, stream);
    }
    return cudaErrorInvalidValue;
  }
  switch (L.KC) {
    case 16: return launch_bn<16>(L, stream);
    case 32: return launch_bn<32>(L, stream);
    case 64: return launch_bn<64>(L, stream);
  }
  return cudaErrorInvalidValue;
}

// cudaFuncSetAttribute is per device: one engine per GPU and several engines per process must each opt in.
cudaError_t conv_init_attributes() {
  static bool done[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
  if ((e = attr_kc8<32>()) != cudaSuccess) return e;
  if ((e = attr_kc8<64>()) != cudaSuccess) return e;
  if ((e = attr_kc8<128>()) != cudaSuccess) return e;
  if ((e = attr_bn<16>()) != cudaSuccess) return e;
  if ((e = attr_bn<32>()) != cudaSuccess) return e;
  if ((e = attr_bn<64>()) != cudaSuccess) return e;
  if ((e = conv_rows_init_attributes()) != cudaSuccess) return e;
  if ((e = conv_slabw_init_attributes()) != cudaSuccess) return e;
  if ((e = conv_fused_init_attributes()) != cudaSuccess) return e;
  if ((e = conv_slab2_init_attributes()) != cudaSuccess) return e;
  if (dev >= 0 && dev < 64) done[dev] = true;
  return cudaSuccess;
}

}  // namespace scv
