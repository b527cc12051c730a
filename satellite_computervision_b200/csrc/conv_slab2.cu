// Instantiations + cluster launch of the CTA-pair slab kernel (conv_slab2.cuh).
#include "conv_slab2.cuh"

namespace scv {

namespace {
template <int KC, int EPI, int NACC>
cudaError_t slab2_launch_t(const ConvLaunch& L, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(L.grid));
  cfg.blockDim = dim3(slab2_threads(NACC));
  cfg.dynamicSmemBytes = L.smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, conv_slab2_kernel<KC, EPI, NACC>, L.tmA, L.tmB, L.tmOut, L.tmPool, L.p);
}
template <int KC, int EPI>
cudaError_t slab2_launch_nacc(const ConvLaunch& L, cudaStream_t s) {
  return L.nacc == 4 ? slab2_launch_t<KC, EPI, 4>(L, s) : slab2_launch_t<KC, EPI, 2>(L, s);
}
template <int KC, int EPI, int NACC>
cudaError_t slab2_attr() {
  return cudaFuncSetAttribute(conv_slab2_kernel<KC, EPI, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}
}  // namespace

cudaError_t conv_slab2_launch(const ConvLaunch& L, cudaStream_t s) {
  if (L.BN != kSlab2BN || L.p.ntaps != 9 || (L.grid & 1) || (L.p.num_m_tiles & 1) || (L.nacc != 2 && L.nacc != 4))
    return cudaErrorInvalidValue;
  if (L.KC == 64 && L.EPI == EPI_STORE) return slab2_launch_nacc<64, EPI_STORE>(L, s);
  if (L.KC == 64 && L.EPI == EPI_POOL_SKIP) return slab2_launch_nacc<64, EPI_POOL_SKIP>(L, s);
  if (L.KC == 32 && L.EPI == EPI_STORE) return slab2_launch_nacc<32, EPI_STORE>(L, s);
  if (L.KC == 32 && L.EPI == EPI_POOL_SKIP) return slab2_launch_nacc<32, EPI_POOL_SKIP>(L, s);
  return cudaErrorInvalidValue;
}

// CTA pairs that can be resident at once (0 when cluster launch is not available)
int conv_slab2_max_pairs(size_t smem, int nacc) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * 64);
  cfg.blockDim = dim3(slab2_threads(nacc));
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  const cudaError_t e = nacc == 4 ? cudaOccupancyMaxActiveClusters(&n, conv_slab2_kernel<64, EPI_STORE, 4>, &cfg)
                                  : cudaOccupancyMaxActiveClusters(&n, conv_slab2_kernel<64, EPI_STORE, 2>, &cfg);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

cudaError_t conv_slab2_init_attributes() {
  cudaError_t e;
  if ((e = slab2_attr<64, EPI_STORE, 2>()) != cudaSuccess) return e;
  if ((e = slab2_attr<64, EPI_STORE, 4>()) != cudaSuccess) return e;
  if ((e = slab2_attr<64, EPI_POOL_SKIP, 2>()) != cudaSuccess) return e;
  if ((e = slab2_attr<64, EPI_POOL_SKIP, 4>()) != cudaSuccess) return e;
  if ((e = slab2_attr<32, EPI_STORE, 2>()) != cudaSuccess) return e;
  if ((e = slab2_attr<32, EPI_STORE, 4>()) != cudaSuccess) return e;
  if ((e = slab2_attr<32, EPI_POOL_SKIP, 2>()) != cudaSuccess) return e;
  if ((e = slab2_attr<32, EPI_POOL_SKIP, 4>()) != cudaSuccess) return e;
  return cudaSuccess;
}

}  // namespace scv
