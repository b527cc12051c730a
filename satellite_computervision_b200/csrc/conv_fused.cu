// Instantiations + cluster launch of the fused two-conv row kernel (conv_fused.cuh).
#include "conv_fused.cuh"

namespace scv {

namespace {
template <int KC1, int EPI2>
cudaError_t fused_launch_t(const ConvLaunch& L, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(L.grid));
  cfg.blockDim = dim3(f2_threads(EPI2));
  cfg.dynamicSmemBytes = L.smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kF2Cluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, conv_fused2_kernel<KC1, EPI2>, L.tmA, L.tmB, L.tmB2, L.tmOut, L.tmPool, L.p);
}
}  // namespace

cudaError_t conv_fused_launch(const ConvLaunch& L, cudaStream_t s) {
  if (L.p.W != kF2Cluster * kRowsPx || (L.p.H & 1) || L.grid % kF2Cluster) return cudaErrorInvalidValue;
  if (L.KC == 64 && L.EPI == EPI_HEAD) return fused_launch_t<64, EPI_HEAD>(L, s);          // decoder tail
  if (L.KC == 16 && L.EPI == EPI_POOL_SKIP) return fused_launch_t<16, EPI_POOL_SKIP>(L, s);  // first encoder block
  return cudaErrorInvalidValue;
}

// Clusters of three 227 KB CTAs that can be resident at once (GPC boundaries make this less than 148 / 3 on some parts).
int conv_fused_max_clusters(size_t smem, int epi2) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kF2Cluster * 64);
  cfg.blockDim = dim3(f2_threads(epi2));
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kF2Cluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  const cudaError_t e = epi2 == EPI_HEAD ? cudaOccupancyMaxActiveClusters(&n, conv_fused2_kernel<64, EPI_HEAD>, &cfg)
                                         : cudaOccupancyMaxActiveClusters(&n, conv_fused2_kernel<16, EPI_POOL_SKIP>, &cfg);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

cudaError_t conv_fused_init_attributes() {
  cudaError_t e = cudaFuncSetAttribute(conv_fused2_kernel<64, EPI_HEAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv_fused2_kernel<16, EPI_POOL_SKIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

}  // namespace scv
