// Instantiations + cluster launch of the fused two-conv row kernel (conv_fused.cuh).
#include "conv_fused.cuh"

namespace scv {

namespace {
template <int KC1>
cudaError_t fused_launch_t(const ConvLaunch& L, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(L.grid));
  cfg.blockDim = dim3(kF2Threads);
  cfg.dynamicSmemBytes = L.smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kF2Cluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, conv_fused2_kernel<KC1>, L.tmA, L.tmB, L.tmB2, L.p);
}
}  // namespace

cudaError_t conv_fused_launch(const ConvLaunch& L, cudaStream_t s) {
  if (L.p.W != kF2Cluster * kRowsPx || (L.p.H & 1) || L.grid % kF2Cluster) return cudaErrorInvalidValue;
  if (L.KC == 64) return fused_launch_t<64>(L, s);
  return cudaErrorInvalidValue;
}

// Clusters of three 227 KB CTAs that can be resident at once (GPC boundaries make this less than 148 / 3 on some parts).
int conv_fused_max_clusters(size_t smem) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kF2Cluster * 64);
  cfg.blockDim = dim3(kF2Threads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kF2Cluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, conv_fused2_kernel<64>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

cudaError_t conv_fused_init_attributes() {
  return cudaFuncSetAttribute(conv_fused2_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

}  // namespace scv
