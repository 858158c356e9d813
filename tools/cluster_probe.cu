// How many thread-block clusters of each size can be co-resident on this GPU when a CTA needs
// (almost) a whole SM?  Answers which cluster shapes can cover all SMs (GPC geometry).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_probe tools/cluster_probe.cu && ./cluster_probe
#include <cuda_runtime.h>
#include <cstdio>

__global__ void __launch_bounds__(320, 1) probe_kernel(int* out) {
  extern __shared__ unsigned char smem[];
  if (out && threadIdx.x == 0 && blockIdx.x == 0) out[0] = smem[0];
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  printf("%s: %d SMs\n", prop.name, prop.multiProcessorCount);
  const int smem = 225 * 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs = 1; cs <= 16; ++cs) {
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(320);
    cfg.gridDim = dim3(cs * (prop.multiProcessorCount / cs));
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe_kernel, &cfg);
    if (e != cudaSuccess) { printf("cluster size %2d: %s\n", cs, cudaGetErrorString(e)); cudaGetLastError(); continue; }
    printf("cluster size %2d: %3d clusters = %3d SMs\n", cs, n, n * cs);
  }
  return 0;
}
