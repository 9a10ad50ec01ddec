// How many thread-block clusters of size C (320 threads, 225 KB of dynamic shared memory per CTA: the
// CTA-pair GEMM's footprint) are co-resident on this GPU?  nvcc -arch=sm_100a -o probe probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
  const int smem = 225 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int c : {1, 2, 3, 4, 5, 6, 8, 10, 12, 16}) {
    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(c * 64); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute a{}; a.id = cudaLaunchAttributeClusterDimension; a.val.clusterDim.x = c; a.val.clusterDim.y = 1; a.val.clusterDim.z = 1;
    cfg.attrs = &a; cfg.numAttrs = 1;
    int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster %2d: %3d clusters = %3d of %d SMs (%s)\n", c, n, n * c, sms, cudaGetErrorString(e));
    cudaGetLastError();
  }
  return 0;
}
