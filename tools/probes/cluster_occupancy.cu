// How many thread-block clusters of size 1/2/4/8/16 with ~200 KB of shared memory per CTA can be co-resident on this GPU
// (decides whether a 4- or 8-CTA cluster with weight multicast can cover all 148 SMs).
#include <cuda_runtime.h>
#include <cstdio>
__global__ void k(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    printf("%s: %d SMs\n", prop.name, prop.multiProcessorCount);
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
        printf("cluster size %2d: max active clusters %d (%d SMs) %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
