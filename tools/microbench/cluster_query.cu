// How many thread-block clusters of the batched kernel's shape (384 threads, ~210 KB of dynamic shared memory, one CTA per
// SM) can be co-resident on this GPU, per cluster size?  nvcc -arch=sm_100a -o cluster_query cluster_query.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(384, 1) probe(int* out) { if (out) out[blockIdx.x] = 1; }
int main() {
    const int smem = 215040;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    printf("SMs %d\n", prop.multiProcessorCount);
    for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((prop.multiProcessorCount / cs) * cs);
        cfg.blockDim = dim3(384);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe, &cfg);
        printf("cluster size %2d: max active clusters %d (%d CTAs) %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
