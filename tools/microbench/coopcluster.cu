// Does the driver accept a launch that is both cooperative (co-residency guaranteed or refused) and clustered?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o coopcluster coopcluster.cu && ./coopcluster
// Used by the hybrid halo transport: clusters of (cx, 1) CTAs that spin on each other through global memory.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256, 1) probe(int* out)
{
    extern __shared__ unsigned char smem[];
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) *out = 1;
    (void)smem;
}

static const char* try_launch(int cx, int clusters, bool coop, size_t smem, int* d)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cx, clusters, 1);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cx; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative;
    at[1].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = coop ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, probe, d);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaGetLastError();
    return cudaGetErrorName(e);
}

int main()
{
    int* d = nullptr;
    cudaMalloc(&d, 4);
    const size_t smem = 200 * 1024;                                     // one CTA per SM, like the fused kernels
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cx : {2, 5, 8, 9}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cx, 64, 1); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cx; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int fit = 0;
        cudaOccupancyMaxActiveClusters(&fit, probe, &cfg);
        printf("clusters of %d: occupancy query says %d co-resident; cluster + cooperative launch of %d clusters: %s, of %d clusters: %s; plain cluster launch of %d: %s\n",
               cx, fit, fit, try_launch(cx, fit, true, smem, d), fit + 1, try_launch(cx, fit + 1, true, smem, d), fit + 1, try_launch(cx, fit + 1, false, smem, d));
    }
    return 0;
}
