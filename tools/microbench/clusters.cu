// Query how many clusters of each size can be co-resident for a 1-CTA-per-SM kernel (255 regs) on this GPU.
#include <cuda_runtime.h>
#include <cstdio>
__global__ void __launch_bounds__(256, 1) big(float* out, int n)
{
    float acc[200];
#pragma unroll
    for (int i = 0; i < 200; ++i) acc[i] = out[(threadIdx.x + i) % n];
    for (int k = 0; k < n; ++k) {
#pragma unroll
        for (int i = 0; i < 200; ++i) acc[i] = fmaf(acc[i], acc[(i + 1) % 200], 1.0f);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 200; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    cudaFuncSetAttribute(big, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, big);
    printf("kernel regs=%d\n", fa.numRegs);
    for (int cs = 1; cs <= 16; ++cs) {
        cudaLaunchConfig_t cfg = {}; cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1; cfg.blockDim = dim3(256); cfg.gridDim = dim3(cs * 16);
        int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, big, &cfg);
        printf("cluster size %2d: max active clusters %3d (SMs used %3d) %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
