// Standalone probe: which 4-D TMA box shapes / coordinates work on this GPU (debugging aid).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap map, float* out, int n, int c0, int c1, int c2, int c3, int bytes)
{
    extern __shared__ __align__(1024) unsigned char raw[];
    __shared__ unsigned long long bar;
    float* st = (float*)raw;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                     :: "r"(smem_u32(st)), "l"(&map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@!p bra W;\n\t}" :: "r"(smem_u32(&bar)) : "memory");
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = st[i];
}
int main()
{
    void* f = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)f;
    const int W = 304, H = 228, C = 8, B = 2;
    std::vector<float> h((size_t)W * H * C * B);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 1000);
    float* d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    float* out; cudaMalloc(&out, 256 * 256 * 4);
    int shapes[][2] = {{64, 16}, {68, 82}, {72, 82}, {80, 82}, {128, 82}};
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (auto& s : shapes) {
        cuuint64_t dims[4] = {W, H, C, B};
        cuuint64_t strides[3] = {W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
        cuuint32_t box[4] = {(cuuint32_t)s[0], (cuuint32_t)s[1], 1, 1}, es[4] = {1, 1, 1, 1};
        alignas(64) CUtensorMap map; memset(&map, 0, sizeof map);
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        int coords[][2] = {{0, 0}, {-4, -1}, {236, 151}, {300, 220}};
        for (auto& c : coords) {
            int n = s[0] * s[1];
            k<<<1, 256, n * 4 + 1024>>>(map, out, n, c[0], c[1], 3, 1, n * 4);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<float> o(n); cudaMemcpy(o.data(), out, n * 4, cudaMemcpyDeviceToHost);
            // check a sample element
            int bad = 0;
            for (int y = 0; y < s[1] && e == cudaSuccess; ++y) for (int x = 0; x < s[0]; ++x) {
                int gx = c[0] + x, gy = c[1] + y;
                float exp = (gx < 0 || gx >= W || gy < 0 || gy >= H) ? 0.f : h[((size_t)(1 * C + 3) * H + gy) * W + gx];
                if (o[y * s[0] + x] != exp) ++bad;
            }
            printf("box %3dx%2d enc=%d coord (%4d,%4d): %s, mismatches %d\n", s[0], s[1], (int)r, c[0], c[1], cudaGetErrorString(e), bad);
            if (e != cudaSuccess) { cudaDeviceReset(); return 1; }
        }
    }
    return 0;
}
