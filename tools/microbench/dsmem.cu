// DSMEM signalling primitives on B200: issue cost and one-way latency of
//  (a) st.async.b64 + complete_tx, (b) st.shared::cluster + mbarrier.arrive.release.cluster (remote),
//  (c) cp.async.bulk shared::cta -> shared::cluster + complete_tx.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
typedef unsigned long long u64;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t d; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(r)); return d; }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W_%=;\n\t}" :: "r"(bar), "r"(parity) : "memory");
}
__global__ void __cluster_dims__(2, 1, 1) k(long long* out, int mode, int nmsg, int lanes)
{
    __shared__ __align__(128) u64 box[2][1024];
    __shared__ __align__(128) u64 stagebuf[1024];
    __shared__ u64 bar[2];
    uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar[0])), "r"(mode == 1 ? lanes : 1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar[1])), "r"(mode == 1 ? lanes : 1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) stagebuf[i] = i;
    __syncthreads();
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    const uint32_t peer = mapa(smem_u32(&box[0][0]), rank ^ 1), pbar = mapa(smem_u32(&bar[0]), rank ^ 1);
    long long t_issue = 0, t_total = 0;
    const int ITER = 64;
    for (int it = 0; it < ITER; ++it) {
        const int par = it & 1;
        const uint32_t dst = peer + par * 8192, rb = pbar + par * 8;
        long long t0 = clock64();
        if (warp == 0) {
            if (mode == 0) {          // st.async per message
                if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar[par])), "r"(8 * nmsg * lanes) : "memory");
                if (lane < lanes)
                    for (int m = 0; m < nmsg; ++m)
                        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" :: "r"(dst + 8 * (m * lanes + lane)), "l"((u64)it), "r"(rb) : "memory");
            } else if (mode == 1) {   // plain DSMEM stores + one remote release-arrive per lane
                if (lane < lanes) {
                    for (int m = 0; m < nmsg; ++m)
                        asm volatile("st.shared::cluster.b64 [%0], %1;" :: "r"(dst + 8 * (m * lanes + lane)), "l"((u64)it) : "memory");
                    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(rb) : "memory");
                }
            } else {                  // one bulk copy smem -> peer smem
                if (threadIdx.x == 0) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar[par])), "r"(8 * nmsg * lanes) : "memory");
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 :: "r"(dst), "r"(smem_u32(stagebuf)), "r"(8 * nmsg * lanes), "r"(rb) : "memory");
                }
            }
        }
        long long t1 = clock64();
        mbar_wait(smem_u32(&bar[par]), (it >> 1) & 1);
        long long t2 = clock64();
        t_issue += t1 - t0; t_total += t2 - t0;
        __syncthreads();
    }
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x == 0) { out[blockIdx.x * 2] = t_issue / ITER; out[blockIdx.x * 2 + 1] = t_total / ITER; }
}
int main()
{
    long long* d; cudaMalloc(&d, 64 * sizeof(long long));
    const char* names[3] = {"st.async.b64", "st.shared::cluster + remote arrive.release", "cp.async.bulk s2s"};
    for (int mode = 0; mode < 3; ++mode)
        for (int lanes : {1, 2, 32})
            for (int nmsg : {2, 10, 20}) {
                if (mode == 2 && (8 * nmsg * lanes) % 16) continue;
                k<<<2, 256>>>(d, mode, nmsg, lanes);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[4]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
                printf("%-44s lanes=%2d msgs/lane=%2d (%4d B): issue %5lld cyc, issue+arrival %5lld cyc  %s\n", names[mode], lanes, nmsg, 8 * nmsg * lanes, h[0], h[1], e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    return 0;
}
