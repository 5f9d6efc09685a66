// Ping-pong of 16-byte {data, tag} messages between CTAs on different SMs through L2 (st.relaxed.gpu / ld.relaxed.gpu),
// the transport of the streamed CSPN kernels: one-way latency = round trip / 2, alone and while every other SM runs
// `bg` warps that poll their own (never changing) lines back to back, as communication warps waiting for a late
// neighbour do.  Also the cost of a burst of 10 such stores from one lane (rim column) followed by the neighbour's poll.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o llping llping.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void st_ll(uint4* p, unsigned a, unsigned tag)
{
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %1, %2};" :: "l"(p), "r"(a), "r"(tag) : "memory");
}
__device__ __forceinline__ uint4 ld_ll(const uint4* p)
{
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

// pairs: CTA 2i pings CTA 2i+1.  box[cta] = one 128-byte line per CTA.  Background warps (threadIdx.x >= 32) poll line bgline[...].
__global__ void k_ping(uint4* box, uint4* bg, int iters, int msgs, int delay, long long* cyc, volatile int* stop)
{
    const int cta = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp > 0) {
        // background pollers: 4 lines per warp, 16 bytes per lane, never satisfied
        const uint4* p = bg + ((size_t)cta * 32 + warp) * 128 + lane;
        unsigned acc = 0;
        while (!*stop) {
            uint4 a = ld_ll(p), b = ld_ll(p + 32), c = ld_ll(p + 64), d = ld_ll(p + 96);
            acc += a.y + b.y + c.y + d.y;
            if (delay) { const long long t = clock64(); while (clock64() - t < delay) { } }
        }
        if (acc == 0x12345678u) bg[0].x = acc;
        return;
    }
    const int peer = cta ^ 1;
    uint4* mine = box + (size_t)cta * 64;      // 64 slots of 16 bytes
    uint4* theirs = box + (size_t)peer * 64;
    const bool first = (cta & 1) == 0;
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        if (first) {
            if (lane < msgs) st_ll(theirs + lane, (unsigned)i, (unsigned)i);
            if (lane < msgs) { uint4 v; do { v = ld_ll(mine + lane); } while (v.y != (unsigned)i || v.w != (unsigned)i); }
            __syncwarp();
        } else {
            if (lane < msgs) { uint4 v; do { v = ld_ll(mine + lane); } while (v.y != (unsigned)i || v.w != (unsigned)i); }
            __syncwarp();
            if (lane < msgs) st_ll(theirs + lane, (unsigned)i, (unsigned)i);
        }
    }
    long long t1 = clock64();
    if (lane == 0) cyc[cta] = t1 - t0;
}

int main()
{
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    uint4 *box, *bg; long long* cyc; int* stop;
    CK(cudaMalloc(&box, (size_t)256 * 64 * 16)); CK(cudaMalloc(&bg, (size_t)256 * 32 * 128 * 16)); CK(cudaMalloc(&cyc, 256 * 8));
    CK(cudaMallocManaged(&stop, 4));
    const int iters = 2000;
    for (int pairs : {1, 70}) for (int bgw : {0, 3}) for (int delay : {0, 300}) for (int msgs : {1, 32}) {
        if (bgw == 0 && delay) continue;
        CK(cudaMemset(box, 0, (size_t)256 * 64 * 16)); CK(cudaMemset(bg, 0, (size_t)256 * 32 * 128 * 16));
        *stop = 0;
        const int grid = 2 * pairs;
        // cooperative-style co-residency: grid <= SMs, one CTA per SM because of the big dynamic smem request
        CK(cudaFuncSetAttribute(k_ping, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        k_ping<<<grid, 32 * (1 + bgw), 200 * 1024>>>(box, bg, iters, msgs, delay, cyc, stop);
        // the pingers finish on their own; then release the background pollers
        cudaStream_t s2; CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
        std::vector<long long> h(grid);
        // wait until cyc of all pingers is written: poll with memcpy on another stream
        for (;;) {
            CK(cudaMemcpyAsync(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost, s2)); CK(cudaStreamSynchronize(s2));
            bool all = true; for (auto v : h) all = all && v != 0;
            if (all) break;
        }
        *stop = 1;
        CK(cudaDeviceSynchronize());
        double s = 0; for (auto v : h) s += (double)v; s /= grid;
        printf("pairs %3d  background polling warps/SM %d (delay %3d)  msgs %2d: one-way %7.1f cycles\n", pairs, bgw, delay, msgs, s / iters / 2);
        CK(cudaMemset(cyc, 0, 256 * 8));
        CK(cudaStreamDestroy(s2));
    }
    return 0;
}
