// Microbenchmarks that ground the CSPN kernel design on B200 (sm_100a).
// Measures per-SM instruction throughputs (FFMA, FFMA2, SHFL, LDS), a mixed
// stencil-like instruction blend, barrier latencies and cluster launch limits.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb mb.cu
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace cg = cooperative_groups;
typedef unsigned long long u64;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
    u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ u64 pack(float lo, float hi) {
    u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d;
}
__device__ __forceinline__ void unpack(u64 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

constexpr int ITERS = 2048;

// 16 independent FFMA chains per thread.
__global__ void k_ffma(float* out, float a, float b, long long* cyc) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// 16 independent FFMA chains with 3 distinct register sources (weights in regs).
__global__ void k_ffma3(float* out, const float* w, long long* cyc) {
    float acc[16], wr[16], xr[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { acc[i] = threadIdx.x + i; wr[i] = w[i]; xr[i] = w[16 + i]; }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(wr[i], xr[(i + 1) & 15], acc[i]);
#pragma unroll
        for (int i = 0; i < 16; ++i) xr[i] = fmaf(wr[(i + 3) & 15], acc[(i + 5) & 15], xr[i]);
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 16; ++i) s += acc[i] + xr[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_ffma2(float* out, const float* w, long long* cyc) {
    u64 acc[16], wr[16], xr[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { acc[i] = pack(threadIdx.x + i, i); wr[i] = pack(w[i], w[i + 1]); xr[i] = pack(w[16 + i], w[17 + i]); }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = ffma2(wr[i], xr[(i + 1) & 15], acc[i]);
#pragma unroll
        for (int i = 0; i < 16; ++i) xr[i] = ffma2(wr[(i + 3) & 15], acc[(i + 5) & 15], xr[i]);
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 16; ++i) { float lo, hi; unpack(acc[i], lo, hi); s += lo + hi; unpack(xr[i], lo, hi); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_shfl(float* out, long long* cyc) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 8 + i;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __shfl_up_sync(0xffffffffu, v[i], 1);
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int VEC>
__global__ void k_lds(float* out, long long* cyc) {
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
    __syncthreads();
    float s = 0;
    int base = (threadIdx.x * VEC) & 4095;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int idx = (base + i * 32 * VEC + (it & 1) * 4) & 8191;
            if (VEC == 1) { s += *(volatile float*)&sm[idx]; }
            else if (VEC == 2) { float2 t; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(t.x), "=f"(t.y) : "r"((unsigned)__cvta_generic_to_shared(&sm[idx & ~1]))); s += t.x + t.y; }
            else { float4 t; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"((unsigned)__cvta_generic_to_shared(&sm[idx & ~3]))); s += t.x + t.y + t.z + t.w; }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// Stencil-like blend per "row": 8 FFMA2 + 2 SHFL + 2 MOV-ish packs, 8 rows per iter.
__global__ void k_mix(float* out, const float* w, long long* cyc) {
    u64 wr[8][8]; u64 r[8];
#pragma unroll
    for (int y = 0; y < 8; ++y) { r[y] = pack(threadIdx.x + y, y);
#pragma unroll
        for (int k = 0; k < 8; ++k) wr[y][k] = pack(w[y * 8 + k], w[y * 8 + k + 1]); }
    long long t0 = clock64();
    for (int it = 0; it < ITERS / 8; ++it) {
        u64 s1[8], s2[8];
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            float lo, hi; unpack(r[y], lo, hi);
            float lh = __shfl_up_sync(0xffffffffu, hi, 1);
            float rl = __shfl_down_sync(0xffffffffu, lo, 1);
            s1[y] = pack(lh, lo); s2[y] = pack(hi, rl);
        }
        u64 nr[8];
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            int ym = (y + 7) & 7, yp = (y + 1) & 7;
            u64 acc = ffma2(wr[y][0], s1[ym], r[y]);
            acc = ffma2(wr[y][1], r[ym], acc);
            acc = ffma2(wr[y][2], s2[ym], acc);
            acc = ffma2(wr[y][3], s1[y], acc);
            acc = ffma2(wr[y][4], s2[y], acc);
            acc = ffma2(wr[y][5], s1[yp], acc);
            acc = ffma2(wr[y][6], r[yp], acc);
            acc = ffma2(wr[y][7], s2[yp], acc);
            nr[y] = acc;
        }
#pragma unroll
        for (int y = 0; y < 8; ++y) r[y] = nr[y];
    }
    long long t1 = clock64();
    float s = 0; for (int y = 0; y < 8; ++y) { float lo, hi; unpack(r[y], lo, hi); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// Same blend with scalar FFMA (no f32x2) for comparison.
__global__ void k_mix_scalar(float* out, const float* w, long long* cyc) {
    float wr[8][8][2]; float r[8][2];
#pragma unroll
    for (int y = 0; y < 8; ++y) { r[y][0] = threadIdx.x + y; r[y][1] = y;
#pragma unroll
        for (int k = 0; k < 8; ++k) { wr[y][k][0] = w[y * 8 + k]; wr[y][k][1] = w[y * 8 + k + 1]; } }
    long long t0 = clock64();
    for (int it = 0; it < ITERS / 8; ++it) {
        float lh[8], rl[8];
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            lh[y] = __shfl_up_sync(0xffffffffu, r[y][1], 1);
            rl[y] = __shfl_down_sync(0xffffffffu, r[y][0], 1);
        }
        float nr[8][2];
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            int ym = (y + 7) & 7, yp = (y + 1) & 7;
            float a0 = r[y][0], a1 = r[y][1];
            a0 = fmaf(wr[y][0][0], lh[ym], a0);   a1 = fmaf(wr[y][0][1], r[ym][0], a1);
            a0 = fmaf(wr[y][1][0], r[ym][0], a0); a1 = fmaf(wr[y][1][1], r[ym][1], a1);
            a0 = fmaf(wr[y][2][0], r[ym][1], a0); a1 = fmaf(wr[y][2][1], rl[ym], a1);
            a0 = fmaf(wr[y][3][0], lh[y], a0);    a1 = fmaf(wr[y][3][1], r[y][0], a1);
            a0 = fmaf(wr[y][4][0], r[y][1], a0);  a1 = fmaf(wr[y][4][1], rl[y], a1);
            a0 = fmaf(wr[y][5][0], lh[yp], a0);   a1 = fmaf(wr[y][5][1], r[yp][0], a1);
            a0 = fmaf(wr[y][6][0], r[yp][0], a0); a1 = fmaf(wr[y][6][1], r[yp][1], a1);
            a0 = fmaf(wr[y][7][0], r[yp][1], a0); a1 = fmaf(wr[y][7][1], rl[yp], a1);
            nr[y][0] = a0; nr[y][1] = a1;
        }
#pragma unroll
        for (int y = 0; y < 8; ++y) { r[y][0] = nr[y][0]; r[y][1] = nr[y][1]; }
    }
    long long t1 = clock64();
    float s = 0; for (int y = 0; y < 8; ++y) s += r[y][0] + r[y][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_syncthreads(long long* cyc) {
    long long t0 = clock64();
    for (int it = 0; it < 1024; ++it) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_cluster_sync(long long* cyc) {
    cg::cluster_group cl = cg::this_cluster();
    long long t0 = clock64();
    for (int it = 0; it < 256; ++it) cl.sync();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// DSMEM neighbour push + cluster barrier (split arrive/wait), per "iteration".
__global__ void k_cluster_push(long long* cyc, float* out) {
    __shared__ float buf[2][512];
    cg::cluster_group cl = cg::this_cluster();
    unsigned rank = cl.block_rank(), n = cl.num_blocks();
    float* remote = cl.map_shared_rank(&buf[0][0], (rank + 1) % n);
    float v = threadIdx.x;
    cl.sync();
    long long t0 = clock64();
    for (int it = 0; it < 256; ++it) {
        if (threadIdx.x < 128) remote[(it & 1) * 512 + threadIdx.x] = v;
        asm volatile("barrier.cluster.arrive.release.aligned;");
        asm volatile("barrier.cluster.wait.acquire.aligned;");
        v += buf[it & 1][threadIdx.x & 127];
    }
    long long t1 = clock64();
    cl.sync();
    out[blockIdx.x * blockDim.x + threadIdx.x] = v;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_empty() {}

static double avg(const std::vector<long long>& v) { double s = 0; for (auto x : v) s += x; return s / v.size(); }

template <typename F>
static void run_tp(const char* name, F launch, int grid, int block, double ops_per_thread, long long* dcyc) {
    launch(); CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> h(grid); CK(cudaMemcpy(h.data(), dcyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    double c = avg(h);
    double ops_per_sm_clk = ops_per_thread * block * (grid / 148.0) / c;   // grid is a multiple of 148 → CTAs/SM
    printf("%-28s grid=%d block=%d cyc=%.0f ms=%.4f  lane-ops/clk/SM=%.2f (warp-instr/clk/SM=%.3f)\n", name, grid, block, c, ms, ops_per_sm_clk, ops_per_sm_clk / 32);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sm_%d%d SMs=%d smem/SM=%zu regs/SM=%d clock=%d kHz\n", p.name, p.major, p.minor, p.multiProcessorCount, p.sharedMemPerMultiprocessor, p.regsPerMultiprocessor, p.clockRate);
    float* out; long long* cyc; float* w;
    CK(cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float))); CK(cudaMalloc(&cyc, 148 * 16 * sizeof(long long))); CK(cudaMalloc(&w, 1024 * sizeof(float)));
    std::vector<float> hw(1024); for (int i = 0; i < 1024; ++i) hw[i] = 0.001f * (i % 7) + 0.1f;
    CK(cudaMemcpy(w, hw.data(), 1024 * sizeof(float), cudaMemcpyHostToDevice));

    for (int block : {128, 256, 512, 1024}) {
        run_tp("FFMA(imm-ish 2 src)", [&] { k_ffma<<<148, block>>>(out, 1.0001f, 0.5f, cyc); }, 148, block, 16.0 * ITERS, cyc);
        run_tp("FFMA 3-reg", [&] { k_ffma3<<<148, block>>>(out, w, cyc); }, 148, block, 32.0 * ITERS, cyc);
        run_tp("FFMA2 (pairs; x2 flop)", [&] { k_ffma2<<<148, block>>>(out, w, cyc); }, 148, block, 32.0 * ITERS, cyc);
    }
    for (int block : {128, 256, 512}) {
        run_tp("SHFL.UP", [&] { k_shfl<<<148, block>>>(out, cyc); }, 148, block, 8.0 * ITERS, cyc);
        run_tp("LDS.32", [&] { k_lds<1><<<148, block, 32768>>>(out, cyc); }, 148, block, 8.0 * ITERS, cyc);
        run_tp("LDS.64", [&] { k_lds<2><<<148, block, 32768>>>(out, cyc); }, 148, block, 8.0 * ITERS, cyc);
        run_tp("LDS.128", [&] { k_lds<4><<<148, block, 32768>>>(out, cyc); }, 148, block, 8.0 * ITERS, cyc);
    }
    // mix: per thread per iter-of-8-rows: 64 FFMA2 (=128 FMA lanes-ops) ; report in FMA lane-ops
    for (int block : {128, 256}) {
        run_tp("MIX f32x2 (FMA lane-ops)", [&] { k_mix<<<148, block>>>(out, w, cyc); }, 148, block, 128.0 * (ITERS / 8), cyc);
        run_tp("MIX scalar (FMA lane-ops)", [&] { k_mix_scalar<<<148, block>>>(out, w, cyc); }, 148, block, 128.0 * (ITERS / 8), cyc);
    }
    // barriers
    for (int block : {128, 256, 512}) {
        k_syncthreads<<<148, block>>>(cyc); CK(cudaDeviceSynchronize());
        std::vector<long long> h(148); CK(cudaMemcpy(h.data(), cyc, 148 * 8, cudaMemcpyDeviceToHost));
        printf("__syncthreads block=%d: %.1f cyc each\n", block, avg(h) / 1024);
    }
    for (int cs : {2, 4, 8, 15, 16}) {
        cudaLaunchConfig_t cfg = {}; cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1; cfg.blockDim = dim3(256); cfg.gridDim = dim3(cs * (128 / cs)); cfg.dynamicSmemBytes = 0;
        cudaError_t e = cudaFuncSetAttribute(k_cluster_sync, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        e = cudaFuncSetAttribute(k_cluster_push, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        int ncl = -1; cudaError_t eo = cudaOccupancyMaxActiveClusters(&ncl, k_cluster_sync, &cfg);
        printf("cluster size %d: maxActiveClusters=%d (%s)\n", cs, ncl, cudaGetErrorString(eo));
        e = cudaLaunchKernelEx(&cfg, k_cluster_sync, cyc);
        cudaError_t e2 = cudaDeviceSynchronize();
        if (e != cudaSuccess || e2 != cudaSuccess) { printf("  cluster_sync launch failed: %s / %s\n", cudaGetErrorString(e), cudaGetErrorString(e2)); cudaGetLastError(); continue; }
        std::vector<long long> h(cfg.gridDim.x); CK(cudaMemcpy(h.data(), cyc, h.size() * 8, cudaMemcpyDeviceToHost));
        printf("  cluster.sync: %.1f cyc each\n", avg(h) / 256);
        e = cudaLaunchKernelEx(&cfg, k_cluster_push, cyc, out);
        e2 = cudaDeviceSynchronize();
        if (e != cudaSuccess || e2 != cudaSuccess) { printf("  cluster_push launch failed: %s / %s\n", cudaGetErrorString(e), cudaGetErrorString(e2)); cudaGetLastError(); continue; }
        CK(cudaMemcpy(h.data(), cyc, h.size() * 8, cudaMemcpyDeviceToHost));
        printf("  DSMEM push + barrier + read: %.1f cyc per iteration\n", avg(h) / 256);
    }
    // 2D cluster shape 5x3 = 15
    {
        cudaLaunchConfig_t cfg = {}; cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 5; at[0].val.clusterDim.y = 3; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1; cfg.blockDim = dim3(256); cfg.gridDim = dim3(5, 3, 8);
        int ncl = -1; cudaError_t eo = cudaOccupancyMaxActiveClusters(&ncl, k_cluster_sync, &cfg);
        cudaError_t e = cudaLaunchKernelEx(&cfg, k_cluster_sync, cyc); cudaError_t e2 = cudaDeviceSynchronize();
        printf("cluster 5x3x1 grid 5x3x8: maxActive=%d (%s) launch=%s sync=%s\n", ncl, cudaGetErrorString(eo), cudaGetErrorString(e), cudaGetErrorString(e2)); cudaGetLastError();
    }
    // launch overhead
    {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int i = 0; i < 10; ++i) k_empty<<<148, 256>>>();
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0)); for (int i = 0; i < 1000; ++i) k_empty<<<148, 256>>>(); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); printf("empty kernel back-to-back: %.2f us each\n", ms);
    }
    return 0;
}
