// Temporally blocked CSPN forward for the 5x5 "PAC variant" (mode OURS, K = 5, 24 guidance channels):
// CSPN_ours.py:24-54 over pac.conv2d (pac.py:75-94), S = 4 propagation steps per launch.
//
// A 5x5 tap set is 24 weights per pixel - too many for the cluster kernel's register tile - but the recurrence is
// short (T = 12 in BASELINE.json) and moves 2 pixels per step, so overlapped (trapezoid) tiles are cheap:
//  * one CTA owns a 64 x 32 pixel region; every thread owns 4 adjacent pixels of one row and keeps their
//    4 x 24 softmax weights (mask folded in) in REGISTERS for all S steps of the launch - the softmax of
//    CSPN_ours.py:35 is evaluated in registers straight from the raw guidance, no normalised-weight tensor
//    ever exists in HBM (the generic path writes and re-reads one every step);
//  * the depth region ping-pongs in shared memory (two 36 x 72 float planes), 15 LDS.128 per thread and step
//    feed 96 FMAs;
//  * after S steps the inner (64 - 4S) x (32 - 4S) pixels are exact and go back to HBM; T = 12 takes 3 launches.
// HBM traffic per launch: 24 + 3 planes read (overlaps between neighbouring regions are served by L2), one
// plane written; no inter-CTA synchronisation.
#include "cspn_common.cuh"

namespace cspn {

namespace {

constexpr int kRW = 64, kRH = 32;                 // region owned by one CTA
constexpr int kQuad = 4;                          // pixels per thread (adjacent in x)
constexpr int kThreads = (kRW / kQuad) * kRH;     // 512
constexpr int kPadX = 4, kPadY = 2;               // zero frame of the shared-memory planes (5x5 reach)
constexpr int kSW = kRW + 2 * kPadX;              // 72 floats per row: rows stay 16-byte aligned
constexpr int kSH = kRH + 2 * kPadY;
constexpr int kMaxS = 4;                          // steps per launch; halo = 2 * S <= 8
constexpr int kHaloX = 8;                         // x halo is fixed so that region origins stay 16-byte aligned
constexpr int kOutW = kRW - 2 * kHaloX;           // 48

__device__ __forceinline__ float fast_exp2(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <typename T> struct Vec4;
template <> struct Vec4<float> { typedef float4 type; };
template <> struct Vec4<__half> { typedef uint2 type; };

__device__ __forceinline__ void unpack4(float4 v, float* o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
__device__ __forceinline__ void unpack4(uint2 v, float* o)
{
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}

// 4 adjacent pixels of a plane row starting at (gy, gx); zero outside the image.  vec: base pointer 16-byte
// aligned (8 for half) and W % 4 == 0, so that an in-range quad never straddles the image edge.
template <typename TP>
__device__ __forceinline__ void load_quad(const TP* plane, int gy, int gx, int H, int W, bool vec, float* o)
{
    o[0] = o[1] = o[2] = o[3] = 0.f;
    if (gy < 0 || gy >= H) return;
    const TP* row = plane + (size_t)gy * W;
    if (vec) {
        if (gx >= 0 && gx + 3 < W) unpack4(*reinterpret_cast<const typename Vec4<TP>::type*>(row + gx), o);
        return;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (gx + q >= 0 && gx + q < W) o[q] = to_f32(row[gx + q]);
}

template <typename T, typename TIn, typename TOut>
__global__ void __launch_bounds__(kThreads, 1)
blocked5x5_kernel(const T* __restrict__ g, int64_t gbs, const TIn* __restrict__ rin, const T* __restrict__ d0,
                  const T* __restrict__ sparse, int sparse_channels, TOut* __restrict__ rout,
                  int C, int H, int W, int steps, int vec)
{
    __shared__ __align__(16) float buf[2][kSH][kSW];
    const int tid = threadIdx.x;
    const int qx = tid % (kRW / kQuad), ry = tid / (kRW / kQuad);
    const int halo_y = 2 * steps, out_h = kRH - 2 * halo_y;
    const int plane = blockIdx.z, b = plane / C, ch = plane - b * C;
    const int ox = blockIdx.x * kOutW - kHaloX, oy = blockIdx.y * out_h - halo_y;      // region origin in the image
    const int gx = ox + qx * kQuad, gy = oy + ry;
    const size_t hw = (size_t)H * W;

    // zero both planes (frame included): what lies outside the image reads as the reference's zero padding
    for (int i = tid; i < 2 * kSH * kSW / 4; i += kThreads) reinterpret_cast<float4*>(&buf[0][0][0])[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    // ---- weights: softmax over the 24 channels (CSPN_ours.py:35), tap j = channel j, row-major 5x5 without the centre
    float w[kQuad][24];
    {
        const T* gb = g + (size_t)b * gbs;
#pragma unroll
        for (int j = 0; j < 24; ++j) {
            float v[4];
            load_quad(gb + (size_t)j * hw, gy, gx, H, W, vec != 0, v);
#pragma unroll
            for (int q = 0; q < kQuad; ++q) w[q][j] = v[q];
        }
    }
    float dq[4], mq[4], rq[4];
    load_quad(d0 + (size_t)plane * hw, gy, gx, H, W, vec != 0, dq);
    load_quad(rin + (size_t)plane * hw, gy, gx, H, W, vec != 0, rq);
    mq[0] = mq[1] = mq[2] = mq[3] = 0.f;
    if (sparse) {
        load_quad(sparse + ((size_t)b * sparse_channels + (sparse_channels == 1 ? 0 : ch)) * hw, gy, gx, H, W, vec != 0, mq);
#pragma unroll
        for (int q = 0; q < kQuad; ++q) mq[q] = signf(mq[q]);
    }
    float cq[kQuad];
#pragma unroll
    for (int q = 0; q < kQuad; ++q) {
        const bool in = gy >= 0 && gy < H && gx + q >= 0 && gx + q < W;
        float mx = w[q][0];
#pragma unroll
        for (int j = 1; j < 24; ++j) mx = fmaxf(mx, w[q][j]);
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 24; ++j) { w[q][j] = fast_exp2((w[q][j] - mx) * 1.4426950408889634f); s += w[q][j]; }   // MUFU.EX2: 2 ulp, the kernel is MUFU-bound here
        // fold the re-injection (1-m) * prop + m * d0 (CSPN_ours.py:51-53); pixels outside the image stay exactly zero
        const float f = in ? (1.f - mq[q]) / s : 0.f;
#pragma unroll
        for (int j = 0; j < 24; ++j) w[q][j] *= f;
        cq[q] = in ? mq[q] * dq[q] : 0.f;
    }
    __syncthreads();                                                                    // zero fill done
    *reinterpret_cast<float4*>(&buf[0][ry + kPadY][kPadX + qx * kQuad]) = make_float4(rq[0], rq[1], rq[2], rq[3]);
    __syncthreads();

    // ---- S steps in shared memory; after step s everything further than 2s pixels from the region edge is exact
    float acc[kQuad];
    for (int s = 0; s < steps; ++s) {
        const float(*cur)[kSW] = buf[s & 1];
#pragma unroll
        for (int q = 0; q < kQuad; ++q) acc[q] = cq[q];
        // rows that can no longer reach the exact inner block are not computed (a warp owns two whole rows, so the
        // test is warp-uniform): the last step only needs the inner rows, the one before 2 more on each side, ...
        const int live = halo_y - 2 * (steps - 1 - s);
        if (ry >= live && ry < kRH - live) {
#pragma unroll
        for (int dy = -2; dy <= 2; ++dy) {
            const float* row = &cur[ry + kPadY + dy][kPadX + qx * kQuad];
            float v[12];
            unpack4(*reinterpret_cast<const float4*>(row - 4), v);
            unpack4(*reinterpret_cast<const float4*>(row), v + 4);
            unpack4(*reinterpret_cast<const float4*>(row + 4), v + 8);
#pragma unroll
            for (int dx = -2; dx <= 2; ++dx) {
                if (dy == 0 && dx == 0) continue;
                const int jj = (dy + 2) * 5 + (dx + 2);
                const int j = jj < 12 ? jj : jj - 1;
#pragma unroll
                for (int q = 0; q < kQuad; ++q) acc[q] = fmaf(w[q][j], v[4 + q + dx], acc[q]);
            }
        }
        }
        *reinterpret_cast<float4*>(&buf[(s + 1) & 1][ry + kPadY][kPadX + qx * kQuad]) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        __syncthreads();
    }

    // ---- the exact inner block goes back to HBM
    const int lx = qx * kQuad;
    if (ry >= halo_y && ry < kRH - halo_y && lx >= kHaloX && lx < kRW - kHaloX && gy < H) {
        TOut* orow = rout + (size_t)plane * hw + (size_t)gy * W;
#pragma unroll
        for (int q = 0; q < kQuad; ++q)
            if (gx + q < W) orow[gx + q] = from_f32<TOut>(acc[q]);
    }
}

template <typename T, typename TIn, typename TOut>
int launch_one(const FwdArgs<T>& a, const TIn* rin, TOut* rout, int steps, bool vec)
{
    const int out_h = kRH - 4 * steps;
    dim3 grid((unsigned)((a.W + kOutW - 1) / kOutW), (unsigned)((a.H + out_h - 1) / out_h), (unsigned)(a.B * a.C));
    blocked5x5_kernel<T, TIn, TOut><<<grid, kThreads, 0, a.stream>>>(a.guidance, a.gbs, rin, a.depth, a.sparse, a.sparse_channels, rout,
                                                                      a.C, a.H, a.W, steps, vec ? 1 : 0);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    ++call_stats().launches;
    return 0;
}

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

bool blocked5x5_supported(int B, int C, int H, int W, int iters, int ksize, int mode)
{
    return mode == CSPN_MODE_OURS && ksize == 5 && iters >= 1 && (long)B * C <= 65535 && (long)H * W <= (1l << 30) && (H + 15) / 16 <= 65535;
}

// two fp32 depth planes for the hand-over between launches (only when more than one launch is needed)
size_t blocked5x5_workspace(int B, int C, int H, int W, int iters)
{
    if (iters <= kMaxS) return 0;
    return (iters <= 2 * kMaxS ? 1 : 2) * align256((size_t)B * C * H * W * sizeof(float));
}

template <typename T>
int blocked5x5_forward(const FwdArgs<T>& a)
{
    const size_t npx = (size_t)a.B * a.C * a.H * a.W;
    const size_t need = blocked5x5_workspace(a.B, a.C, a.H, a.W, a.iters);
    if (need && (!a.ws || a.ws_bytes < need)) return CSPN_ERR_WORKSPACE;
    float* r0 = (float*)a.ws;
    float* r1 = (float*)((char*)a.ws + align256(npx * sizeof(float)));
    const int vsz = 4 * (int)sizeof(T);
    // vector path: every plane row starts 4-pixel aligned (the fp32 hand-over planes are 256-byte aligned)
    const bool vec = (a.W % 4 == 0) && ((uintptr_t)a.guidance % vsz == 0) && ((uintptr_t)a.depth % vsz == 0) && (a.gbs % 4 == 0) &&
                     (!a.sparse || (uintptr_t)a.sparse % vsz == 0) && ((uintptr_t)a.out % vsz == 0);
    int done = 0, rc = 0;
    const float* cur = nullptr;
    while (done < a.iters && rc == 0) {
        const int steps = a.iters - done < kMaxS ? a.iters - done : kMaxS;
        const bool first = done == 0, last = done + steps == a.iters;
        float* nxt = cur == r0 ? r1 : r0;
        if (first && last) rc = launch_one<T, T, T>(a, a.depth, a.out, steps, vec);
        else if (first) rc = launch_one<T, T, float>(a, a.depth, nxt, steps, vec);
        else if (last) rc = launch_one<T, float, T>(a, cur, a.out, steps, vec);
        else rc = launch_one<T, float, float>(a, cur, nxt, steps, vec);
        cur = nxt;
        done += steps;
    }
    return rc;
}

template int blocked5x5_forward<float>(const FwdArgs<float>&);
template int blocked5x5_forward<__half>(const FwdArgs<__half>&);

}  // namespace cspn
