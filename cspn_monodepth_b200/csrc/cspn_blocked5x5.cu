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

// REV = false: the forward recurrence.  REV = true: the adjoint recurrence of the backward pass,
//   G^t(q) = sum_j k'_j(q - o_j) G^{t+1}(q - o_j),  k' = (1 - m) softmax  (pac.py:96-121 through the loop of CSPN_ours.py:47-53),
// which is the SAME 24-tap gather once every thread holds the transposed weights wt_j(q) = k'_j(q - o_j): they are fetched from
// the neighbours through shared memory once per launch (12 rounds of 2 tap planes), then the S steps run exactly like the forward.
// hist != nullptr: the exact inner block of every step's result (the first hist_count steps) also goes to the fp32 plane set
// hist + s * hist_step - the backward pass needs every r^t and every G^t.  rout may then be nullptr.
template <typename T, typename TIn, typename TOut, bool REV>
__global__ void __launch_bounds__(kThreads, 1)
blocked5x5_kernel(const T* __restrict__ g, int64_t gbs, const TIn* __restrict__ rin, const T* __restrict__ d0,
                  const T* __restrict__ sparse, int sparse_channels, TOut* __restrict__ rout,
                  float* __restrict__ hist, long long hist_step, int hist_count,
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
    dq[0] = dq[1] = dq[2] = dq[3] = 0.f;
    if (!REV) load_quad(d0 + (size_t)plane * hw, gy, gx, H, W, vec != 0, dq);
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
        cq[q] = (in && !REV) ? mq[q] * dq[q] : 0.f;
    }
    __syncthreads();                                                                    // zero fill done
    if (REV) {
        // transposed weights: tap j of pixel q is the weight pixel q - o_j applies to ITS tap j.  Pixels whose source lies
        // outside the region read the zero frame - they are within 2 pixels of the region edge, never part of an exact result.
#pragma unroll
        for (int r = 0; r < 12; ++r) {
            // taps r and 23 - r have opposite offsets (o_{23-r} = -o_r): in gather form the tap with offset o_r must carry
            // k'_{23-r}(q + o_r), the tap with offset -o_r carries k'_r(q - o_r)
            *reinterpret_cast<float4*>(&buf[0][ry + kPadY][kPadX + qx * kQuad]) = make_float4(w[0][r], w[1][r], w[2][r], w[3][r]);
            *reinterpret_cast<float4*>(&buf[1][ry + kPadY][kPadX + qx * kQuad]) = make_float4(w[0][23 - r], w[1][23 - r], w[2][23 - r], w[3][23 - r]);
            __syncthreads();
            const int dy = r / 5 - 2, dx = r % 5 - 2;
#pragma unroll
            for (int q = 0; q < kQuad; ++q) {
                w[q][23 - r] = buf[0][ry + kPadY - dy][kPadX + qx * kQuad + q - dx];
                w[q][r] = buf[1][ry + kPadY + dy][kPadX + qx * kQuad + q + dx];
            }
            __syncthreads();
        }
    }
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
        if (hist != nullptr && s < hist_count && ry >= halo_y && ry < kRH - halo_y && qx * kQuad >= kHaloX && qx * kQuad < kRW - kHaloX && gy < H) {
            float* hrow = hist + (long long)s * hist_step + (size_t)plane * hw + (size_t)gy * W;
            if (vec) {
                if (gx + 3 < W) *reinterpret_cast<float4*>(hrow + gx) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            } else {
#pragma unroll
                for (int q = 0; q < kQuad; ++q)
                    if (gx + q < W) hrow[gx + q] = acc[q];
            }
        }
        __syncthreads();
    }

    // ---- the exact inner block goes back to HBM
    const int lx = qx * kQuad;
    if (rout != nullptr && ry >= halo_y && ry < kRH - halo_y && lx >= kHaloX && lx < kRW - kHaloX && gy < H) {
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
    blocked5x5_kernel<T, TIn, TOut, false><<<grid, kThreads, 0, a.stream>>>(a.guidance, a.gbs, rin, a.depth, a.sparse, a.sparse_channels, rout,
                                                                             nullptr, 0, 0, a.C, a.H, a.W, steps, vec ? 1 : 0);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    ++call_stats().launches;
    return 0;
}

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

__device__ __forceinline__ void store_quad(float* row, int gx, int W, bool vec, const float* v)
{
    if (vec) { if (gx + 3 < W) *reinterpret_cast<float4*>(row + gx) = make_float4(v[0], v[1], v[2], v[3]); return; }
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (gx + q < W) row[gx + q] = v[q];
}
__device__ __forceinline__ void store_quad(__half* row, int gx, int W, bool vec, const float* v)
{
    if (vec) {
        if (gx + 3 < W) {
            uint2 u;
            *reinterpret_cast<__half2*>(&u.x) = __floats2half2_rn(v[0], v[1]);
            *reinterpret_cast<__half2*>(&u.y) = __floats2half2_rn(v[2], v[3]);
            *reinterpret_cast<uint2*>(row + gx) = u;
        }
        return;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (gx + q < W) row[gx + q] = __float2half_rn(v[q]);
}

// stage the (kRH + 4) x (kRW + 8) window of a plane around a region (origin (oy, ox) = buf row kPadY, column kPadX) into shared memory
template <typename TP>
__device__ __forceinline__ void stage_window(float (*dst)[kSW], const TP* plane, int oy, int ox, int H, int W, bool vec)
{
    for (int i = threadIdx.x; i < kSH * (kSW / 4); i += kThreads) {
        const int row = i / (kSW / 4), qc = i - row * (kSW / 4);
        float v[4];
        load_quad(plane, oy - kPadY + row, ox - kPadX + 4 * qc, H, W, vec, v);
        *reinterpret_cast<float4*>(&dst[row][4 * qc]) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// Jacobians of the 5x5 backward.  The recurrences have left every r^t (rhist: r^1 .. r^{T-1}; r^0 = depth) and every G^t
// (ghist: G^0 .. G^{T-1}; G^T = grad_out) in HBM; what remains has no recurrence and is one pass:
//   dL/dk_j(p)   = sum_ch sum_t (1 - m) G^{t+1}(p) r^t(p + o_j)          24 accumulators per pixel in registers,
//                                                                         r^t windows double-buffered in shared memory
//   dL/dguided_c = s_c (dL/dk_c - sum_c' s_c' dL/dk_c')                   softmax Jacobian (CSPN_ours.py:35), s recomputed
//   dL/dx        = m * sum_{t>=1} G^t + G^0                               (CSPN_ours.py:53: the re-injected input)
template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
jacobian5x5_kernel(const T* __restrict__ g, int64_t gbs, int Cg, const T* __restrict__ depth, const T* __restrict__ grad_out,
                   const T* __restrict__ sparse, int sparse_channels, const float* __restrict__ rhist, const float* __restrict__ ghist,
                   size_t npx, T* __restrict__ grad_guidance, T* __restrict__ grad_depth, int C, int H, int W, int iters, int vec)
{
    __shared__ __align__(16) float buf[2][kSH][kSW];
    const int tid = threadIdx.x;
    const int qx = tid % (kRW / kQuad), ry = tid / (kRW / kQuad);
    const int b = blockIdx.z;
    const int ox = blockIdx.x * kRW, oy = blockIdx.y * kRH;
    const int gx = ox + qx * kQuad, gy = oy + ry;
    const size_t hw = (size_t)H * W;
    const bool vz = vec != 0;

    float acc[kQuad][24];
#pragma unroll
    for (int q = 0; q < kQuad; ++q)
#pragma unroll
        for (int j = 0; j < 24; ++j) acc[q][j] = 0.f;

    int it = 0;
    for (int ch = 0; ch < C; ++ch) {
        const size_t plane = (size_t)b * C + ch;
        float mq[4] = {0.f, 0.f, 0.f, 0.f}, gsum[4] = {0.f, 0.f, 0.f, 0.f};
        if (sparse) {
            load_quad(sparse + ((size_t)b * sparse_channels + (sparse_channels == 1 ? 0 : ch)) * hw, gy, gx, H, W, vz, mq);
#pragma unroll
            for (int q = 0; q < kQuad; ++q) mq[q] = signf(mq[q]);
        }
        for (int t = 0; t < iters; ++t, ++it) {
            float (*cur)[kSW] = buf[it & 1];
            if (t == 0) stage_window(cur, depth + plane * hw, oy, ox, H, W, vz);
            else stage_window(cur, rhist + (size_t)(t - 1) * npx + plane * hw, oy, ox, H, W, vz);
            float G[4];
            if (t + 1 == iters) load_quad(grad_out + plane * hw, gy, gx, H, W, vz, G);
            else load_quad(ghist + (size_t)(t + 1) * npx + plane * hw, gy, gx, H, W, vz, G);
            float u[4];
#pragma unroll
            for (int q = 0; q < kQuad; ++q) { u[q] = (1.f - mq[q]) * G[q]; gsum[q] = fmaf(mq[q], G[q], gsum[q]); }
            __syncthreads();
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
                    for (int q = 0; q < kQuad; ++q) acc[q][j] = fmaf(u[q], v[4 + q + dx], acc[q][j]);
                }
            }
        }
        if (gy < H) {
            float G0[4];
            load_quad(ghist + plane * hw, gy, gx, H, W, vz, G0);
#pragma unroll
            for (int q = 0; q < kQuad; ++q) G0[q] += gsum[q];
            store_quad(grad_depth + plane * hw + (size_t)gy * W, gx, W, vz, G0);
        }
    }
    if (gy >= H) return;
    // softmax Jacobian, one pixel at a time (the 96 accumulators leave no room for 4 x 24 probabilities); in place
    const T* gb = g + (size_t)b * gbs + (size_t)gy * W;
#pragma unroll
    for (int q = 0; q < kQuad; ++q) {
        if (gx + q >= W) continue;
        float mx = -INFINITY;
        for (int j = 0; j < 24; ++j) mx = fmaxf(mx, to_f32(gb[(size_t)j * hw + gx + q]));
        float ssum = 0.f, dot = 0.f;
#pragma unroll
        for (int j = 0; j < 24; ++j) {
            const float e = fast_exp2((to_f32(gb[(size_t)j * hw + gx + q]) - mx) * 1.4426950408889634f);
            ssum += e;
            dot = fmaf(e, acc[q][j], dot);
        }
        const float inv = 1.f / ssum;
        dot *= inv;
#pragma unroll
        for (int j = 0; j < 24; ++j) {
            const float e = fast_exp2((to_f32(gb[(size_t)j * hw + gx + q]) - mx) * 1.4426950408889634f);
            acc[q][j] = e * inv * (acc[q][j] - dot);
        }
    }
    T* ggb = grad_guidance + (size_t)b * Cg * hw + (size_t)gy * W;
#pragma unroll
    for (int j = 0; j < 24; ++j) {
        const float v[4] = {acc[0][j], acc[1][j], acc[2][j], acc[3][j]};
        store_quad(ggb + (size_t)j * hw, gx, W, vz, v);
    }
    const float z[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 24; j < Cg; ++j) store_quad(ggb + (size_t)j * hw, gx, W, vz, z);      // channels the forward never reads
}

template <typename T, typename TIn, bool REV>
int launch_hist(const BwdArgs<T>& a, const TIn* rin, float* hist, long long hist_step, int steps, bool vec)
{
    const int out_h = kRH - 4 * steps;
    dim3 grid((unsigned)((a.W + kOutW - 1) / kOutW), (unsigned)((a.H + out_h - 1) / out_h), (unsigned)(a.B * a.C));
    blocked5x5_kernel<T, TIn, float, REV><<<grid, kThreads, 0, a.stream>>>(a.guidance, a.gbs, rin, a.depth, a.sparse, a.sparse_channels, (float*)nullptr,
                                                                            hist, hist_step, steps, a.C, a.H, a.W, steps, vec ? 1 : 0);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    ++call_stats().launches;
    return 0;
}

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

bool blocked5x5_bwd_supported(int B, int C, int H, int W, int iters, int ksize, int mode)
{
    return blocked5x5_supported(B, C, H, W, iters, ksize, mode) && (H + kRH - 1) / kRH <= 65535;
}

// history of both recurrences: r^1 .. r^{T-1} and G^0 .. G^{T-1}, fp32 planes
size_t blocked5x5_bwd_workspace(int B, int C, int H, int W, int iters)
{
    return (size_t)(2 * iters - 1) * align256((size_t)B * C * H * W * sizeof(float));
}

// Backward of the 5x5 variant in ceil((T-1)/4) + ceil(T/4) + 1 launches (7 for T = 12) instead of the generic path's 2T + 3:
// forward recompute and adjoint recurrence both run temporally blocked with the weights in registers and keep every
// intermediate plane; the Jacobians are one pass over those planes.
template <typename T>
int blocked5x5_backward(const BwdArgs<T>& a)
{
    const size_t npx = (size_t)a.B * a.C * a.H * a.W;
    const size_t need = blocked5x5_bwd_workspace(a.B, a.C, a.H, a.W, a.iters);
    if (!a.ws || a.ws_bytes < need || ((uintptr_t)a.ws & 15)) return CSPN_ERR_WORKSPACE;
    const size_t pstride = align256(npx * sizeof(float)) / sizeof(float);
    float* rhist = (float*)a.ws;
    float* ghist = rhist + (size_t)(a.iters - 1) * pstride;
    const int vsz = 4 * (int)sizeof(T);
    const bool vec = (a.W % 4 == 0) && ((uintptr_t)a.guidance % vsz == 0) && ((uintptr_t)a.depth % vsz == 0) && (a.gbs % 4 == 0) &&
                     (!a.sparse || (uintptr_t)a.sparse % vsz == 0) && ((uintptr_t)a.grad_out % vsz == 0) &&
                     ((uintptr_t)a.grad_guidance % vsz == 0) && ((uintptr_t)a.grad_depth % vsz == 0);
    int rc = 0;
    for (int done = 0; done < a.iters - 1 && rc == 0;) {                 // r^1 .. r^{T-1}
        const int steps = a.iters - 1 - done < kMaxS ? a.iters - 1 - done : kMaxS;
        if (done == 0) rc = launch_hist<T, T, false>(a, a.depth, rhist, (long long)pstride, steps, vec);
        else rc = launch_hist<T, float, false>(a, rhist + (size_t)(done - 1) * pstride, rhist + (size_t)done * pstride, (long long)pstride, steps, vec);
        done += steps;
    }
    for (int done = 0; done < a.iters && rc == 0;) {                     // G^{T-1} .. G^0
        const int steps = a.iters - done < kMaxS ? a.iters - done : kMaxS;
        const int tcur = a.iters - done;
        if (done == 0) rc = launch_hist<T, T, true>(a, a.grad_out, ghist + (size_t)(tcur - 1) * pstride, -(long long)pstride, steps, vec);
        else rc = launch_hist<T, float, true>(a, ghist + (size_t)tcur * pstride, ghist + (size_t)(tcur - 1) * pstride, -(long long)pstride, steps, vec);
        done += steps;
    }
    if (rc != 0) return rc;
    dim3 grid((unsigned)((a.W + kRW - 1) / kRW), (unsigned)((a.H + kRH - 1) / kRH), (unsigned)a.B);
    jacobian5x5_kernel<T><<<grid, kThreads, 0, a.stream>>>(a.guidance, a.gbs, a.Cg, a.depth, a.grad_out, a.sparse, a.sparse_channels, rhist, ghist,
                                                           pstride, a.grad_guidance, a.grad_depth, a.C, a.H, a.W, a.iters, vec ? 1 : 0);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    ++call_stats().launches;
    return 0;
}

template int blocked5x5_backward<float>(const BwdArgs<float>&);
template int blocked5x5_backward<__half>(const BwdArgs<__half>&);
template int blocked5x5_forward<float>(const FwdArgs<float>&);
template int blocked5x5_forward<__half>(const FwdArgs<__half>&);

}  // namespace cspn
