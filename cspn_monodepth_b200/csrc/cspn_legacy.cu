// Legacy max-of-8 CSPN: network/libs/post_process/CSPN.py:19-56 (AffinityPropagate, with sparse samples)
// and :132-164 (AffinityPropagate_prediction, without).  Per step and gate k (|guidance[:, k]|, NOT shifted):
//     out_k(p) = box3x3(g_k * r)(p) / box3x3(g_k)(p)      zero padded, centre included (CSPN.py:81-102)
//     r(p)     = max_k out_k(p), NaN propagating           (:48-51, :113-123)
//     r        = (1 - m) r + m * sparse                    (:53; the SPARSE SAMPLE, which also seeds r^0 at :33)
// The reference issues 16 x (8 x (2 conv2d + mul + div) + 7 max + 4 blend ops) = ~700 ATen launches with 8 full-size
// temporaries per step.  Here the recurrence is temporally blocked like the 5x5 kernel: a CTA owns a 64 x 32 region, a thread 4
// adjacent pixels whose 8 gates and 8 box sums stay in registers, S = 4 steps per launch:
//   * per step the 8 products g_k * r get their horizontal 3-sums in registers (the two outer terms come from the lane
//     neighbours by shuffle), go to shared memory once (8 STS.128), and the vertical 3-sums come back with 24 LDS.128 - the box
//     filter is separable, 6 adds per gate instead of 8;
//   * the divisions are IEEE (the max picks between near-equal candidates; a reciprocal would flip winners);
//   * two sets of planes alternate, one __syncthreads per step; after S steps the inner 56 x (32 - 2S) block is exact.
// HBM traffic per launch: 8 + 3 planes read (region overlaps are L2 hits), one written.  No inter-CTA synchronisation.
#include "cspn_common.cuh"

namespace cspn {

namespace {

constexpr int kRW = 64, kRH = 32, kQuad = 4;
constexpr int kThreads = (kRW / kQuad) * kRH;        // 512
constexpr int kSW = kRW, kSH = kRH + 2;              // one zero row above and below
constexpr int kMaxS = 4;
constexpr int kHaloX = 4;                            // fixed: region origins stay 16-byte aligned
constexpr int kOutW = kRW - 2 * kHaloX;              // 56
constexpr size_t kSmemBytes = (size_t)2 * 8 * kSH * kSW * sizeof(float);     // 139,264 B

__device__ __forceinline__ void unpack4(float4 v, float* o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
__device__ __forceinline__ void unpack4(uint2 v, float* o)
{
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
template <typename T> struct Vec4;
template <> struct Vec4<float> { typedef float4 type; };
template <> struct Vec4<__half> { typedef uint2 type; };

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

// horizontal 3-sum of a quad; the outer terms come from the lane neighbours (16 quads per region row, two rows per warp)
__device__ __forceinline__ float4 hsum3(const float* v, int qx)
{
    float l = __shfl_up_sync(0xffffffffu, v[3], 1), r = __shfl_down_sync(0xffffffffu, v[0], 1);
    if (qx == 0) l = 0.f;                       // region edge: inside the halo, never part of an exact result
    if (qx == kRW / kQuad - 1) r = 0.f;
    return make_float4(l + v[0] + v[1], v[0] + v[1] + v[2], v[1] + v[2] + v[3], v[2] + v[3] + r);
}

template <typename T, typename TIn, typename TOut>
__global__ void __launch_bounds__(kThreads, 1)
legacy_kernel(const T* __restrict__ g, int64_t gbs, const TIn* __restrict__ rin, const T* __restrict__ d0, const T* __restrict__ sparse,
              TOut* __restrict__ rout, int H, int W, int steps, int first, int vec)
{
    extern __shared__ __align__(16) float smem[];
    float (*hp)[8][kSH][kSW] = reinterpret_cast<float (*)[8][kSH][kSW]>(smem);
    const int tid = threadIdx.x;
    const int qx = tid % (kRW / kQuad), ry = tid / (kRW / kQuad);
    const int halo_y = steps, out_h = kRH - 2 * halo_y;
    const int b = blockIdx.z;
    const int ox = blockIdx.x * kOutW - kHaloX, oy = blockIdx.y * out_h - halo_y;
    const int gx = ox + qx * kQuad, gy = oy + ry;
    const size_t hw = (size_t)H * W;

    // zero rows above / below both sets of planes
    for (int i = tid; i < 2 * 8 * 2 * kSW; i += kThreads) {
        const int x = i % kSW, rowsel = (i / kSW) & 1, k = (i / (2 * kSW)) & 7, set = i / (16 * kSW);
        hp[set][k][rowsel ? kSH - 1 : 0][x] = 0.f;
    }

    float gk[8][kQuad], wsum[8][kQuad];
    {
        const T* gb = g + (size_t)b * gbs;
#pragma unroll
        for (int k = 0; k < 8; ++k) load_quad(gb + (size_t)k * hw, gy, gx, H, W, vec != 0, gk[k]);      // all loads first: the shuffles below would serialise them
    }
    float rq[kQuad], mq[kQuad], sq[kQuad], in[kQuad];
    load_quad(rin + (size_t)b * hw, gy, gx, H, W, vec != 0, rq);
    mq[0] = mq[1] = mq[2] = mq[3] = 0.f;
    sq[0] = sq[1] = sq[2] = sq[3] = 0.f;
    if (sparse) load_quad(sparse + (size_t)b * hw, gy, gx, H, W, vec != 0, sq);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int q = 0; q < kQuad; ++q) gk[k][q] = fabsf(gk[k][q]);                           // CSPN.py:22-29
        *reinterpret_cast<float4*>(&hp[0][k][ry + 1][qx * kQuad]) = hsum3(gk[k], qx);
    }
#pragma unroll
    for (int q = 0; q < kQuad; ++q) {
        in[q] = (gy >= 0 && gy < H && gx + q >= 0 && gx + q < W) ? 1.f : 0.f;
        mq[q] = signf(sq[q]);                                                                  // :31
        if (first) rq[q] = (1.f - mq[q]) * rq[q] + mq[q] * sq[q];                              // :33
        (void)d0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float a[4], c[4], e[4];
        unpack4(*reinterpret_cast<const float4*>(&hp[0][k][ry][qx * kQuad]), a);
        unpack4(*reinterpret_cast<const float4*>(&hp[0][k][ry + 1][qx * kQuad]), c);
        unpack4(*reinterpret_cast<const float4*>(&hp[0][k][ry + 2][qx * kQuad]), e);
#pragma unroll
        for (int q = 0; q < kQuad; ++q) wsum[k][q] = a[q] + c[q] + e[q];                        // box3x3(g_k), zero padded (:98)
    }
    // (the first step writes set 1, so set 0 is free again after the barrier inside the loop)

    for (int s = 0; s < steps; ++s) {
        const int set = (s + 1) & 1;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float pr[kQuad];
#pragma unroll
            for (int q = 0; q < kQuad; ++q) pr[q] = gk[k][q] * rq[q];                          // :99 weight_matrix * blur_matrix
            *reinterpret_cast<float4*>(&hp[set][k][ry + 1][qx * kQuad]) = hsum3(pr, qx);
        }
        __syncthreads();
        float best[kQuad];
        bool bad[kQuad];
#pragma unroll
        for (int q = 0; q < kQuad; ++q) { best[q] = -INFINITY; bad[q] = false; }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float a[4], c[4], e[4];
            unpack4(*reinterpret_cast<const float4*>(&hp[set][k][ry][qx * kQuad]), a);
            unpack4(*reinterpret_cast<const float4*>(&hp[set][k][ry + 1][qx * kQuad]), c);
            unpack4(*reinterpret_cast<const float4*>(&hp[set][k][ry + 2][qx * kQuad]), e);
#pragma unroll
            for (int q = 0; q < kQuad; ++q) {
                const float o = (a[q] + c[q] + e[q]) / wsum[k][q];                             // :101, IEEE division
                bad[q] = bad[q] || (o != o);                                                   // torch.max propagates NaN
                best[q] = fmaxf(best[q], o);
            }
        }
#pragma unroll
        for (int q = 0; q < kQuad; ++q) {
            const float r = bad[q] ? __int_as_float(0x7fc00000) : best[q];
            rq[q] = in[q] != 0.f ? (1.f - mq[q]) * r + mq[q] * sq[q] : 0.f;                    // :53; outside the image stays zero padding
        }
    }

    const int lx = qx * kQuad;
    if (ry >= halo_y && ry < kRH - halo_y && lx >= kHaloX && lx < kRW - kHaloX && gy < H) {
        TOut* orow = rout + (size_t)b * hw + (size_t)gy * W;
#pragma unroll
        for (int q = 0; q < kQuad; ++q)
            if (gx + q < W) orow[gx + q] = from_f32<TOut>(rq[q]);
    }
}

template <typename T, typename TIn, typename TOut>
int launch_one(const T* g, int64_t gbs, const TIn* rin, const T* depth, const T* sparse, TOut* rout, int B, int H, int W, int steps, bool first,
               bool vec, cudaStream_t stream)
{
    auto kern = legacy_kernel<T, TIn, TOut>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    const int out_h = kRH - 2 * steps;
    dim3 grid((unsigned)((W + kOutW - 1) / kOutW), (unsigned)((H + out_h - 1) / out_h), (unsigned)B);
    kern<<<grid, kThreads, kSmemBytes, stream>>>(g, gbs, rin, depth, sparse, rout, H, W, steps, first ? 1 : 0, vec ? 1 : 0);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    ++call_stats().launches;
    return 0;
}

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

size_t legacy_workspace(int B, int H, int W, int iters)
{
    if (iters <= kMaxS) return 0;
    return (iters <= 2 * kMaxS ? 1 : 2) * align256((size_t)B * H * W * sizeof(float));
}

template <typename T>
int legacy_forward(const T* guidance, int64_t gbs, const T* depth, const T* sparse, T* out, int B, int H, int W, int iters, void* ws,
                   size_t ws_bytes, cudaStream_t stream)
{
    const size_t npx = (size_t)B * H * W;
    const size_t need = legacy_workspace(B, H, W, iters);
    if (need && (!ws || ws_bytes < need || ((uintptr_t)ws & 15))) return CSPN_ERR_WORKSPACE;
    if ((long)B > 65535 || (H + kRH - 2 * kMaxS - 1) / (kRH - 2 * kMaxS) > 65535) return CSPN_ERR_BAD_SHAPE;
    float* r0 = (float*)ws;
    float* r1 = (float*)((char*)ws + align256(npx * sizeof(float)));
    const int vsz = 4 * (int)sizeof(T);
    const bool vec = (W % 4 == 0) && ((uintptr_t)guidance % vsz == 0) && ((uintptr_t)depth % vsz == 0) && (gbs % 4 == 0) &&
                     (!sparse || (uintptr_t)sparse % vsz == 0) && ((uintptr_t)out % vsz == 0);
    int done = 0, rc = 0;
    const float* cur = nullptr;
    while (done < iters && rc == 0) {
        const int steps = iters - done < kMaxS ? iters - done : kMaxS;
        const bool first = done == 0, last = done + steps == iters;
        float* nxt = cur == r0 ? r1 : r0;
        if (first && last) rc = launch_one<T, T, T>(guidance, gbs, depth, depth, sparse, out, B, H, W, steps, true, vec, stream);
        else if (first) rc = launch_one<T, T, float>(guidance, gbs, depth, depth, sparse, nxt, B, H, W, steps, true, vec, stream);
        else if (last) rc = launch_one<T, float, T>(guidance, gbs, cur, depth, sparse, out, B, H, W, steps, false, vec, stream);
        else rc = launch_one<T, float, float>(guidance, gbs, cur, depth, sparse, nxt, B, H, W, steps, false, vec, stream);
        cur = nxt;
        done += steps;
    }
    return rc;
}

template int legacy_forward<float>(const float*, int64_t, const float*, const float*, float*, int, int, int, int, void*, size_t, cudaStream_t);
template int legacy_forward<__half>(const __half*, int64_t, const __half*, const __half*, __half*, int, int, int, int, void*, size_t, cudaStream_t);

}  // namespace cspn
