// In-place activated batch normalisation - the reference's only native code:
//   network/libs/inplace_abn/functions.py:70-163 (InPlaceABN.forward / backward), :166-297 (InPlaceABNSync)
//   network/libs/inplace_abn/src/bn.cu:125-232 (mean_var / forward / edz_eydz / backward kernels), :290-377 (activations)
// used by both UNets when more than one GPU is present (unet_cspn_nyu.py:19-24).  Semantics kept:
//   mean, var (biased) per channel over N x S;  gamma = |weight| + eps, beta = bias;  invstd = 1/sqrt(var + eps) (0 if var = eps = 0)
//   z = act((x - mean) * invstd * gamma + beta) written OVER x;  act in {leaky_relu(slope), elu, none}
//   backward from the saved OUTPUT: undo the activation (leaky: z/slope, dz*slope for z < 0; elu: log1p(z), dz*(z+1)),
//   y = (z - beta)/gamma, edz = E[dz], eydz = E[y dz], dx = (dz - edz - y eydz) gamma invstd,
//   dweight = sign(weight) eydz count, dbias = edz count;  running stats with the n/(n-1) correction.
// What is different from the reference's kernels (one CTA per channel = 64 CTAs for a 64-channel layer, two passes over x for
// the statistics, six separate passes for activation / undo / reduce / apply):
//   * reductions run on a (slab, channel) grid sized to the 148 SMs, one pass (sum and sum of squares accumulated in double),
//     per-CTA partials added in a fixed order by a second tiny kernel: deterministic, no floating-point atomics;
//   * the per-channel sums leave the device-side pipeline as a [2C] double vector - the synchronised variant all-reduces
//     exactly that vector across ranks (NCCL, one process per GPU) between the two stages instead of the reference's
//     master/worker queues over DataParallel threads;
//   * normalise + affine + activation is one float4 pass in place; the backward recomputes the activation undo inside both of
//     its passes instead of rewriting z and dz in memory.
// HBM traffic per element: forward 4 (stats) + 8 (normalise) bytes, backward 8 (reduce) + 12 (apply) bytes.
#include "cspn_common.cuh"

namespace cspn {
namespace {

constexpr int kAT = 256;                 // threads per CTA
constexpr int kMaxSlabs = 64;            // per-channel split of the reductions
constexpr int kSegElems = 4096;          // elements of a plane per CTA in the elementwise kernels

enum { kActNone = 0, kActLeaky = 1, kActElu = 2 };

__device__ __forceinline__ float act_forward(float z, int act, float slope)
{
    if (act == kActLeaky) return z < 0.f ? z * slope : z;
    if (act == kActElu) return z < 0.f ? expf(z) - 1.f : z;
    return z;
}
// undo the activation on the saved output and route the incoming gradient through it
__device__ __forceinline__ void act_undo(float& z, float& dz, int act, float slope)
{
    if (act == kActLeaky) { if (z < 0.f) { dz *= slope; z *= 1.f / slope; } }
    else if (act == kActElu) { if (z < 0.f) { dz *= z + 1.f; z = log1pf(z); } }
}

__device__ __forceinline__ void block_sum2(double& a, double& b, double* sh /* [2][kAT / 32] */)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sh[warp] = a; sh[kAT / 32 + warp] = b; }
    __syncthreads();
    if (warp == 0) {
        a = lane < kAT / 32 ? sh[lane] : 0.0;
        b = lane < kAT / 32 ? sh[kAT / 32 + lane] : 0.0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
    }
}

// Walk the elements [m0, m1) of channel c (flat index m = n * S + s): piece by piece (one piece per image the range touches), with
// four independent 16-byte loads per thread in flight when vec (S, m0, m1 multiples of 4 and 16-byte aligned planes), else scalar.
// ld4(i) / ld1(i) load at float index i, use4 / use1 consume what they returned.
template <typename L4, typename U4, typename L1, typename U1>
__device__ __forceinline__ void for_channel_range(long m0, long m1, int c, int C, int S, bool vec, L4 ld4, U4 use4, L1 ld1, U1 use1)
{
    long m = m0;
    while (m < m1) {
        const int n = (int)(m / S), s0 = (int)(m - (long)n * S);
        const int len = (int)((m1 - m) < (long)(S - s0) ? (m1 - m) : (long)(S - s0));
        const size_t base = ((size_t)n * C + c) * S + s0;
        if (vec) {
            int i = 4 * threadIdx.x;
            for (; i + 12 * kAT < len; i += 16 * kAT) {
                const auto v0 = ld4(base + i), v1 = ld4(base + i + 4 * kAT), v2 = ld4(base + i + 8 * kAT), v3 = ld4(base + i + 12 * kAT);
                use4(v0); use4(v1); use4(v2); use4(v3);
            }
            for (; i < len; i += 4 * kAT) use4(ld4(base + i));
        } else {
            for (int i = threadIdx.x; i < len; i += kAT) use1(ld1(base + i));
        }
        m += len;
    }
}

// slab s of a channel covers [s * L, (s + 1) * L) of its N * S elements; L is a multiple of 4 so that vector pieces stay aligned
__device__ __forceinline__ void slab_range(int N, int S, int slabs, int slab, long& m0, long& m1)
{
    const long M = (long)N * S;
    const long L = (((M + slabs - 1) / slabs) + 3) & ~3l;
    m0 = (long)slab * L < M ? (long)slab * L : M;
    m1 = m0 + L < M ? m0 + L : M;
}

// stage 1 of the statistics: partial[c][slab] = {sum x, sum x^2} over the slab's elements (groups of four summed in fp32, groups in double)
__global__ void __launch_bounds__(kAT) abn_stats_kernel(const float* __restrict__ x, double* __restrict__ partial, int N, int C, int S, int slabs, int vec)
{
    __shared__ double sh[2 * (kAT / 32)];
    const int c = blockIdx.y, slab = blockIdx.x;
    long m0, m1;
    slab_range(N, S, slabs, slab, m0, m1);
    double a = 0.0, b = 0.0;
    for_channel_range(m0, m1, c, C, S, vec != 0,
        [&](size_t i) { return __ldg(reinterpret_cast<const float4*>(x + i)); },
        [&](float4 v) { a += (double)((v.x + v.y) + (v.z + v.w)); b += (double)(fmaf(v.x, v.x, v.y * v.y) + fmaf(v.z, v.z, v.w * v.w)); },
        [&](size_t i) { return __ldg(x + i); },
        [&](float v) { a += (double)v; b = fma((double)v, (double)v, b); });
    block_sum2(a, b, sh);
    if (threadIdx.x == 0) { partial[((size_t)c * slabs + slab) * 2] = a; partial[((size_t)c * slabs + slab) * 2 + 1] = b; }
}

// stage 1 of the backward: partial = {sum dz', sum y dz'} with the activation undone on the fly
struct ZD4 { float4 z, d; };
struct ZD1 { float z, d; };
__global__ void __launch_bounds__(kAT) abn_bwd_reduce_kernel(const float* __restrict__ z, const float* __restrict__ dz, const float* __restrict__ weight,
                                                             const float* __restrict__ bias, double* __restrict__ partial, int N, int C, int S, int slabs,
                                                             float eps, int act, float slope, int vec)
{
    __shared__ double sh[2 * (kAT / 32)];
    const int c = blockIdx.y, slab = blockIdx.x;
    long m0, m1;
    slab_range(N, S, slabs, slab, m0, m1);
    const float gamma = weight ? fabsf(weight[c]) + eps : 1.f, beta = bias ? bias[c] : 0.f;
    const float inv_gamma = 1.f / gamma;               // y = (z - beta) * (1 / gamma): one rounding more than the reference's division, 30 % fewer instructions
    double a = 0.0, b = 0.0;
    auto one = [&](float zz, float d, float& sd, float& syd) {
        act_undo(zz, d, act, slope);
        const float y = (zz - beta) * inv_gamma;
        sd += d;
        syd = fmaf(y, d, syd);
    };
    for_channel_range(m0, m1, c, C, S, vec != 0,
        [&](size_t i) { return ZD4{__ldg(reinterpret_cast<const float4*>(z + i)), __ldg(reinterpret_cast<const float4*>(dz + i))}; },
        [&](ZD4 v) {
            float sd = 0.f, syd = 0.f;
            one(v.z.x, v.d.x, sd, syd); one(v.z.y, v.d.y, sd, syd); one(v.z.z, v.d.z, sd, syd); one(v.z.w, v.d.w, sd, syd);
            a += (double)sd; b += (double)syd;
        },
        [&](size_t i) { return ZD1{__ldg(z + i), __ldg(dz + i)}; },
        [&](ZD1 v) { float sd = 0.f, syd = 0.f; one(v.z, v.d, sd, syd); a += (double)sd; b += (double)syd; });
    block_sum2(a, b, sh);
    if (threadIdx.x == 0) { partial[((size_t)c * slabs + slab) * 2] = a; partial[((size_t)c * slabs + slab) * 2 + 1] = b; }
}

// stage 2 of both reductions: sums[c] = fixed-order sum of the partials (what the synchronised variant all-reduces)
__global__ void abn_sum_partials_kernel(const double* __restrict__ partial, double* __restrict__ sums, int C, int slabs)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double a = 0.0, b = 0.0;
    for (int s = 0; s < slabs; ++s) { a += partial[((size_t)c * slabs + s) * 2]; b += partial[((size_t)c * slabs + s) * 2 + 1]; }
    sums[2 * c] = a;
    sums[2 * c + 1] = b;
}

// mean / var from the (possibly all-reduced) sums; running statistics as functions.py:90-92
__global__ void abn_finalize_kernel(const double* __restrict__ sums, double count, float* __restrict__ mean, float* __restrict__ var,
                                    float* __restrict__ running_mean, float* __restrict__ running_var, float momentum, int C)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double m = sums[2 * c] / count;
    double v = sums[2 * c + 1] / count - m * m;
    v = v < 0.0 ? 0.0 : v;
    mean[c] = (float)m;
    var[c] = (float)v;
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
    if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(v * count / (count - 1.0));
}

__device__ __forceinline__ float inv_std(float var, float eps) { return (var != 0.f || eps != 0.f) ? 1.f / sqrtf(var + eps) : 0.f; }

// z = act((x - mean) invstd gamma + beta), in place
__global__ void __launch_bounds__(kAT) abn_forward_kernel(float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ var,
                                                          const float* __restrict__ weight, const float* __restrict__ bias, int C, int S, float eps, int act,
                                                          float slope, int vec)
{
    const int plane = blockIdx.x, c = plane % C;
    const float mu = mean[c], is = inv_std(var[c], eps);
    const float gamma = weight ? fabsf(weight[c]) + eps : 1.f, beta = bias ? bias[c] : 0.f;
    float* p = x + (size_t)plane * S;
    const int s0 = blockIdx.y * kSegElems, s1 = s0 + kSegElems < S ? s0 + kSegElems : S;
    if (vec) {
        // a segment is 4 x 16 bytes per thread: all four loads are issued before the first store
        constexpr int kPer = kSegElems / (4 * kAT);
        float4 v[kPer];
#pragma unroll
        for (int k = 0; k < kPer; ++k) { const int s = s0 + 4 * (threadIdx.x + k * kAT); if (s < s1) v[k] = *reinterpret_cast<const float4*>(p + s); }
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int s = s0 + 4 * (threadIdx.x + k * kAT);
            if (s >= s1) continue;
            v[k].x = act_forward((v[k].x - mu) * is * gamma + beta, act, slope);
            v[k].y = act_forward((v[k].y - mu) * is * gamma + beta, act, slope);
            v[k].z = act_forward((v[k].z - mu) * is * gamma + beta, act, slope);
            v[k].w = act_forward((v[k].w - mu) * is * gamma + beta, act, slope);
            *reinterpret_cast<float4*>(p + s) = v[k];
        }
    } else {
        for (int s = s0 + threadIdx.x; s < s1; s += kAT) p[s] = act_forward((p[s] - mu) * is * gamma + beta, act, slope);
    }
}

// dx = (dz' - edz - y eydz) gamma invstd; sums == nullptr: inference mode (edz = eydz = 0, functions.py:147-150)
__global__ void __launch_bounds__(kAT) abn_bwd_apply_kernel(const float* __restrict__ z, const float* __restrict__ dz, float* __restrict__ dx,
                                                            const float* __restrict__ var, const float* __restrict__ weight, const float* __restrict__ bias,
                                                            const double* __restrict__ sums, double count_total, double count_local,
                                                            float* __restrict__ dweight, float* __restrict__ dbias, int C, int S, float eps, int act,
                                                            float slope, int vec)
{
    const int plane = blockIdx.x, c = plane % C;
    const float gamma = weight ? fabsf(weight[c]) + eps : 1.f, beta = bias ? bias[c] : 0.f;
    const double edz_d = sums ? sums[2 * c] / count_total : 0.0, eydz_d = sums ? sums[2 * c + 1] / count_total : 0.0;
    const float edz = (float)edz_d, eydz = (float)eydz_d;
    if (plane < C && blockIdx.y == 0 && threadIdx.x == 0) {          // first image's CTA of every channel also writes the parameter gradients
        if (dweight) { const float wv = weight[c]; dweight[c] = wv > 0.f ? (float)(eydz_d * count_local) : (wv < 0.f ? (float)(-eydz_d * count_local) : 0.f); }
        if (dbias) dbias[c] = (float)(edz_d * count_local);
    }
    if (!dx) return;
    const float mul = gamma * inv_std(var[c], eps);
    const float* pz = z + (size_t)plane * S;
    const float* pd = dz + (size_t)plane * S;
    float* po = dx + (size_t)plane * S;
    const int s0 = blockIdx.y * kSegElems, s1 = s0 + kSegElems < S ? s0 + kSegElems : S;
    const float inv_gamma = 1.f / gamma;
    auto one = [&](float zz, float d) {
        act_undo(zz, d, act, slope);
        const float y = (zz - beta) * inv_gamma;
        return (d - edz - y * eydz) * mul;
    };
    if (vec) {
        constexpr int kPer = kSegElems / (4 * kAT);
        float4 a[kPer], g[kPer];
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int s = s0 + 4 * (threadIdx.x + k * kAT);
            if (s < s1) { a[k] = *reinterpret_cast<const float4*>(pz + s); g[k] = *reinterpret_cast<const float4*>(pd + s); }
        }
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int s = s0 + 4 * (threadIdx.x + k * kAT);
            if (s < s1) *reinterpret_cast<float4*>(po + s) = make_float4(one(a[k].x, g[k].x), one(a[k].y, g[k].y), one(a[k].z, g[k].z), one(a[k].w, g[k].w));
        }
    } else {
        for (int s = s0 + threadIdx.x; s < s1; s += kAT) po[s] = one(pz[s], pd[s]);
    }
}

inline int slabs_for(int N, int C, int S)
{
    const long M = (long)N * S;
    long want = (4 * 148 + C - 1) / C;                        // ~4 CTAs per SM over all channels
    const long cap = (M + 2047) / 2048;                       // at least 2k elements per CTA
    if (want > cap) want = cap;
    if (want > kMaxSlabs) want = kMaxSlabs;
    return (int)(want < 1 ? 1 : want);
}

#define ABN_CHECK_LAUNCH()                                   \
    do {                                                     \
        const cudaError_t e_ = cudaGetLastError();           \
        if (e_ != cudaSuccess) return (int)e_;               \
        ++call_stats().launches;                             \
    } while (0)

}  // namespace

size_t abn_workspace_bytes(int C) { return (size_t)C * kMaxSlabs * 2 * sizeof(double); }

int abn_stats(const float* x, int N, int C, int S, double* sums, void* ws, cudaStream_t stream)
{
    const int slabs = slabs_for(N, C, S);
    const int vec = (S % 4 == 0) && ((uintptr_t)x % 16 == 0);
    abn_stats_kernel<<<dim3((unsigned)slabs, (unsigned)C), kAT, 0, stream>>>(x, (double*)ws, N, C, S, slabs, vec);
    ABN_CHECK_LAUNCH();
    abn_sum_partials_kernel<<<(C + 127) / 128, 128, 0, stream>>>((const double*)ws, sums, C, slabs);
    ABN_CHECK_LAUNCH();
    return 0;
}

int abn_finalize(const double* sums, double count, float* mean, float* var, float* running_mean, float* running_var, float momentum, int C,
                 cudaStream_t stream)
{
    abn_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(sums, count, mean, var, running_mean, running_var, momentum, C);
    ABN_CHECK_LAUNCH();
    return 0;
}

int abn_forward(float* x, const float* mean, const float* var, const float* weight, const float* bias, int N, int C, int S, float eps, int act,
                float slope, cudaStream_t stream)
{
    const int vec = (S % 4 == 0) && ((uintptr_t)x % 16 == 0);
    abn_forward_kernel<<<dim3((unsigned)(N * C), (unsigned)((S + kSegElems - 1) / kSegElems)), kAT, 0, stream>>>(x, mean, var, weight, bias, C, S, eps, act,
                                                                                                                 slope, vec);
    ABN_CHECK_LAUNCH();
    return 0;
}

int abn_bwd_reduce(const float* z, const float* dz, const float* weight, const float* bias, int N, int C, int S, float eps, int act, float slope,
                   double* sums, void* ws, cudaStream_t stream)
{
    const int slabs = slabs_for(N, C, S);
    const int vec = (S % 4 == 0) && ((uintptr_t)z % 16 == 0) && ((uintptr_t)dz % 16 == 0);
    abn_bwd_reduce_kernel<<<dim3((unsigned)slabs, (unsigned)C), kAT, 0, stream>>>(z, dz, weight, bias, (double*)ws, N, C, S, slabs, eps, act, slope, vec);
    ABN_CHECK_LAUNCH();
    abn_sum_partials_kernel<<<(C + 127) / 128, 128, 0, stream>>>((const double*)ws, sums, C, slabs);
    ABN_CHECK_LAUNCH();
    return 0;
}

int abn_bwd_apply(const float* z, const float* dz, float* dx, const float* var, const float* weight, const float* bias, const double* sums,
                  double count_total, double count_local, float* dweight, float* dbias, int N, int C, int S, float eps, int act, float slope,
                  cudaStream_t stream)
{
    const int vec = (S % 4 == 0) && ((uintptr_t)z % 16 == 0) && ((uintptr_t)dz % 16 == 0) && ((uintptr_t)dx % 16 == 0);
    abn_bwd_apply_kernel<<<dim3((unsigned)(N * C), (unsigned)((S + kSegElems - 1) / kSegElems)), kAT, 0, stream>>>(
        z, dz, dx, var, weight, bias, sums, count_total, count_local, dweight, dbias, C, S, eps, act, slope, vec);
    ABN_CHECK_LAUNCH();
    return 0;
}

}  // namespace cspn
