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

// Walk the elements [m0, m1) of channel c (flat index m = n * S + s) and feed them to f(value index in memory).
template <typename F>
__device__ __forceinline__ void for_channel_range(long m0, long m1, int c, int C, int S, F f)
{
    long m = m0;
    while (m < m1) {
        const int n = (int)(m / S), s0 = (int)(m - (long)n * S);
        const int len = (int)((m1 - m) < (long)(S - s0) ? (m1 - m) : (long)(S - s0));
        const size_t base = ((size_t)n * C + c) * S + s0;
        for (int i = threadIdx.x; i < len; i += kAT) f(base + i);
        m += len;
    }
}

// stage 1 of the statistics: partial[c][slab] = {sum x, sum x^2} over the slab's elements
__global__ void __launch_bounds__(kAT) abn_stats_kernel(const float* __restrict__ x, double* __restrict__ partial, int N, int C, int S, int slabs)
{
    __shared__ double sh[2 * (kAT / 32)];
    const int c = blockIdx.y, slab = blockIdx.x;
    const long M = (long)N * S, L = (M + slabs - 1) / slabs;
    const long m0 = (long)slab * L, m1 = m0 + L < M ? m0 + L : M;
    double a = 0.0, b = 0.0;
    for_channel_range(m0, m1, c, C, S, [&](size_t i) { const double v = (double)x[i]; a += v; b = fma(v, v, b); });
    block_sum2(a, b, sh);
    if (threadIdx.x == 0) { partial[((size_t)c * slabs + slab) * 2] = a; partial[((size_t)c * slabs + slab) * 2 + 1] = b; }
}

// stage 1 of the backward: partial = {sum dz', sum y dz'} with the activation undone on the fly
__global__ void __launch_bounds__(kAT) abn_bwd_reduce_kernel(const float* __restrict__ z, const float* __restrict__ dz, const float* __restrict__ weight,
                                                             const float* __restrict__ bias, double* __restrict__ partial, int N, int C, int S, int slabs,
                                                             float eps, int act, float slope)
{
    __shared__ double sh[2 * (kAT / 32)];
    const int c = blockIdx.y, slab = blockIdx.x;
    const long M = (long)N * S, L = (M + slabs - 1) / slabs;
    const long m0 = (long)slab * L, m1 = m0 + L < M ? m0 + L : M;
    const float gamma = weight ? fabsf(weight[c]) + eps : 1.f, beta = bias ? bias[c] : 0.f;
    double a = 0.0, b = 0.0;
    for_channel_range(m0, m1, c, C, S, [&](size_t i) {
        float zz = z[i], d = dz[i];
        act_undo(zz, d, act, slope);
        const float y = (zz - beta) / gamma;
        a += (double)d;
        b += (double)(y * d);
    });
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
        for (int s = s0 + 4 * threadIdx.x; s < s1; s += 4 * kAT) {
            float4 v = *reinterpret_cast<float4*>(p + s);
            v.x = act_forward((v.x - mu) * is * gamma + beta, act, slope);
            v.y = act_forward((v.y - mu) * is * gamma + beta, act, slope);
            v.z = act_forward((v.z - mu) * is * gamma + beta, act, slope);
            v.w = act_forward((v.w - mu) * is * gamma + beta, act, slope);
            *reinterpret_cast<float4*>(p + s) = v;
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
    auto one = [&](float zz, float d) {
        act_undo(zz, d, act, slope);
        const float y = (zz - beta) / gamma;
        return (d - edz - y * eydz) * mul;
    };
    if (vec) {
        for (int s = s0 + 4 * threadIdx.x; s < s1; s += 4 * kAT) {
            const float4 a = *reinterpret_cast<const float4*>(pz + s), g = *reinterpret_cast<const float4*>(pd + s);
            *reinterpret_cast<float4*>(po + s) = make_float4(one(a.x, g.x), one(a.y, g.y), one(a.z, g.z), one(a.w, g.w));
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
    abn_stats_kernel<<<dim3((unsigned)slabs, (unsigned)C), kAT, 0, stream>>>(x, (double*)ws, N, C, S, slabs);
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
    abn_bwd_reduce_kernel<<<dim3((unsigned)slabs, (unsigned)C), kAT, 0, stream>>>(z, dz, weight, bias, (double*)ws, N, C, S, slabs, eps, act, slope);
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
