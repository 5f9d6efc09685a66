// Loss and metrics directly downstream of the CSPN module, each a single pass over (pred, target):
//   masked L1 loss          libs/criterion/criteria.py:27-39 (MaskedL1Loss, used through Criterion_No_DSN :170-188):
//                            mean |target - pred| over the pixels with target > 0; its gradient -sign(target - pred) / count
//   depth metrics           libs/metrics.py:49-83 (Result.evaluate): irmse, imae, mse, rmse, mae, absrel, lg10, delta1-3
// The reference issues ~6 (loss) and ~25 (metrics) ATen ops with boolean-mask gathers, i.e. that many passes over the two
// planes plus a host synchronisation per float().  Here one kernel streams both planes once (8 B/px fp32: HBM-bound) and the
// reduction is DETERMINISTIC: every CTA writes its partial sums to scratch, the last CTA to finish (atomic ticket) adds the
// partials in a fixed order and writes the results - no floating-point atomics, bit-reproducible from run to run.
#include "cspn_common.cuh"

namespace cspn {
namespace {

constexpr int kLossThreads = 256;
constexpr int kLossMaxBlocks = 148 * 8;
constexpr int kMetricSums = 10;          // count, sum|d|, sum d^2, sum|log10 p - log10 t|, sum|d|/t, #d1, #d2, #d3, sum (1/p-1/t)^2, sum|1/p-1/t|

template <int N>
__device__ __forceinline__ void block_reduce(double (&v)[N], double* smem /* [N][kLossThreads / 32] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < N; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
        if (lane == 0) smem[k * (kLossThreads / 32) + warp] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            double x = lane < kLossThreads / 32 ? smem[k * (kLossThreads / 32) + lane] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
            v[k] = x;
        }
    }
}

// scratch layout (doubles): [0] ticket (as unsigned), then per block N partial sums
template <typename T, int N, typename Body, typename Final>
__device__ __forceinline__ void reduce_stream(const T* pred, const T* target, size_t n, double* scratch, Body body, Final fin)
{
    __shared__ double smem[N * (kLossThreads / 32)];
    __shared__ bool last;
    double v[N];
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) body(to_f32(pred[i]), to_f32(target[i]), v);
    block_reduce<N>(v, smem);
    double* part = scratch + 1;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) part[(size_t)blockIdx.x * N + k] = v[k];
        __threadfence();
        const unsigned t = atomicAdd(reinterpret_cast<unsigned*>(scratch), 1u);
        last = t == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x < 32) {
        __threadfence();
        double tot[N];
#pragma unroll
        for (int k = 0; k < N; ++k) {
            double x = 0.0;
            for (unsigned b = threadIdx.x; b < gridDim.x; b += 32) x += part[(size_t)b * N + k];      // fixed order: deterministic
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
            tot[k] = x;
        }
        if (threadIdx.x == 0) {
            fin(tot);
            *reinterpret_cast<unsigned*>(scratch) = 0u;                 // ticket ready for the next launch / graph replay
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kLossThreads) masked_l1_kernel(const T* __restrict__ pred, const T* __restrict__ target, size_t n,
                                                                 float* __restrict__ loss, double* __restrict__ scratch)
{
    reduce_stream<T, 2>(pred, target, n, scratch,
        [](float p, float t, double (&v)[2]) { if (t > 0.f) { v[0] += 1.0; v[1] += (double)fabsf(t - p); } },
        [&](const double (&tot)[2]) { loss[0] = (float)(tot[1] / tot[0]); loss[1] = (float)tot[0]; });        // 0 valid pixels: 0/0 = NaN like mean of empty
}

// d loss / d pred = -sign(target - pred) / count on valid pixels (torch.abs' backward: sign(0) = 0), times the incoming gradient
template <typename T>
__global__ void __launch_bounds__(kLossThreads) masked_l1_bwd_kernel(const T* __restrict__ pred, const T* __restrict__ target, size_t n,
                                                                     const float* __restrict__ loss_count, const float* __restrict__ grad_loss,
                                                                     T* __restrict__ grad_pred)
{
    const float scale = (grad_loss ? grad_loss[0] : 1.f) / loss_count[1];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float p = to_f32(pred[i]), t = to_f32(target[i]);
        grad_pred[i] = from_f32<T>(t > 0.f ? -signf(t - p) * scale : 0.f);
    }
}

template <typename T>
__global__ void __launch_bounds__(kLossThreads) depth_metrics_kernel(const T* __restrict__ pred, const T* __restrict__ target, size_t n,
                                                                     float* __restrict__ out, double* __restrict__ scratch)
{
    reduce_stream<T, kMetricSums>(pred, target, n, scratch,
        [](float p, float t, double (&v)[kMetricSums]) {
            if (t > 0.f) {
                const float d = fabsf(p - t);
                const float ratio = fmaxf(p / t, t / p);                                  // metrics.py:68
                const float inv = fabsf(1.f / p - 1.f / t);                               // :76-78
                v[0] += 1.0; v[1] += (double)d; v[2] += (double)d * d;
                v[3] += (double)fabsf((logf(p) - logf(t)) * 0.43429448190325176f);        // log10 = ln / ln 10 (:14-16)
                v[4] += (double)(d / t);
                v[5] += ratio < 1.25f ? 1.0 : 0.0; v[6] += ratio < 1.5625f ? 1.0 : 0.0; v[7] += ratio < 1.953125f ? 1.0 : 0.0;
                v[8] += (double)inv * inv; v[9] += (double)inv;
            }
        },
        [&](const double (&t)[kMetricSums]) {
            const double c = t[0];
            out[0] = (float)sqrt(t[8] / c);  out[1] = (float)(t[9] / c);                  // irmse, imae
            out[2] = (float)(t[2] / c);      out[3] = (float)sqrt(t[2] / c);  out[4] = (float)(t[1] / c);   // mse, rmse, mae
            out[5] = (float)(t[4] / c);      out[6] = (float)(t[3] / c);                  // absrel, lg10
            out[7] = (float)(t[5] / c);      out[8] = (float)(t[6] / c);  out[9] = (float)(t[7] / c);        // delta1..3
            out[10] = (float)c;
        });
}

inline unsigned loss_blocks(size_t n)
{
    const size_t b = (n + kLossThreads * 8 - 1) / ((size_t)kLossThreads * 8);
    return (unsigned)(b < 1 ? 1 : (b > kLossMaxBlocks ? kLossMaxBlocks : b));
}

}  // namespace

size_t loss_workspace_bytes() { return (1 + (size_t)kLossMaxBlocks * kMetricSums) * sizeof(double); }

template <typename T>
int masked_l1_forward(const T* pred, const T* target, size_t n, float* loss2, void* ws, size_t ws_bytes, cudaStream_t stream)
{
    if (!pred || !target || !loss2) return CSPN_ERR_NULL_POINTER;
    if (!ws || ws_bytes < loss_workspace_bytes() || ((uintptr_t)ws & 7)) return CSPN_ERR_WORKSPACE;
    masked_l1_kernel<T><<<loss_blocks(n), kLossThreads, 0, stream>>>(pred, target, n, loss2, (double*)ws);
    const cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) ++call_stats().launches;
    return (int)e;
}
template <typename T>
int masked_l1_backward(const T* pred, const T* target, size_t n, const float* loss2, const float* grad_loss, T* grad_pred, cudaStream_t stream)
{
    if (!pred || !target || !loss2 || !grad_pred) return CSPN_ERR_NULL_POINTER;
    masked_l1_bwd_kernel<T><<<loss_blocks(n), kLossThreads, 0, stream>>>(pred, target, n, loss2, grad_loss, grad_pred);
    const cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) ++call_stats().launches;
    return (int)e;
}
template <typename T>
int depth_metrics(const T* pred, const T* target, size_t n, float* out11, void* ws, size_t ws_bytes, cudaStream_t stream)
{
    if (!pred || !target || !out11) return CSPN_ERR_NULL_POINTER;
    if (!ws || ws_bytes < loss_workspace_bytes() || ((uintptr_t)ws & 7)) return CSPN_ERR_WORKSPACE;
    depth_metrics_kernel<T><<<loss_blocks(n), kLossThreads, 0, stream>>>(pred, target, n, out11, (double*)ws);
    const cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) ++call_stats().launches;
    return (int)e;
}

template int masked_l1_forward<float>(const float*, const float*, size_t, float*, void*, size_t, cudaStream_t);
template int masked_l1_forward<__half>(const __half*, const __half*, size_t, float*, void*, size_t, cudaStream_t);
template int masked_l1_backward<float>(const float*, const float*, size_t, const float*, const float*, float*, cudaStream_t);
template int masked_l1_backward<__half>(const __half*, const __half*, size_t, const float*, const float*, __half*, cudaStream_t);
template int depth_metrics<float>(const float*, const float*, size_t, float*, void*, size_t, cudaStream_t);
template int depth_metrics<__half>(const __half*, const __half*, size_t, float*, void*, size_t, cudaStream_t);

}  // namespace cspn
