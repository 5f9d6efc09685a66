// Shared declarations of the CSPN B200 library (internal; the public ABI is include/cspn_b200.h).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cspn_b200.h"

namespace cspn {

constexpr int kMaxTaps = 48;  // 7x7 - 1

// Tap k reads depth at p + (dy[k], dx[k]).  Mode NEW: table of CSPN_new.py:43-67 (+ crop :87);
// mode OURS: row-major K*K taps without the centre (CSPN_ours.py:37-41, pac.py:89).
struct TapTable {
    int n;
    int8_t dy[kMaxTaps];
    int8_t dx[kMaxTaps];
};

inline bool make_taps(int mode, int ksize, TapTable* t)
{
    if (mode == CSPN_MODE_NEW) {
        if (ksize != 3) return false;
        const int8_t ady[8] = {+1, +1, +1, 0, 0, -1, -1, -1};
        const int8_t adx[8] = {+1, 0, -1, +1, -1, +1, 0, -1};
        t->n = 8;
        for (int k = 0; k < 8; ++k) { t->dy[k] = ady[k]; t->dx[k] = adx[k]; }
        return true;
    }
    if (mode != CSPN_MODE_OURS || ksize < 3 || ksize > 7 || (ksize & 1) == 0) return false;
    const int p = ksize / 2;
    int n = 0;
    for (int iy = 0; iy < ksize; ++iy)
        for (int ix = 0; ix < ksize; ++ix) {
            if (iy == p && ix == p) continue;
            t->dy[n] = (int8_t)(iy - p);
            t->dx[n] = (int8_t)(ix - p);
            ++n;
        }
    t->n = n;
    return true;
}

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

__device__ __forceinline__ float signf(float v) { return (float)((v > 0.f) - (v < 0.f)); }

// Problem description shared by all launchers.
template <typename T>
struct FwdArgs {
    const T* guidance; int64_t gbs;
    const T* depth; const T* sparse; int sparse_channels;
    T* out;
    int B, C, H, W, iters, ksize, mode;
    void* ws; size_t ws_bytes;
    cudaStream_t stream;
};
template <typename T>
struct BwdArgs {
    const T* grad_out; const T* guidance; int64_t gbs; int Cg;
    const T* depth; const T* sparse; int sparse_channels;
    T* grad_guidance; T* grad_depth;
    int B, C, H, W, iters, ksize, mode;
    void* ws; size_t ws_bytes;
    cudaStream_t stream;
};

// Per-host-thread bookkeeping reported through cspn_last_path / cspn_last_launch_count.
struct CallStats { int path; int launches; };
CallStats& call_stats();

// generic path (cspn_generic.cu): one launch per iteration, any tap table
size_t generic_fwd_workspace(int B, int C, int H, int W, int taps);
size_t generic_bwd_workspace(int B, int C, int H, int W, int iters, int taps);
template <typename T> int generic_forward(const FwdArgs<T>& a, const TapTable& tt);
template <typename T> int generic_backward(const BwdArgs<T>& a, const TapTable& tt);

// ---- fused path: whole recurrence in one launch, 3x3 only (cspn_fused3x3*.cu) -------------------------------
constexpr long kMaxGlobalExchangeCtas = 1l << 20;  // upper bound on the tiles of a streamed problem (one inbox each)
// Cluster shape (cx x cy CTAs) and grid of cluster tiles (ntx x nty per image) covering an image.  stream: the image is
// one virtual cluster walked by a persistent grid with the halo exchange in global memory (no hardware cluster).
struct Tiling { int cx, cy, ntx, nty, stepx, stepy, ew, eh; long ctas; bool ok; bool stream; bool hyb; };
// What the GPU holds at once for one kernel configuration: sms = CTA slots of the whole GPU (SMs x resident CTAs per
// SM), clusters[n] = co-resident hardware clusters of n CTAs.
struct Capacity { int sms; int clusters[17]; };
bool fused_supported(int B, int C, int H, int W, int iters, int ksize, int mode);
template <typename T> int fused_forward(const FwdArgs<T>& a);   // kDualFallback: no fused kernel takes the problem, use another path
bool fused_single_possible(int B, int C, int H, int W, int iters, int mode);   // the single-tile kernel has a plan (any alignment)
size_t fused_workspace(int B, int C, int H, int W, int iters, int ksize, int mode);   // optional scratch (0 = none)

// dual-slot streamed forward (cspn_dual3x3.cu): two register tiles per CTA worked on alternately so that the halo
// messages of one travel while the other computes; the default 3x3 forward whenever the guidance is TMA-addressable.
constexpr int kDualFallback = -1000;   // internal: "not this kernel" (unaligned guidance etc.), the caller takes the single-tile kernel
bool dual_supported(int B, int C, int H, int W, int iters, int ksize, int mode);
size_t dual_workspace(int B, int C, int H, int W, int iters);
void dual_describe(int B, int C, int H, int W, int iters, int* out9);
void single_describe(int B, int C, int H, int W, int iters, int mode, int* out9);   // {transport, cx, cy, ntx, nty, CTAs per plane, 0, 0, 0}
template <typename T> int dual_forward(const FwdArgs<T>& a);

// temporally blocked forward for the 5x5 variant (cspn_blocked5x5.cu): 4 steps per launch, weights in registers
bool blocked5x5_supported(int B, int C, int H, int W, int iters, int ksize, int mode);
template <typename T> int blocked5x5_forward(const FwdArgs<T>& a);
size_t blocked5x5_workspace(int B, int C, int H, int W, int iters);
// ... and its backward: both recurrences blocked with full history, Jacobians in one pass (7 launches for T = 12)
bool blocked5x5_bwd_supported(int B, int C, int H, int W, int iters, int ksize, int mode);
template <typename T> int blocked5x5_backward(const BwdArgs<T>& a);
size_t blocked5x5_bwd_workspace(int B, int C, int H, int W, int iters);

// fused backward (cspn_fused3x3_bwd.cu): recompute + reverse sweep + Jacobians in one launch, 3x3, one depth channel
bool fused_bwd_supported(int C, int H, int W, int iters, int ksize, int mode);
template <typename T> int fused_backward(const BwdArgs<T>& a);
size_t fused_bwd_workspace(int B, int C, int H, int W, int iters);

// the two output heads upstream of the module (cspn_heads.cu): unpool x2 + conv3x3 without the zeros, both heads in one pass
size_t heads_workspace_bytes();
bool heads_supported(int n1, int n2, int Cin, int h, int w, int H, int W);
template <typename T> int heads_forward(const T* x, const T* w1, const T* w2, T* out1, T* out2, int n1, int n2, int B, int Cin, int h, int w, int H, int W,
                                        cudaStream_t stream);
template <typename T> int heads_backward(const T* x, const T* w1, const T* w2, const T* go1, const T* go2, T* gx, T* gw1, T* gw2, int n1, int n2, int B,
                                         int Cin, int h, int w, int H, int W, void* ws, cudaStream_t stream);

// in-place activated batch normalisation (cspn_abn.cu): fp32, NCHW; act: 0 none, 1 leaky_relu, 2 elu
size_t abn_workspace_bytes(int C);
int abn_stats(const float* x, int N, int C, int S, double* sums, void* ws, cudaStream_t stream);
int abn_finalize(const double* sums, double count, float* mean, float* var, float* running_mean, float* running_var, float momentum, int C, cudaStream_t stream);
int abn_forward(float* x, const float* mean, const float* var, const float* weight, const float* bias, int N, int C, int S, float eps, int act, float slope,
                cudaStream_t stream);
int abn_bwd_reduce(const float* z, const float* dz, const float* weight, const float* bias, int N, int C, int S, float eps, int act, float slope, double* sums,
                   void* ws, cudaStream_t stream);
int abn_bwd_apply(const float* z, const float* dz, float* dx, const float* var, const float* weight, const float* bias, const double* sums, double count_total,
                  double count_local, float* dweight, float* dbias, int N, int C, int S, float eps, int act, float slope, cudaStream_t stream);

// legacy max-of-8 CSPN (cspn_legacy.cu): temporally blocked forward, 4 steps per launch
size_t legacy_workspace(int B, int H, int W, int iters);
template <typename T> int legacy_forward(const T* guidance, int64_t gbs, const T* depth, const T* sparse, T* out, int B, int H, int W, int iters,
                                         void* ws, size_t ws_bytes, cudaStream_t stream);

// loss / metrics downstream of the module (cspn_loss.cu): one deterministic streaming pass each
size_t loss_workspace_bytes();
template <typename T> int masked_l1_forward(const T* pred, const T* target, size_t n, float* loss2, void* ws, size_t ws_bytes, cudaStream_t stream);
template <typename T> int masked_l1_backward(const T* pred, const T* target, size_t n, const float* loss2, const float* grad_loss, T* grad_pred, cudaStream_t stream);
template <typename T> int depth_metrics(const T* pred, const T* target, size_t n, float* out11, void* ws, size_t ws_bytes, cudaStream_t stream);

}  // namespace cspn
