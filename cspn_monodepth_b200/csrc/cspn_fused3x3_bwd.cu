// Backward instantiations and host entry points of the fused 3x3 CSPN kernel (cspn_fused3x3.cuh, BWD = true):
// forward recompute with history + reverse sweep + normalisation / abs / softmax Jacobian in ONE launch.
// Replaces autograd through CSPN_new.py:80-90 and Conv2dFn.backward (pac.py:96-121) x T.
#include <algorithm>

#include "cspn_fused3x3.cuh"

namespace cspn {

namespace {
constexpr int kTHBwd = kNWBwd * kPBwd;
inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }
// History scratch slots of a launch: one per CTA when the launch has at most kHistSlots CTAs (always in stream mode: the persistent
// grid has at most one CTA per SM), else one per SM id.
inline int hist_slots_for(const Tiling& tl, int B, const Capacity& cap)
{
    const long total = tl.ctas * (long)B;
    const long ctas = tl.stream ? (total < cap.sms ? total : (long)cap.sms) : total;
    return ctas <= kHistSlots ? (int)ctas : kHistSlots;
}
inline bool hist_by_cta(const Tiling& tl, int B, const Capacity& cap)
{
    const long total = tl.ctas * (long)B;
    return (tl.stream ? (total < cap.sms ? total : (long)cap.sms) : total) <= kHistSlots;
}
inline size_t hist_bytes(int iters, int slots) { return (size_t)slots * (size_t)iters * kTHBwd * kTileW * sizeof(float); }
// grad_guidance of one more depth channel, accumulated into the result (the affinity is shared by the channels of an image,
// pac.py:118-119): channels 0..7 only, the rest stays zero
template <typename T>
__global__ void accumulate_gg_kernel(T* __restrict__ gg, const T* __restrict__ add, size_t per_image, size_t image_stride, size_t total)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / per_image, r = i - b * per_image;
        const size_t o = b * image_stride + r;
        gg[o] = from_f32<T>(to_f32(gg[o]) + to_f32(add[i]));
    }
}
// scratch for the extra channels' grad_guidance: 8 channels per image, sized for fp32 (the query does not know the dtype)
inline size_t gg_scratch_bytes(int B, int C, int H, int W) { return C > 1 ? up256((size_t)B * 8 * H * W * sizeof(float)) : 0; }

inline size_t bwd_inbox_bytes(const Tiling& tl, int B)
{
    if (!tl.stream) return 0;
    return up256(kStatusBytes + (size_t)(tl.ctas * (long)B) * inbox_bytes<kTHBwd>());     // status word + one inbox per tile
}
}  // namespace

bool fused_bwd_supported(int C, int H, int W, int iters, int ksize, int mode)
{
    (void)mode;
    // several depth channels share the affinity: one launch per channel plus an accumulation of grad_guidance; refresh tags fit 7 bits
    if (C < 1 || C > 16 || ksize != 3 || iters < 1 || iters > 60) return false;
    if ((long)H * W > (1l << 30)) return false;
    return choose_tiling(H, W, iters, kTHBwd, 1, default_capacity()).ok;
}

size_t fused_bwd_workspace(int B, int C, int H, int W, int iters)
{
    const Capacity cap = capacity<kPBwd, kNWBwd, true>();
    const Tiling tl = choose_tiling(H, W, iters, kTHBwd, (long)B, cap);
    if (!tl.ok) return 0;
    return bwd_inbox_bytes(tl, B) + hist_bytes(iters, hist_slots_for(tl, B, cap)) + gg_scratch_bytes(B, C, H, W);
}

template <typename T>
int fused_backward(const BwdArgs<T>& a)
{
    const Capacity cap = capacity<kPBwd, kNWBwd, true>();
    const Tiling tl = choose_tiling(a.H, a.W, a.iters, kTHBwd, (long)a.B, cap);
    if (!tl.ok || a.C < 1 || a.C > 16) return CSPN_ERR_BAD_KERNEL_SIZE;
    if ((long)a.B > 65535) return CSPN_ERR_BAD_SHAPE;
    const int slots = hist_slots_for(tl, a.B, cap);
    const size_t inbox = bwd_inbox_bytes(tl, a.B), hist = hist_bytes(a.iters, slots);
    if (!a.ws || a.ws_bytes < inbox + hist + gg_scratch_bytes(a.B, a.C, a.H, a.W)) return CSPN_ERR_WORKSPACE;
    const size_t hw = (size_t)a.H * a.W;
    if (a.Cg > 8) {
        // channels the forward never reads get exact zeros (SURVEY.md A.4 invariant 6)
        cudaError_t e = cudaMemset2DAsync(a.grad_guidance + 8 * hw, (size_t)a.Cg * hw * sizeof(T), 0, (size_t)(a.Cg - 8) * hw * sizeof(T), (size_t)a.B, a.stream);
        if (e != cudaSuccess) return (int)e;
    }
    FusedParams<T> p{};
    p.g = a.guidance; p.gbs = a.gbs; p.depth = a.depth; p.sparse = a.sparse; p.sparse_channels = a.sparse_channels; p.out = nullptr;
    p.C = 1; p.H = a.H; p.W = a.W; p.iters = a.iters;
    p.gout = a.grad_out; p.gd = a.grad_depth; p.Ctot = a.C;
    p.hist = (float*)((char*)a.ws + inbox); p.hist_slots = slots; p.hist_by_cta = hist_by_cta(tl, a.B, cap) ? 1 : 0;
    T* const scratch = (T*)((char*)a.ws + inbox + hist);
    // One launch per depth channel (the channels of an image share the affinity, pac.py:77-78,118-119): channel 0 writes
    // grad_guidance, every further channel writes an 8-channel scratch that is added on.
    for (int ch = 0; ch < a.C; ++ch) {
        p.ch0 = ch;
        p.gg = ch == 0 ? a.grad_guidance : scratch;
        p.Cg = ch == 0 ? a.Cg : 8;
        const int rc = a.mode == CSPN_MODE_NEW ? launch<T, kPBwd, kNWBwd, CSPN_MODE_NEW, true>(p, tl, a.B, inbox ? a.ws : nullptr, inbox, a.stream)
                                               : launch<T, kPBwd, kNWBwd, CSPN_MODE_OURS, true>(p, tl, a.B, inbox ? a.ws : nullptr, inbox, a.stream);
        if (rc != 0) return rc;
        if (ch > 0) {
            const size_t per_image = 8 * hw, total = (size_t)a.B * per_image;
            accumulate_gg_kernel<T><<<(unsigned)std::min<size_t>((total + 255) / 256, 148 * 8), 256, 0, a.stream>>>(a.grad_guidance, scratch, per_image, (size_t)a.Cg * hw, total);
            const cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return (int)e;
            ++call_stats().launches;
        }
    }
    return 0;
}

template int fused_backward<float>(const BwdArgs<float>&);
template int fused_backward<__half>(const BwdArgs<__half>&);

}  // namespace cspn
