// Backward instantiations and host entry points of the fused 3x3 CSPN kernel (cspn_fused3x3.cuh, BWD = true):
// forward recompute with history + reverse sweep + normalisation / abs / softmax Jacobian in ONE launch.
// Replaces autograd through CSPN_new.py:80-90 and Conv2dFn.backward (pac.py:96-121) x T.
#include "cspn_fused3x3.cuh"

namespace cspn {

namespace {
constexpr int kTHBwd = kNWBwd * kPBwd;
inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }
inline size_t hist_bytes(int iters) { return (size_t)kHistSlots * (size_t)iters * kTHBwd * kTileW * sizeof(float); }
inline size_t bwd_inbox_bytes(const Tiling& tl, int B)
{
    if (!tl.stream) return 0;
    return up256((size_t)(tl.ctas * (long)B) * inbox_bytes<kTHBwd>());
}
}  // namespace

bool fused_bwd_supported(int C, int H, int W, int iters, int ksize, int mode)
{
    (void)mode;
    // one depth channel (the gradient of the shared affinity would otherwise sum over channels), refresh tags fit 7 bits
    if (C != 1 || ksize != 3 || iters < 1 || iters > 60) return false;
    if ((long)H * W > (1l << 30)) return false;
    return choose_tiling(H, W, iters, kTHBwd, 1, default_capacity()).ok;
}

size_t fused_bwd_workspace(int B, int C, int H, int W, int iters)
{
    (void)C;
    const Tiling tl = choose_tiling(H, W, iters, kTHBwd, (long)B, capacity<kPBwd, kNWBwd, true>());
    if (!tl.ok) return 0;
    return bwd_inbox_bytes(tl, B) + hist_bytes(iters);
}

template <typename T>
int fused_backward(const BwdArgs<T>& a)
{
    const Tiling tl = choose_tiling(a.H, a.W, a.iters, kTHBwd, (long)a.B, capacity<kPBwd, kNWBwd, true>());
    if (!tl.ok || a.C != 1) return CSPN_ERR_BAD_KERNEL_SIZE;
    if ((long)a.B > 65535) return CSPN_ERR_BAD_SHAPE;
    const size_t inbox = bwd_inbox_bytes(tl, a.B);
    if (!a.ws || a.ws_bytes < inbox + hist_bytes(a.iters)) return CSPN_ERR_WORKSPACE;
    const size_t hw = (size_t)a.H * a.W;
    if (a.Cg > 8) {
        // channels the forward never reads get exact zeros (SURVEY.md A.4 invariant 6)
        cudaError_t e = cudaMemset2DAsync(a.grad_guidance + 8 * hw, (size_t)a.Cg * hw * sizeof(T), 0, (size_t)(a.Cg - 8) * hw * sizeof(T), (size_t)a.B, a.stream);
        if (e != cudaSuccess) return (int)e;
    }
    FusedParams<T> p{};
    p.g = a.guidance; p.gbs = a.gbs; p.depth = a.depth; p.sparse = a.sparse; p.sparse_channels = a.sparse_channels; p.out = nullptr;
    p.C = 1; p.H = a.H; p.W = a.W; p.iters = a.iters;
    p.gout = a.grad_out; p.gg = a.grad_guidance; p.gd = a.grad_depth; p.Cg = a.Cg;
    p.hist = (float*)((char*)a.ws + inbox); p.hist_slots = kHistSlots;
    return a.mode == CSPN_MODE_NEW ? launch<T, kPBwd, kNWBwd, CSPN_MODE_NEW, true>(p, tl, a.B, inbox ? a.ws : nullptr, inbox, a.stream)
                                   : launch<T, kPBwd, kNWBwd, CSPN_MODE_OURS, true>(p, tl, a.B, inbox ? a.ws : nullptr, inbox, a.stream);
}

template int fused_backward<float>(const BwdArgs<float>&);
template int fused_backward<__half>(const BwdArgs<__half>&);

}  // namespace cspn
