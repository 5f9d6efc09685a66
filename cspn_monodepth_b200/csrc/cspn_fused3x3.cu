// Forward host entry points of the fused 3x3 CSPN kernel (cspn_fused3x3.cuh) and its forward instantiations
// (64 x 80 pixel register tile, 8 warps, one CTA per SM).
#include "cspn_fused3x3.cuh"

namespace cspn {

#ifdef CSPN_TRACE
extern "C" __attribute__((visibility("default"))) int cspn_debug_set_trace(void* buf)
{
    long long* ptr = (long long*)buf;
    return (int)cudaMemcpyToSymbol(g_trace, &ptr, sizeof ptr);
}
#endif

namespace {
constexpr int kTHBig = kNW * kPFwd;

Tiling plan_forward(int B, int C, int H, int W, int iters)
{
    return choose_tiling(H, W, iters, kTHBig, (long)B * C, capacity<kPFwd, kNW, false>());
}
}  // namespace

bool fused_supported(int C, int H, int W, int iters, int ksize, int mode)
{
    (void)C; (void)mode;
    if (ksize != 3 || iters < 1) return false;
    if ((long)H * W > (1l << 30)) return false;
    return choose_tiling(H, W, iters, kTHBig, 1, default_capacity()).ok;
}

size_t fused_workspace(int B, int C, int H, int W, int iters)
{
    const Tiling tl = plan_forward(B, C, H, W, iters);
    if (!tl.ok || !tl.stream) return 0;
    return (size_t)(tl.ctas * (long)B * C) * inbox_bytes<kTHBig>();     // inboxes of the global-memory exchange, one per tile
}

template <typename T>
int fused_forward(const FwdArgs<T>& a)
{
    const Tiling tl = plan_forward(a.B, a.C, a.H, a.W, a.iters);
    if (!tl.ok) return CSPN_ERR_BAD_KERNEL_SIZE;
    if ((long)a.B * a.C > 65535) return CSPN_ERR_BAD_SHAPE;
    FusedParams<T> p = forward_params(a);
    return a.mode == CSPN_MODE_NEW ? launch<T, kPFwd, kNW, CSPN_MODE_NEW, false>(p, tl, a.B, a.ws, a.ws_bytes, a.stream)
                                   : launch<T, kPFwd, kNW, CSPN_MODE_OURS, false>(p, tl, a.B, a.ws, a.ws_bytes, a.stream);
}

template int fused_forward<float>(const FwdArgs<float>&);
template int fused_forward<__half>(const FwdArgs<__half>&);

}  // namespace cspn
