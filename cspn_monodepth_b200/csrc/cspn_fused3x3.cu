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

Tiling plan_forward(int B, int C, int H, int W, int iters, int mode)
{
    // hybrid transport: by cost for mode OURS only when CSPN_HYBRID_AUTO=1 (measured 7 % faster there), otherwise opt-in through
    // CSPN_EXCHANGE=hybrid.  Not on by default: Nsight Compute cannot replay cooperative + cluster launches (LaunchFailed / illegal
    // instruction under ncu 2025.2, fine under compute-sanitizer and in normal runs) - a default path must stay profilable.
    static const bool hyb_auto = [] { const char* v = getenv("CSPN_HYBRID_AUTO"); return v && v[0] == '1'; }();
    return choose_tiling(H, W, iters, kTHBig, (long)B * C, capacity<kPFwd, kNW, false>(), (hyb_auto && mode == CSPN_MODE_OURS) ? 2 : 1);
}
}  // namespace

// The single-tile kernel can run the problem: the planner (same inputs as the launch: real device capacity, all
// planes) finds a tiling and the plane count fits gridDim.z.
static bool single_ok(int B, int C, int H, int W, int iters, int mode)
{
    if ((long)H * W > (1l << 30) || (long)B * C > 65535) return false;
    return plan_forward(B, C, H, W, iters, mode).ok;
}

bool fused_single_possible(int B, int C, int H, int W, int iters, int mode) { return single_ok(B, C, H, W, iters, mode); }

void single_describe(int B, int C, int H, int W, int iters, int mode, int* out9)
{
    for (int i = 0; i < 9; ++i) out9[i] = 0;
    if (!single_ok(B, C, H, W, iters, mode)) return;
    const Tiling tl = plan_forward(B, C, H, W, iters, mode);
    out9[0] = tl.hyb ? CSPN_TRANSPORT_HYBRID : (tl.stream ? CSPN_TRANSPORT_STREAM : CSPN_TRANSPORT_CLUSTER);
    out9[1] = tl.cx; out9[2] = tl.cy; out9[3] = tl.ntx; out9[4] = tl.nty; out9[5] = (int)tl.ctas;
}

bool fused_supported(int B, int C, int H, int W, int iters, int ksize, int mode)
{
    if (ksize != 3 || iters < 1 || B < 1) return false;
    return dual_supported(B, C, H, W, iters, ksize, mode) || single_ok(B, C, H, W, iters, mode);
}

static size_t single_workspace(int B, int C, int H, int W, int iters, int mode)
{
    if (!single_ok(B, C, H, W, iters, mode)) return 0;
    const Tiling tl = plan_forward(B, C, H, W, iters, mode);
    if (!tl.stream && !tl.hyb) return 0;
    return kStatusBytes + (size_t)(tl.ctas * (long)B * C) * inbox_bytes<kTHBig>();     // status word + inboxes of the global-memory exchange, one per tile
}

// Sized for whichever kernel the launch ends up with: the dual-slot kernel can still hand over to the single-tile one
// at launch time (guidance pointer / batch stride not TMA-addressable).
size_t fused_workspace(int B, int C, int H, int W, int iters, int ksize, int mode)
{
    const size_t a = dual_supported(B, C, H, W, iters, ksize, mode) ? dual_workspace(B, C, H, W, iters) : 0;
    const size_t b = single_workspace(B, C, H, W, iters, mode);
    return a > b ? a : b;
}

template <typename T>
int fused_forward(const FwdArgs<T>& a)
{
    if (dual_supported(a.B, a.C, a.H, a.W, a.iters, a.ksize, a.mode)) {
        const int rc = dual_forward<T>(a);
        if (rc != kDualFallback) return rc;
    }
    if (!single_ok(a.B, a.C, a.H, a.W, a.iters, a.mode)) return kDualFallback;       // nothing fused fits: the caller falls through
    const Tiling tl = plan_forward(a.B, a.C, a.H, a.W, a.iters, a.mode);
    FusedParams<T> p = forward_params(a);
    return a.mode == CSPN_MODE_NEW ? launch<T, kPFwd, kNW, CSPN_MODE_NEW, false>(p, tl, a.B, a.ws, a.ws_bytes, a.stream)
                                   : launch<T, kPFwd, kNW, CSPN_MODE_OURS, false>(p, tl, a.B, a.ws, a.ws_bytes, a.stream);
}

template int fused_forward<float>(const FwdArgs<float>&);
template int fused_forward<__half>(const FwdArgs<__half>&);

}  // namespace cspn
