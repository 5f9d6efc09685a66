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

// A batch that one wave of hardware clusters cannot hold (8 NYU images: only 7 clusters of 15 CTAs are co-resident on a
// B200) used to go to stream mode as a whole, which is ~30 % slower per image.  Split instead: as many images as fit run as
// hardware clusters, the remainder is streamed at the same time on the SMs the clusters leave free (second stream, forked and
// joined with events: graph-capturable).  Only offered when the remainder fits those SMs and no transport is forced.
// Measured on B200, 8 NYU images: 28.4 us all streamed, 27.2 us split.  (Streaming the remainder as 35 tiles of 64 x 40
// instead of 15 of 64 x 80 - shorter refresh periods - was built and measured too: 37.7 us, the 35-CTA cooperative launch no
// longer overlaps the clusters.)
struct SplitPlan { bool ok; int n_cluster; Tiling tc, ts; };

static Tiling stream_tiling(int H, int W, int th)
{
    const int step_y = th - 2 * kHaloY;
    Tiling t{}; t.stream = true; t.ntx = t.nty = 1;
    t.cx = W <= kTileW ? 1 : (W - kTileW + kStepX - 1) / kStepX + 1;
    t.cy = H <= th ? 1 : (H - th + step_y - 1) / step_y + 1;
    t.ew = kStepX * (t.cx - 1) + kTileW; t.eh = step_y * (t.cy - 1) + th;
    t.stepx = t.ew; t.stepy = t.eh;
    t.ctas = (long)t.cx * t.cy; t.ok = t.cx * t.cy > 1;
    return t;
}

SplitPlan plan_split(int B, int C, int H, int W, int iters)
{
    SplitPlan sp{}; sp.ok = false;
    static const bool off = [] { const char* v = getenv("CSPN_SPLIT"); return v && atoi(v) == 0; }();      // A/B knob
    if (off || C != 1 || B < 2 || iters > 60 || exchange_override() != 0) return sp;
    const Capacity cap = capacity<kPFwd, kNW, false>();
    const Tiling all = choose_tiling(H, W, iters, kTHBig, (long)B, cap);
    if (!all.ok || !all.stream) return sp;                                  // the planner already prefers clusters
    const Tiling one = choose_tiling(H, W, iters, kTHBig, 1, cap);
    if (!one.ok || one.stream || one.ntx * one.nty != 1) return sp;         // an image must be exactly one cluster
    const int held = cap.clusters[one.cx * one.cy];
    // exactly one image more than the clusters hold: with two streamed images beside 7 clusters the launches no longer
    // overlap well on this B200 (9 NYU images: 39.7 us split vs 28.9 us all streamed)
    if (held < 1 || B != held + 1) return sp;
    const Tiling ts = stream_tiling(H, W, kTHBig);
    const long free_sms = cap.sms - (long)held * one.cx * one.cy;
    if (!ts.ok || ts.ctas > free_sms) return sp;
    sp.ok = true; sp.n_cluster = held; sp.tc = one; sp.ts = ts;
    return sp;
}

// per host thread and device: the second stream and the fork / join events of a split launch
struct SplitCtx { bool ok; cudaStream_t aux; cudaEvent_t fork, join; };
SplitCtx* split_ctx()
{
    static thread_local SplitCtx ctxs[16] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) { cudaGetLastError(); return nullptr; }
    SplitCtx& c = ctxs[dev];
    if (!c.ok) {
        if (cudaStreamCreateWithFlags(&c.aux, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&c.fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&c.join, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        c.ok = true;
    }
    return &c;
}

// The single-tile kernel can run the problem: the planner (same inputs as the launch: real device capacity, all
// planes) finds a tiling and the plane count fits gridDim.z.
static bool single_ok(int B, int C, int H, int W, int iters)
{
    if ((long)H * W > (1l << 30) || (long)B * C > 65535) return false;
    return plan_forward(B, C, H, W, iters).ok;
}

bool fused_single_possible(int B, int C, int H, int W, int iters) { return single_ok(B, C, H, W, iters); }

bool fused_supported(int B, int C, int H, int W, int iters, int ksize, int mode)
{
    if (ksize != 3 || iters < 1 || B < 1) return false;
    return dual_supported(B, C, H, W, iters, ksize, mode) || single_ok(B, C, H, W, iters);
}

static size_t single_workspace(int B, int C, int H, int W, int iters)
{
    if (!single_ok(B, C, H, W, iters)) return 0;
    const SplitPlan sp = plan_split(B, C, H, W, iters);
    if (sp.ok) return kStatusBytes + (size_t)(sp.ts.ctas * (long)(B - sp.n_cluster)) * inbox_bytes<kTHBig>();
    const Tiling tl = plan_forward(B, C, H, W, iters);
    if (!tl.stream) return 0;
    return kStatusBytes + (size_t)(tl.ctas * (long)B * C) * inbox_bytes<kTHBig>();     // status word + inboxes of the global-memory exchange, one per tile
}

// Sized for whichever kernel the launch ends up with: the dual-slot kernel can still hand over to the single-tile one
// at launch time (guidance pointer / batch stride not TMA-addressable).
size_t fused_workspace(int B, int C, int H, int W, int iters, int ksize, int mode)
{
    const size_t a = dual_supported(B, C, H, W, iters, ksize, mode) ? dual_workspace(B, C, H, W, iters) : 0;
    const size_t b = single_workspace(B, C, H, W, iters);
    return a > b ? a : b;
}

template <typename T>
static int launch_mode(const FusedParams<T>& p, int mode, const Tiling& tl, int B, void* ws, size_t ws_bytes, cudaStream_t stream)
{
    return mode == CSPN_MODE_NEW ? launch<T, kPFwd, kNW, CSPN_MODE_NEW, false>(p, tl, B, ws, ws_bytes, stream)
                                 : launch<T, kPFwd, kNW, CSPN_MODE_OURS, false>(p, tl, B, ws, ws_bytes, stream);
}

template <typename T>
int fused_forward(const FwdArgs<T>& a)
{
    if (dual_supported(a.B, a.C, a.H, a.W, a.iters, a.ksize, a.mode)) {
        const int rc = dual_forward<T>(a);
        if (rc != kDualFallback) return rc;
    }
    if (!single_ok(a.B, a.C, a.H, a.W, a.iters)) return kDualFallback;       // nothing fused fits: the caller falls through
    FusedParams<T> p = forward_params(a);
    const SplitPlan sp = plan_split(a.B, a.C, a.H, a.W, a.iters);
    SplitCtx* sc = sp.ok ? split_ctx() : nullptr;
    if (sc) {
        const size_t hw = (size_t)a.H * a.W;
        const int n1 = sp.n_cluster, n2 = a.B - n1;
        cudaError_t e = cudaEventRecord(sc->fork, a.stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(sc->aux, sc->fork, 0);
        if (e != cudaSuccess) return (int)e;
        // clusters first: they need whole GPCs; the streamed remainder then takes SMs the clusters cannot use anyway
        FusedParams<T> p2 = p;
        p2.g = p.g + (size_t)n1 * p.gbs; p2.depth = p.depth + (size_t)n1 * hw; p2.out = p.out + (size_t)n1 * hw;
        if (p.sparse) p2.sparse = p.sparse + (size_t)n1 * p.sparse_channels * hw;
        int rc = launch_mode<T>(p, a.mode, sp.tc, n1, nullptr, 0, a.stream);
        const int rc2 = launch_mode<T>(p2, a.mode, sp.ts, n2, a.ws, a.ws_bytes, sc->aux);
        e = cudaEventRecord(sc->join, sc->aux);                              // join even after an error: the caller's stream must not run ahead of aux
        if (e == cudaSuccess) e = cudaStreamWaitEvent(a.stream, sc->join, 0);
        if (rc == 0) rc = rc2;
        return rc != 0 ? rc : (int)e;
    }
    const Tiling tl = plan_forward(a.B, a.C, a.H, a.W, a.iters);
    return launch_mode<T>(p, a.mode, tl, a.B, a.ws, a.ws_bytes, a.stream);
}

template int fused_forward<float>(const FwdArgs<float>&);
template int fused_forward<__half>(const FwdArgs<__half>&);

}  // namespace cspn
