// C ABI of the CSPN B200 library (include/cspn_b200.h): argument validation, path selection
// and the host-buffer convenience entry points.  No torch types anywhere in this library.
#include <atomic>
#include <cstdlib>
#include <cstring>

#include "cspn_common.cuh"

namespace cspn {

CallStats& call_stats()
{
    static thread_local CallStats s = {0, 0};
    return s;
}

namespace {

std::atomic<int> g_path{CSPN_PATH_AUTO};

int validate_common(const void* guidance, int64_t gbs, const void* depth, const void* sparse, int sparse_channels,
                    int B, int C, int H, int W, int iters, int ksize, int mode, TapTable* tt)
{
    if (mode != CSPN_MODE_NEW && mode != CSPN_MODE_OURS) return CSPN_ERR_BAD_MODE;
    if (!make_taps(mode, ksize, tt)) return CSPN_ERR_BAD_KERNEL_SIZE;
    if (B < 0 || C < 1 || H < 1 || W < 1 || iters < 0) return CSPN_ERR_BAD_SHAPE;
    if ((uint64_t)B * (uint64_t)C * (uint64_t)H * (uint64_t)W > (1ull << 40)) return CSPN_ERR_BAD_SHAPE;
    if (B == 0) return CSPN_OK;
    if (!guidance || !depth) return CSPN_ERR_NULL_POINTER;
    if (gbs < (int64_t)tt->n * H * W) return CSPN_ERR_BAD_STRIDE;
    if (sparse && sparse_channels != 1 && sparse_channels != C) return CSPN_ERR_BAD_SPARSE_CHANNELS;
    return CSPN_OK;
}

bool overlaps(const void* a, size_t an, const void* b, size_t bn)
{
    const char* pa = (const char*)a; const char* pb = (const char*)b;
    return a && b && pa < pb + bn && pb < pa + an;
}

bool use_fused(int B, int C, int H, int W, int iters, int ksize, int mode, int* err)
{
    const int path = g_path.load(std::memory_order_relaxed);
    const bool ok = fused_supported(B, C, H, W, iters, ksize, mode);
    if (path == CSPN_PATH_FUSED && !ok && err) *err = CSPN_ERR_BAD_KERNEL_SIZE;
    return path != CSPN_PATH_GENERIC && ok;
}

bool use_blocked(int B, int C, int H, int W, int iters, int ksize, int mode)
{
    return g_path.load(std::memory_order_relaxed) == CSPN_PATH_AUTO && blocked5x5_supported(B, C, H, W, iters, ksize, mode);
}

bool use_blocked_bwd(int B, int C, int H, int W, int iters, int ksize, int mode)
{
    return g_path.load(std::memory_order_relaxed) == CSPN_PATH_AUTO && blocked5x5_bwd_supported(B, C, H, W, iters, ksize, mode);
}

bool use_fused_bwd(int C, int H, int W, int iters, int ksize, int mode, int* err)
{
    const int path = g_path.load(std::memory_order_relaxed);
    const bool ok = fused_bwd_supported(C, H, W, iters, ksize, mode);
    if (path == CSPN_PATH_FUSED && !ok && err) *err = CSPN_ERR_BAD_KERNEL_SIZE;
    return path != CSPN_PATH_GENERIC && ok;
}

template <typename T>
int forward_impl(const T* guidance, int64_t gbs, const T* depth, const T* sparse, int sparse_channels, T* out,
                 int B, int C, int H, int W, int iters, int ksize, int mode, void* ws, size_t ws_bytes, void* stream)
{
    TapTable tt;
    int rc = validate_common(guidance, gbs, depth, sparse, sparse_channels, B, C, H, W, iters, ksize, mode, &tt);
    if (rc != CSPN_OK || B == 0) return rc;
    if (!out) return CSPN_ERR_NULL_POINTER;
    const size_t hw = (size_t)H * W, nout = (size_t)B * C * hw * sizeof(T);
    if (overlaps(out, nout, depth, nout) || overlaps(out, nout, guidance, (size_t)B * gbs * sizeof(T)) ||
        (sparse && overlaps(out, nout, sparse, (size_t)B * sparse_channels * hw * sizeof(T))))
        return CSPN_ERR_ALIAS;
    FwdArgs<T> a{guidance, gbs, depth, sparse, sparse ? sparse_channels : 1, out, B, C, H, W, iters, ksize, mode, ws, ws_bytes, (cudaStream_t)stream};
    call_stats().launches = 0;
    int err = CSPN_OK;
    if (iters > 0 && use_fused(B, C, H, W, iters, ksize, mode, &err)) {
        rc = fused_forward<T>(a);
        if (rc == CSPN_OK) call_stats().path = CSPN_PATH_FUSED;
        if (rc != kDualFallback) return rc;
        // no fused kernel took the problem after all (guidance not TMA-addressable and no single-tile plan)
        if (g_path.load(std::memory_order_relaxed) == CSPN_PATH_FUSED) return CSPN_ERR_BAD_KERNEL_SIZE;
    }
    if (err != CSPN_OK) return err;
    if (iters > 0 && use_blocked(B, C, H, W, iters, ksize, mode)) {
        rc = blocked5x5_forward<T>(a);
        if (rc == CSPN_OK) call_stats().path = CSPN_PATH_BLOCKED;
        return rc;
    }
    if (iters > 0) {
        const size_t need = generic_fwd_workspace(B, C, H, W, tt.n);
        if (!ws || ws_bytes < need) return CSPN_ERR_WORKSPACE;
    }
    rc = generic_forward<T>(a, tt);
    if (rc == CSPN_OK) call_stats().path = CSPN_PATH_GENERIC;
    return rc;
}

template <typename T>
int backward_impl(const T* grad_out, const T* guidance, int64_t gbs, int Cg, const T* depth, const T* sparse,
                  int sparse_channels, T* grad_guidance, T* grad_depth, int B, int C, int H, int W, int iters,
                  int ksize, int mode, void* ws, size_t ws_bytes, void* stream)
{
    TapTable tt;
    int rc = validate_common(guidance, gbs, depth, sparse, sparse_channels, B, C, H, W, iters, ksize, mode, &tt);
    if (rc != CSPN_OK || B == 0) return rc;
    if (!grad_out || !grad_guidance || !grad_depth) return CSPN_ERR_NULL_POINTER;
    if (Cg < tt.n) return CSPN_ERR_BAD_STRIDE;
    BwdArgs<T> a{grad_out, guidance, gbs, Cg, depth, sparse, sparse ? sparse_channels : 1, grad_guidance, grad_depth,
                 B, C, H, W, iters, ksize, mode, ws, ws_bytes, (cudaStream_t)stream};
    call_stats().launches = 0;
    int err = CSPN_OK;
    if (iters > 0 && use_fused_bwd(C, H, W, iters, ksize, mode, &err)) {
        rc = fused_backward<T>(a);
        if (rc == CSPN_OK) call_stats().path = CSPN_PATH_FUSED;
        return rc;
    }
    if (err != CSPN_OK) return err;
    if (iters > 0 && use_blocked_bwd(B, C, H, W, iters, ksize, mode)) {
        rc = blocked5x5_backward<T>(a);
        if (rc == CSPN_OK) call_stats().path = CSPN_PATH_BLOCKED;
        return rc;
    }
    if (iters > 0) {
        const size_t need = generic_bwd_workspace(B, C, H, W, iters, tt.n);
        if (!ws || ws_bytes < need) return CSPN_ERR_WORKSPACE;
    }
    rc = generic_backward<T>(a, tt);
    if (rc == CSPN_OK) call_stats().path = CSPN_PATH_GENERIC;
    return rc;
}

template <typename T>
int heads_fwd_impl(const T* x, const T* w1, const T* w2, T* out1, T* out2, int B, int Cin, int h, int w, int H, int W, int n1, int n2, void* stream)
{
    if (B < 0 || !heads_supported(n1, n2, Cin, h, w, H, W)) return CSPN_ERR_BAD_SHAPE;
    if (B == 0) return CSPN_OK;
    if (!x || !w1 || !out1 || (n2 > 0 && (!w2 || !out2))) return CSPN_ERR_NULL_POINTER;
    call_stats().launches = 0;
    return heads_forward<T>(x, w1, w2, out1, out2, n1, n2, B, Cin, h, w, H, W, (cudaStream_t)stream);
}

template <typename T>
int heads_bwd_impl(const T* x, const T* w1, const T* w2, const T* go1, const T* go2, T* gx, T* gw1, T* gw2, int B, int Cin, int h, int w, int H, int W,
                   int n1, int n2, void* ws, size_t ws_bytes, void* stream)
{
    if (B < 0 || !heads_supported(n1, n2, Cin, h, w, H, W)) return CSPN_ERR_BAD_SHAPE;
    if (B == 0) return CSPN_OK;
    if (!x || !w1 || !go1 || (n2 > 0 && (!w2 || !go2))) return CSPN_ERR_NULL_POINTER;
    if (gw1 && n2 > 0 && !gw2) return CSPN_ERR_NULL_POINTER;
    if (gw1 && (!ws || ws_bytes < heads_workspace_bytes() || ((uintptr_t)ws & 15))) return CSPN_ERR_WORKSPACE;
    call_stats().launches = 0;
    return heads_backward<T>(x, w1, w2, go1, go2, gx, gw1, gw2, n1, n2, B, Cin, h, w, H, W, ws, (cudaStream_t)stream);
}

template <typename T>
int legacy_impl(const T* guidance, int64_t gbs, const T* depth, const T* sparse, T* out, int B, int H, int W, int iters, void* ws, size_t ws_bytes,
                void* stream)
{
    if (B < 0 || H < 1 || W < 1 || iters < 1) return CSPN_ERR_BAD_SHAPE;
    if ((uint64_t)B * (uint64_t)H * (uint64_t)W > (1ull << 40)) return CSPN_ERR_BAD_SHAPE;
    if (B == 0) return CSPN_OK;
    if (!guidance || !depth || !out) return CSPN_ERR_NULL_POINTER;
    const size_t hw = (size_t)H * W, nout = (size_t)B * hw * sizeof(T);
    if (gbs < (int64_t)8 * (int64_t)hw) return CSPN_ERR_BAD_STRIDE;
    if (overlaps(out, nout, depth, nout) || overlaps(out, nout, guidance, (size_t)B * gbs * sizeof(T)) || (sparse && overlaps(out, nout, sparse, nout)))
        return CSPN_ERR_ALIAS;
    call_stats().launches = 0;
    const int rc = legacy_forward<T>(guidance, gbs, depth, sparse, out, B, H, W, iters, ws, ws_bytes, (cudaStream_t)stream);
    if (rc == CSPN_OK) call_stats().path = CSPN_PATH_BLOCKED;
    return rc;
}

// Per host thread and device state of the host entry points, created once:
//  * a private stream-ordered memory pool that KEEPS its memory (release threshold = max): with the default pool every
//    cudaStreamSynchronize hands the scratch back to the driver and the next call pays a fresh allocation (~0.3 ms);
//  * two helper streams so the small depth / sparse copies overlap the guidance copy instead of queueing behind it
//    (each small copy costs ~90 us on its own, tools/e2e_breakdown.py).
constexpr int kAuxLanes = 2;
struct HostCtx { bool ok; cudaMemPool_t pool; cudaStream_t s[kAuxLanes]; cudaEvent_t fork, join[kAuxLanes]; };
HostCtx* host_ctx()
{
    static thread_local HostCtx ctxs[16] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) { cudaGetLastError(); return nullptr; }
    HostCtx& c = ctxs[dev];
    if (!c.ok) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        bool good = cudaMemPoolCreate(&c.pool, &props) == cudaSuccess;
        if (good) {
            unsigned long long keep = ~0ull;
            good = cudaMemPoolSetAttribute(c.pool, cudaMemPoolAttrReleaseThreshold, &keep) == cudaSuccess;
        }
        good = good && cudaEventCreateWithFlags(&c.fork, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; good && i < kAuxLanes; ++i)
            good = cudaStreamCreateWithFlags(&c.s[i], cudaStreamNonBlocking) == cudaSuccess &&
                   cudaEventCreateWithFlags(&c.join[i], cudaEventDisableTiming) == cudaSuccess;
        if (!good) { cudaGetLastError(); return nullptr; }
        c.ok = true;
    }
    return &c;
}

// H2D -> forward -> D2H with stream-ordered scratch; returns once `out` holds the result.
template <typename T>
int forward_host_impl(const T* guidance, int64_t gbs, const T* depth, const T* sparse, int sparse_channels, T* out,
                      int B, int C, int H, int W, int iters, int ksize, int mode, void* stream_)
{
    TapTable tt;
    int rc = validate_common(guidance, gbs, depth, sparse, sparse_channels, B, C, H, W, iters, ksize, mode, &tt);
    if (rc != CSPN_OK || B == 0) return rc;
    if (!out) return CSPN_ERR_NULL_POINTER;
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t hw = (size_t)H * W;
    // Only the K*K-1 channels that are read travel over PCIe; the device copy is packed (batch stride taps*H*W).
    const size_t g_img = (size_t)tt.n * hw * sizeof(T);
    const size_t g_bytes = (size_t)B * g_img, d_bytes = (size_t)B * C * hw * sizeof(T);
    const size_t s_bytes = sparse ? (size_t)B * sparse_channels * hw * sizeof(T) : 0;
    const size_t ws_bytes = cspn_fwd_workspace_bytes(B, C, H, W, iters, ksize, mode);
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t total = up(g_bytes) + 2 * up(d_bytes) + up(s_bytes) + up(ws_bytes);
    char* base = nullptr;
    static const bool plain = [] { const char* v = getenv("CSPN_HOST_PLAIN"); return v && atoi(v) == 1; }();     // A/B knob: default pool, one stream
    HostCtx* ctx = plain ? nullptr : host_ctx();
    cudaError_t e = ctx ? cudaMallocFromPoolAsync((void**)&base, total, ctx->pool, stream) : cudaMallocAsync((void**)&base, total, stream);
    if (e != cudaSuccess) return (int)e;
    char* p = base;
    T* dg = (T*)p; p += up(g_bytes);
    T* dd = (T*)p; p += up(d_bytes);
    T* dout = (T*)p; p += up(d_bytes);
    T* ds = sparse ? (T*)p : nullptr; p += up(s_bytes);
    void* dws = ws_bytes ? (void*)p : nullptr;
    const bool fork = ctx && g_bytes >= ((size_t)1 << 20);
    cudaStream_t sd = fork ? ctx->s[0] : stream, ss = fork ? ctx->s[1] : stream;
    if (fork) {
        // the helper streams may touch the allocation only after it exists in `stream` order
        e = cudaEventRecord(ctx->fork, stream);
        for (int i = 0; e == cudaSuccess && i < kAuxLanes; ++i) e = cudaStreamWaitEvent(ctx->s[i], ctx->fork, 0);
    }
    if (e == cudaSuccess) {
        if (gbs == (int64_t)tt.n * (int64_t)hw) e = cudaMemcpyAsync(dg, guidance, g_bytes, cudaMemcpyHostToDevice, stream);
        else e = cudaMemcpy2DAsync(dg, g_img, guidance, (size_t)gbs * sizeof(T), g_img, (size_t)B, cudaMemcpyHostToDevice, stream);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(dd, depth, d_bytes, cudaMemcpyHostToDevice, sd);
    if (e == cudaSuccess && sparse) e = cudaMemcpyAsync(ds, sparse, s_bytes, cudaMemcpyHostToDevice, ss);
    if (fork)
        for (int i = 0; i < kAuxLanes; ++i) {
            // join even after an error: the free below is ordered on `stream`
            cudaError_t ej = cudaEventRecord(ctx->join[i], ctx->s[i]);
            if (ej == cudaSuccess) ej = cudaStreamWaitEvent(stream, ctx->join[i], 0);
            if (e == cudaSuccess) e = ej;
        }
    if (e == cudaSuccess) {
        rc = forward_impl<T>(dg, (int64_t)tt.n * (int64_t)hw, dd, ds, sparse_channels, dout, B, C, H, W, iters, ksize, mode, dws, ws_bytes, stream);
        if (rc == CSPN_OK) e = cudaMemcpyAsync(out, dout, d_bytes, cudaMemcpyDeviceToHost, stream);
    }
    cudaError_t e2 = cudaFreeAsync(base, stream);
    cudaError_t e3 = cudaStreamSynchronize(stream);
    if (rc != CSPN_OK) return rc;
    if (e != cudaSuccess) return (int)e;
    if (e2 != cudaSuccess) return (int)e2;
    return (int)e3;
}


// ---- pipelined host entry points --------------------------------------------------------------------------------------
// Per host thread and device: a ring of kPipeDepth slots, each with its own stream, persistent device buffers (grown on
// demand, never shrunk) and a pinned status word.  A submitted call runs H2D -> kernel -> D2H on its slot's stream, so the
// H2D copy of call i+1 overlaps the kernel and the D2H copy of call i (PCIe is full duplex, the copy engines are
// independent of the SMs).  The synchronous cspn_fwd_host_* are kept as they were (they honour the caller's stream).
constexpr int kPipeDepth = 3;
struct PipeSlot {
    cudaStream_t stream; cudaEvent_t done; char* dev; size_t cap; int* status_host; int rc; bool busy; int ticket;
};
struct HostPipe { bool ok; PipeSlot slot[kPipeDepth]; int next_ticket; };

HostPipe* host_pipe()
{
    static thread_local HostPipe pipes[16] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) { cudaGetLastError(); return nullptr; }
    HostPipe& hp = pipes[dev];
    if (!hp.ok) {
        bool good = true;
        for (int i = 0; good && i < kPipeDepth; ++i) {
            PipeSlot& sl = hp.slot[i];
            sl = PipeSlot{};
            good = cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking) == cudaSuccess &&
                   cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming | cudaEventBlockingSync) == cudaSuccess &&
                   cudaHostAlloc((void**)&sl.status_host, 64, cudaHostAllocDefault) == cudaSuccess;
        }
        if (!good) { cudaGetLastError(); return nullptr; }
        hp.next_ticket = 1;
        hp.ok = true;
    }
    return &hp;
}

int pipe_wait_slot(PipeSlot& sl)
{
    if (!sl.busy) return CSPN_OK;
    const cudaError_t e = cudaEventSynchronize(sl.done);
    sl.busy = false;
    if (sl.rc != CSPN_OK) return sl.rc;
    if (e != cudaSuccess) return (int)e;
    return *sl.status_host ? CSPN_ERR_EXCHANGE_TIMEOUT : CSPN_OK;
}

template <typename T>
int forward_host_submit_impl(const T* guidance, int64_t gbs, const T* depth, const T* sparse, int sparse_channels, T* out,
                             int B, int C, int H, int W, int iters, int ksize, int mode, int* ticket)
{
    if (ticket) *ticket = 0;
    TapTable tt;
    int rc = validate_common(guidance, gbs, depth, sparse, sparse_channels, B, C, H, W, iters, ksize, mode, &tt);
    if (rc != CSPN_OK) return rc;
    if (!ticket) return CSPN_ERR_NULL_POINTER;
    if (B == 0) return CSPN_OK;                       // ticket 0: nothing to wait for
    if (!out) return CSPN_ERR_NULL_POINTER;
    HostPipe* hp = host_pipe();
    if (!hp) return (int)cudaErrorInitializationError;
    const int tk = hp->next_ticket++;
    PipeSlot& sl = hp->slot[tk % kPipeDepth];
    pipe_wait_slot(sl);                               // back-pressure: the call that used this slot kPipeDepth submits ago (its status is dropped)
    const size_t hw = (size_t)H * W;
    const size_t g_img = (size_t)tt.n * hw * sizeof(T);
    const size_t g_bytes = (size_t)B * g_img, d_bytes = (size_t)B * C * hw * sizeof(T);
    const size_t s_bytes = sparse ? (size_t)B * sparse_channels * hw * sizeof(T) : 0;
    const size_t ws_bytes = cspn_fwd_workspace_bytes(B, C, H, W, iters, ksize, mode);
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t total = up(g_bytes) + 2 * up(d_bytes) + up(s_bytes) + up(ws_bytes) + 256;
    if (sl.cap < total) {
        if (sl.dev) cudaFree(sl.dev);
        sl.dev = nullptr; sl.cap = 0;
        const cudaError_t e = cudaMalloc((void**)&sl.dev, total);
        if (e != cudaSuccess) return (int)e;
        sl.cap = total;
    }
    char* q = sl.dev;
    T* dg = (T*)q; q += up(g_bytes);
    T* dd = (T*)q; q += up(d_bytes);
    T* dout = (T*)q; q += up(d_bytes);
    T* ds = sparse ? (T*)q : nullptr; q += up(s_bytes);
    void* dws = (void*)q;
    // fused plans that exchange halos through global memory (stream mode, dual-slot kernel) report a timeout through the
    // first int of their workspace; hardware-cluster plans need no workspace and cannot time out
    const bool has_status = iters > 0 && use_fused(B, C, H, W, iters, ksize, mode, nullptr) && fused_workspace(B, C, H, W, iters, ksize, mode) > 0;
    *sl.status_host = 0;
    cudaError_t e = cudaSuccess;
    {
        if (gbs == (int64_t)tt.n * (int64_t)hw) e = cudaMemcpyAsync(dg, guidance, g_bytes, cudaMemcpyHostToDevice, sl.stream);
        else e = cudaMemcpy2DAsync(dg, g_img, guidance, (size_t)gbs * sizeof(T), g_img, (size_t)B, cudaMemcpyHostToDevice, sl.stream);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(dd, depth, d_bytes, cudaMemcpyHostToDevice, sl.stream);
    if (e == cudaSuccess && sparse) e = cudaMemcpyAsync(ds, sparse, s_bytes, cudaMemcpyHostToDevice, sl.stream);
    rc = CSPN_OK;
    if (e == cudaSuccess) {
        rc = forward_impl<T>(dg, (int64_t)tt.n * (int64_t)hw, dd, ds, sparse_channels, dout, B, C, H, W, iters, ksize, mode, dws, ws_bytes + 256, sl.stream);
        if (rc == CSPN_OK) e = cudaMemcpyAsync(out, dout, d_bytes, cudaMemcpyDeviceToHost, sl.stream);
        if (rc == CSPN_OK && e == cudaSuccess && has_status) e = cudaMemcpyAsync(sl.status_host, dws, 4, cudaMemcpyDeviceToHost, sl.stream);
    }
    const cudaError_t e2 = cudaEventRecord(sl.done, sl.stream);
    sl.rc = rc != CSPN_OK ? rc : (e != cudaSuccess ? (int)e : (int)e2);
    sl.busy = true; sl.ticket = tk;
    *ticket = tk;
    return sl.rc;
}

}  // namespace
}  // namespace cspn

using namespace cspn;

extern "C" {

int cspn_abi_version(void) { return CSPN_B200_ABI_VERSION; }

const char* cspn_error_string(int code)
{
    switch (code) {
        case CSPN_OK: return "success";
        case CSPN_ERR_NULL_POINTER: return "cspn: required pointer is NULL";
        case CSPN_ERR_BAD_SHAPE: return "cspn: bad shape (need B >= 0, C,H,W >= 1, iters >= 0)";
        case CSPN_ERR_BAD_KERNEL_SIZE: return "cspn: unsupported kernel size for this mode/path (mode NEW needs 3; mode OURS needs odd 3..7)";
        case CSPN_ERR_BAD_MODE: return "cspn: unknown mode";
        case CSPN_ERR_BAD_STRIDE: return "cspn: guidance has fewer than K*K-1 channels (batch stride / Cg too small)";
        case CSPN_ERR_WORKSPACE: return "cspn: workspace missing or smaller than cspn_*_workspace_bytes()";
        case CSPN_ERR_BAD_SPARSE_CHANNELS: return "cspn: sparse must have 1 or C channels";
        case CSPN_ERR_ALIAS: return "cspn: out aliases an input";
        case CSPN_ERR_EXCHANGE_TIMEOUT: return "cspn: a tile never received its neighbours' halo (kernel not co-resident?); the output is NaN-filled";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "cspn: unknown error";
}

int cspn_set_path(int path)
{
    if (path < CSPN_PATH_AUTO || path > CSPN_PATH_FUSED) return g_path.load();
    return g_path.exchange(path);
}
int cspn_last_path(void) { return call_stats().path; }
int cspn_last_launch_count(void) { return call_stats().launches; }

size_t cspn_fwd_workspace_bytes(int B, int C, int H, int W, int iters, int ksize, int mode)
{
    TapTable tt;
    if (!make_taps(mode, ksize, &tt) || B < 1 || C < 1 || H < 1 || W < 1 || iters < 1) return 0;
    if (use_fused(B, C, H, W, iters, ksize, mode, nullptr)) {
        // a dual-slot-only problem may still fall through to the generic path at launch time (unaligned guidance pointer)
        const size_t f = fused_workspace(B, C, H, W, iters, ksize, mode);
        const size_t g = generic_fwd_workspace(B, C, H, W, tt.n);
        return fused_single_possible(B, C, H, W, iters, mode) ? f : (f > g ? f : g);
    }
    if (use_blocked(B, C, H, W, iters, ksize, mode)) return blocked5x5_workspace(B, C, H, W, iters);
    return generic_fwd_workspace(B, C, H, W, tt.n);
}

int cspn_fwd_plan(int B, int C, int H, int W, int iters, int ksize, int mode, int* plan10)
{
    TapTable tt;
    if (!plan10) return CSPN_ERR_NULL_POINTER;
    for (int i = 0; i < 10; ++i) plan10[i] = 0;
    if (!make_taps(mode, ksize, &tt)) return CSPN_ERR_BAD_KERNEL_SIZE;
    if (B < 1 || C < 1 || H < 1 || W < 1 || iters < 1) return CSPN_ERR_BAD_SHAPE;
    if (use_fused(B, C, H, W, iters, ksize, mode, nullptr)) {
        if (dual_supported(B, C, H, W, iters, ksize, mode)) { plan10[0] = CSPN_KERNEL_DUAL; dual_describe(B, C, H, W, iters, plan10 + 1); }
        else { plan10[0] = CSPN_KERNEL_SINGLE; single_describe(B, C, H, W, iters, mode, plan10 + 1); }
    } else if (use_blocked(B, C, H, W, iters, ksize, mode)) plan10[0] = CSPN_KERNEL_BLOCKED;
    else plan10[0] = CSPN_KERNEL_GENERIC;
    return CSPN_OK;
}

size_t cspn_bwd_workspace_bytes(int B, int C, int H, int W, int iters, int ksize, int mode)
{
    TapTable tt;
    if (!make_taps(mode, ksize, &tt) || B < 1 || C < 1 || H < 1 || W < 1 || iters < 1) return 0;
    if (use_fused_bwd(C, H, W, iters, ksize, mode, nullptr)) return fused_bwd_workspace(B, C, H, W, iters);
    if (use_blocked_bwd(B, C, H, W, iters, ksize, mode)) return blocked5x5_bwd_workspace(B, C, H, W, iters);
    return generic_bwd_workspace(B, C, H, W, iters, tt.n);
}

int cspn_fwd_f32(const float* guidance, int64_t gbs, const float* depth, const float* sparse, int sparse_channels,
                 float* out, int B, int C, int H, int W, int iters, int ksize, int mode, void* ws, size_t ws_bytes, void* stream)
{
    return forward_impl<float>(guidance, gbs, depth, sparse, sparse_channels, out, B, C, H, W, iters, ksize, mode, ws, ws_bytes, stream);
}
int cspn_fwd_f16(const void* guidance, int64_t gbs, const void* depth, const void* sparse, int sparse_channels,
                 void* out, int B, int C, int H, int W, int iters, int ksize, int mode, void* ws, size_t ws_bytes, void* stream)
{
    return forward_impl<__half>((const __half*)guidance, gbs, (const __half*)depth, (const __half*)sparse, sparse_channels,
                                (__half*)out, B, C, H, W, iters, ksize, mode, ws, ws_bytes, stream);
}
int cspn_bwd_f32(const float* grad_out, const float* guidance, int64_t gbs, int Cg, const float* depth, const float* sparse,
                 int sparse_channels, float* grad_guidance, float* grad_depth, int B, int C, int H, int W, int iters,
                 int ksize, int mode, void* ws, size_t ws_bytes, void* stream)
{
    return backward_impl<float>(grad_out, guidance, gbs, Cg, depth, sparse, sparse_channels, grad_guidance, grad_depth,
                                B, C, H, W, iters, ksize, mode, ws, ws_bytes, stream);
}
int cspn_bwd_f16(const void* grad_out, const void* guidance, int64_t gbs, int Cg, const void* depth, const void* sparse,
                 int sparse_channels, void* grad_guidance, void* grad_depth, int B, int C, int H, int W, int iters,
                 int ksize, int mode, void* ws, size_t ws_bytes, void* stream)
{
    return backward_impl<__half>((const __half*)grad_out, (const __half*)guidance, gbs, Cg, (const __half*)depth,
                                 (const __half*)sparse, sparse_channels, (__half*)grad_guidance, (__half*)grad_depth,
                                 B, C, H, W, iters, ksize, mode, ws, ws_bytes, stream);
}
int cspn_fwd_host_f32(const float* guidance, int64_t gbs, const float* depth, const float* sparse, int sparse_channels,
                      float* out, int B, int C, int H, int W, int iters, int ksize, int mode, void* stream)
{
    return forward_host_impl<float>(guidance, gbs, depth, sparse, sparse_channels, out, B, C, H, W, iters, ksize, mode, stream);
}
static int abn_shape_ok(int N, int C, int S, int act)
{
    return N >= 1 && C >= 1 && S >= 1 && (long)N * C <= 2147483647l && act >= 0 && act <= 2;
}
size_t cspn_abn_workspace_bytes(int C) { return C < 1 ? 0 : abn_workspace_bytes(C); }
int cspn_abn_stats_f32(const float* x, int N, int C, int S, double* sums, void* ws, size_t ws_bytes, void* stream)
{
    if (!abn_shape_ok(N, C, S, 0)) return CSPN_ERR_BAD_SHAPE;
    if (!x || !sums) return CSPN_ERR_NULL_POINTER;
    if (!ws || ws_bytes < abn_workspace_bytes(C) || ((uintptr_t)ws & 7)) return CSPN_ERR_WORKSPACE;
    call_stats().launches = 0;
    return abn_stats(x, N, C, S, sums, ws, (cudaStream_t)stream);
}
int cspn_abn_finalize_f32(const double* sums, double count, float* mean, float* var, float* running_mean, float* running_var, float momentum, int C,
                          void* stream)
{
    if (C < 1 || !(count >= 1.0)) return CSPN_ERR_BAD_SHAPE;
    if (!sums || !mean || !var) return CSPN_ERR_NULL_POINTER;
    call_stats().launches = 0;
    return abn_finalize(sums, count, mean, var, running_mean, running_var, momentum, C, (cudaStream_t)stream);
}
int cspn_abn_forward_f32(float* x, const float* mean, const float* var, const float* weight, const float* bias, int N, int C, int S, float eps,
                         int activation, float slope, void* stream)
{
    if (!abn_shape_ok(N, C, S, activation)) return CSPN_ERR_BAD_SHAPE;
    if (!x || !mean || !var) return CSPN_ERR_NULL_POINTER;
    call_stats().launches = 0;
    return abn_forward(x, mean, var, weight, bias, N, C, S, eps, activation, slope, (cudaStream_t)stream);
}
int cspn_abn_bwd_reduce_f32(const float* z, const float* dz, const float* weight, const float* bias, int N, int C, int S, float eps, int activation,
                            float slope, double* sums, void* ws, size_t ws_bytes, void* stream)
{
    if (!abn_shape_ok(N, C, S, activation)) return CSPN_ERR_BAD_SHAPE;
    if (!z || !dz || !sums) return CSPN_ERR_NULL_POINTER;
    if (!ws || ws_bytes < abn_workspace_bytes(C) || ((uintptr_t)ws & 7)) return CSPN_ERR_WORKSPACE;
    call_stats().launches = 0;
    return abn_bwd_reduce(z, dz, weight, bias, N, C, S, eps, activation, slope, sums, ws, (cudaStream_t)stream);
}
int cspn_abn_bwd_apply_f32(const float* z, const float* dz, float* dx, const float* var, const float* weight, const float* bias, const double* sums,
                           double count_total, double count_local, float* dweight, float* dbias, int N, int C, int S, float eps, int activation,
                           float slope, void* stream)
{
    if (!abn_shape_ok(N, C, S, activation) || !(count_total >= 1.0)) return CSPN_ERR_BAD_SHAPE;
    if (!z || !dz || !var || (dweight && !weight)) return CSPN_ERR_NULL_POINTER;
    call_stats().launches = 0;
    return abn_bwd_apply(z, dz, dx, var, weight, bias, sums, count_total, count_local, dweight, dbias, N, C, S, eps, activation, slope, (cudaStream_t)stream);
}
size_t cspn_heads_workspace_bytes(void) { return heads_workspace_bytes(); }
int cspn_heads_fwd_f32(const float* x, const float* w1, const float* w2, float* out1, float* out2, int B, int Cin, int h, int w, int H, int W, int n1, int n2,
                       void* stream)
{
    return heads_fwd_impl<float>(x, w1, w2, out1, out2, B, Cin, h, w, H, W, n1, n2, stream);
}
int cspn_heads_fwd_f16(const void* x, const void* w1, const void* w2, void* out1, void* out2, int B, int Cin, int h, int w, int H, int W, int n1, int n2,
                       void* stream)
{
    return heads_fwd_impl<__half>((const __half*)x, (const __half*)w1, (const __half*)w2, (__half*)out1, (__half*)out2, B, Cin, h, w, H, W, n1, n2, stream);
}
int cspn_heads_bwd_f32(const float* x, const float* w1, const float* w2, const float* go1, const float* go2, float* gx, float* gw1, float* gw2, int B, int Cin,
                       int h, int w, int H, int W, int n1, int n2, void* ws, size_t ws_bytes, void* stream)
{
    return heads_bwd_impl<float>(x, w1, w2, go1, go2, gx, gw1, gw2, B, Cin, h, w, H, W, n1, n2, ws, ws_bytes, stream);
}
int cspn_heads_bwd_f16(const void* x, const void* w1, const void* w2, const void* go1, const void* go2, void* gx, void* gw1, void* gw2, int B, int Cin,
                       int h, int w, int H, int W, int n1, int n2, void* ws, size_t ws_bytes, void* stream)
{
    return heads_bwd_impl<__half>((const __half*)x, (const __half*)w1, (const __half*)w2, (const __half*)go1, (const __half*)go2, (__half*)gx, (__half*)gw1,
                                  (__half*)gw2, B, Cin, h, w, H, W, n1, n2, ws, ws_bytes, stream);
}
size_t cspn_legacy_workspace_bytes(int B, int H, int W, int iters)
{
    if (B < 1 || H < 1 || W < 1 || iters < 1) return 0;
    return legacy_workspace(B, H, W, iters);
}
int cspn_legacy_fwd_f32(const float* guidance, int64_t gbs, const float* depth, const float* sparse, float* out, int B, int H, int W, int iters,
                        void* ws, size_t ws_bytes, void* stream)
{
    return legacy_impl<float>(guidance, gbs, depth, sparse, out, B, H, W, iters, ws, ws_bytes, stream);
}
int cspn_legacy_fwd_f16(const void* guidance, int64_t gbs, const void* depth, const void* sparse, void* out, int B, int H, int W, int iters,
                        void* ws, size_t ws_bytes, void* stream)
{
    return legacy_impl<__half>((const __half*)guidance, gbs, (const __half*)depth, (const __half*)sparse, (__half*)out, B, H, W, iters, ws, ws_bytes, stream);
}
size_t cspn_loss_workspace_bytes(void) { return loss_workspace_bytes(); }
int cspn_masked_l1_fwd_f32(const float* pred, const float* target, int64_t n, float* loss2, void* ws, size_t ws_bytes, void* stream)
{
    if (n < 0) return CSPN_ERR_BAD_SHAPE;
    call_stats().launches = 0;
    return masked_l1_forward<float>(pred, target, (size_t)n, loss2, ws, ws_bytes, (cudaStream_t)stream);
}
int cspn_masked_l1_fwd_f16(const void* pred, const void* target, int64_t n, float* loss2, void* ws, size_t ws_bytes, void* stream)
{
    if (n < 0) return CSPN_ERR_BAD_SHAPE;
    call_stats().launches = 0;
    return masked_l1_forward<__half>((const __half*)pred, (const __half*)target, (size_t)n, loss2, ws, ws_bytes, (cudaStream_t)stream);
}
int cspn_masked_l1_bwd_f32(const float* pred, const float* target, int64_t n, const float* loss2, const float* grad_loss, float* grad_pred, void* stream)
{
    if (n < 0) return CSPN_ERR_BAD_SHAPE;
    call_stats().launches = 0;
    return masked_l1_backward<float>(pred, target, (size_t)n, loss2, grad_loss, grad_pred, (cudaStream_t)stream);
}
int cspn_masked_l1_bwd_f16(const void* pred, const void* target, int64_t n, const float* loss2, const float* grad_loss, void* grad_pred, void* stream)
{
    if (n < 0) return CSPN_ERR_BAD_SHAPE;
    call_stats().launches = 0;
    return masked_l1_backward<__half>((const __half*)pred, (const __half*)target, (size_t)n, loss2, grad_loss, (__half*)grad_pred, (cudaStream_t)stream);
}
int cspn_depth_metrics_f32(const float* pred, const float* target, int64_t n, float* out11, void* ws, size_t ws_bytes, void* stream)
{
    if (n < 0) return CSPN_ERR_BAD_SHAPE;
    call_stats().launches = 0;
    return depth_metrics<float>(pred, target, (size_t)n, out11, ws, ws_bytes, (cudaStream_t)stream);
}
int cspn_depth_metrics_f16(const void* pred, const void* target, int64_t n, float* out11, void* ws, size_t ws_bytes, void* stream)
{
    if (n < 0) return CSPN_ERR_BAD_SHAPE;
    call_stats().launches = 0;
    return depth_metrics<__half>((const __half*)pred, (const __half*)target, (size_t)n, out11, ws, ws_bytes, (cudaStream_t)stream);
}

int cspn_fwd_host_submit_f32(const float* guidance, int64_t gbs, const float* depth, const float* sparse, int sparse_channels,
                             float* out, int B, int C, int H, int W, int iters, int ksize, int mode, int* ticket)
{
    return forward_host_submit_impl<float>(guidance, gbs, depth, sparse, sparse_channels, out, B, C, H, W, iters, ksize, mode, ticket);
}
int cspn_fwd_host_submit_f16(const void* guidance, int64_t gbs, const void* depth, const void* sparse, int sparse_channels,
                             void* out, int B, int C, int H, int W, int iters, int ksize, int mode, int* ticket)
{
    return forward_host_submit_impl<__half>((const __half*)guidance, gbs, (const __half*)depth, (const __half*)sparse, sparse_channels,
                                            (__half*)out, B, C, H, W, iters, ksize, mode, ticket);
}
int cspn_host_wait(int ticket)
{
    if (ticket <= 0) return CSPN_OK;
    HostPipe* hp = host_pipe();
    if (!hp) return (int)cudaErrorInitializationError;
    PipeSlot& sl = hp->slot[ticket % kPipeDepth];
    if (!sl.busy || sl.ticket != ticket) return CSPN_OK;      // already waited for (or retired by a later submit)
    return pipe_wait_slot(sl);
}
int cspn_host_pipeline_depth(void) { return kPipeDepth; }

int cspn_fwd_host_f16(const void* guidance, int64_t gbs, const void* depth, const void* sparse, int sparse_channels,
                      void* out, int B, int C, int H, int W, int iters, int ksize, int mode, void* stream)
{
    return forward_host_impl<__half>((const __half*)guidance, gbs, (const __half*)depth, (const __half*)sparse, sparse_channels,
                                     (__half*)out, B, C, H, W, iters, ksize, mode, stream);
}

}  // extern "C"
