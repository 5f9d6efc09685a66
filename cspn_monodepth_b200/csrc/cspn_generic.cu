// Generic CSPN path: one launch per propagation step, any tap table (3x3 / 5x5 / 7x7),
// any image shape.  It is the reference-shaped schedule (CSPN_new.py:80-90 /
// CSPN_ours.py:47-53: T dependent sweeps) with each sweep collapsed into one kernel and the
// loop-invariant normalisation hoisted into a prologue.  The fused single-launch kernel in
// cspn_fused3x3.cu is the fast path; this file covers the configurations it does not
// (5x5 "PAC variant", ...), the backward pass of round 1, and serves as an on-device
// cross-check.
//
// HBM layout of the fp32 workspace (forward):
//   nw   [B][taps][H*W]   normalised tap weights n_k(p), weight located at the CENTRE pixel
//   r0,r1[B][C][H*W]      ping-pong depth planes
// Algorithmic HBM traffic per pixel per sweep: taps + 1 (r) + 1 (d0) + 1 (sparse) reads, 1 write.
#include "cspn_common.cuh"

namespace cspn {

namespace {

constexpr int kThreads = 256;

inline unsigned blocks_for(size_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }

// n_k(p) for one image.  Mode NEW: |g_k(p+o_k)| / S(p), S over in-bounds neighbours, 0/0 = NaN
// kept (CSPN_new.py:29-70,124,127).  Mode OURS: softmax over the taps at p (CSPN_ours.py:35).
template <typename T>
__global__ void prep_weights_kernel(const T* __restrict__ g, int64_t gbs, float* __restrict__ nw,
                                    float* __restrict__ ssum, int B, int H, int W, int mode, const __grid_constant__ TapTable tt)
{
    const size_t hw = (size_t)H * W;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)B * hw) return;
    const int b = (int)(idx / hw);
    const size_t p = idx - (size_t)b * hw;
    const int y = (int)(p / W), x = (int)(p - (size_t)y * W);
    const T* gb = g + (size_t)b * gbs;
    float* nb = nw + (size_t)b * tt.n * hw;
    if (mode == CSPN_MODE_NEW) {
        float wk[8];
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int yy = y + tt.dy[k], xx = x + tt.dx[k];
            const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
            wk[k] = in ? fabsf(to_f32(gb[k * hw + (size_t)yy * W + xx])) : 0.f;
            s += wk[k];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) nb[k * hw + p] = wk[k] / s;
        if (ssum) ssum[idx] = s;
    } else {
        float mx = -INFINITY;
        for (int k = 0; k < tt.n; ++k) mx = fmaxf(mx, to_f32(gb[k * hw + p]));
        float s = 0.f;
        for (int k = 0; k < tt.n; ++k) s += expf(to_f32(gb[k * hw + p]) - mx);
        for (int k = 0; k < tt.n; ++k) nb[k * hw + p] = expf(to_f32(gb[k * hw + p]) - mx) / s;
    }
}

// One propagation step: rout(p) = (1-m)*sum_k n_k(p)*rin(p+o_k) + m*d0(p).
// TIn: type of rin (T for the first sweep, which reads the depth input directly, else float).
template <typename T, typename TIn, typename TOut>
__global__ void sweep_kernel(const float* __restrict__ nw, const TIn* __restrict__ rin,
                             const T* __restrict__ d0, const T* __restrict__ sparse, int sparse_channels,
                             TOut* __restrict__ rout, int B, int C, int H, int W, const __grid_constant__ TapTable tt)
{
    const size_t hw = (size_t)H * W;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)B * C * hw) return;
    const int bc = (int)(idx / hw);
    const int b = bc / C, c = bc - b * C;
    const size_t p = idx - (size_t)bc * hw;
    const int y = (int)(p / W), x = (int)(p - (size_t)y * W);
    const float* nb = nw + (size_t)b * tt.n * hw + p;
    const TIn* rp = rin + (size_t)bc * hw;
    float acc = 0.f;
    for (int k = 0; k < tt.n; ++k) {
        const int yy = y + tt.dy[k], xx = x + tt.dx[k];
        // out-of-image taps read the zero padding but are still multiplied: a pixel whose gathered weights
        // are all zero has n_k = 0/0 and must come out NaN like the reference's 0/0 (CSPN_new.py:127)
        const float rv = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? to_f32(rp[(size_t)yy * W + xx]) : 0.f;
        acc = fmaf(nb[k * hw], rv, acc);
    }
    if (sparse) {
        const float m = signf(to_f32(sparse[((size_t)b * sparse_channels + (sparse_channels == 1 ? 0 : c)) * hw + p]));
        acc = (1.f - m) * acc + m * to_f32(d0[idx]);      // arithmetic blend, not a select (CSPN_new.py:90)
    }
    rout[idx] = from_f32<TOut>(acc);
}

template <typename T>
__global__ void copy_kernel(const T* __restrict__ in, T* __restrict__ out, size_t n)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) out[idx] = in[idx];
}

// ---- backward ---------------------------------------------------------------------------
// One reverse step for all channels of pixel p = (b, y, x):
//   u(p)      = (1-m(p)) * G(p)
//   gd0(p)   += m(p) * G(p)
//   gn_k(p)  += sum_c u_c(p) * r_c^t(p+o_k)
//   Gnext(p)  = sum_k u(p-o_k) * n_k(p-o_k)            (gather with reversed offsets)
template <typename T, typename TR, typename TG>
__global__ void bwd_sweep_kernel(const float* __restrict__ nw, const TR* __restrict__ rt,
                                 const TG* __restrict__ G, const T* __restrict__ sparse, int sparse_channels,
                                 float* __restrict__ Gnext, float* __restrict__ gn, float* __restrict__ gd0,
                                 int B, int C, int H, int W, int first, const __grid_constant__ TapTable tt)
{
    const size_t hw = (size_t)H * W;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)B * hw) return;
    const int b = (int)(idx / hw);
    const size_t p = idx - (size_t)b * hw;
    const int y = (int)(p / W), x = (int)(p - (size_t)y * W);
    const float* nb = nw + (size_t)b * tt.n * hw;
    float* gnb = gn + (size_t)b * tt.n * hw + p;
    for (int c = 0; c < C; ++c) {
        const size_t plane = ((size_t)b * C + c) * hw;
        const T* sp = sparse ? sparse + ((size_t)b * sparse_channels + (sparse_channels == 1 ? 0 : c)) * hw : nullptr;
        const float m = sp ? signf(to_f32(sp[p])) : 0.f;
        const float gp = to_f32(G[plane + p]);
        const float u = (1.f - m) * gp;
        if (sp) gd0[plane + p] = (first ? 0.f : gd0[plane + p]) + m * gp;
        else if (first) gd0[plane + p] = 0.f;
        float acc = 0.f;
        for (int k = 0; k < tt.n; ++k) {
            const int yy = y + tt.dy[k], xx = x + tt.dx[k];
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                const float add = u * to_f32(rt[plane + (size_t)yy * W + xx]);
                gnb[k * hw] = ((first && c == 0) ? 0.f : gnb[k * hw]) + add;
            } else if (first && c == 0) {
                gnb[k * hw] = 0.f;
            }
            const int ys = y - tt.dy[k], xs = x - tt.dx[k];
            if (ys >= 0 && ys < H && xs >= 0 && xs < W) {
                const size_t s = (size_t)ys * W + xs;
                const float ms = sp ? signf(to_f32(sp[s])) : 0.f;
                acc = fmaf((1.f - ms) * to_f32(G[plane + s]), nb[k * hw + s], acc);
            }
        }
        Gnext[plane + p] = acc;
    }
}

template <typename T>
__global__ void bwd_finish_depth_kernel(const float* __restrict__ gd0, const float* __restrict__ G0,
                                        T* __restrict__ grad_depth, size_t n)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) grad_depth[idx] = from_f32<T>(gd0[idx] + G0[idx]);
}

// Mode NEW: dL/dW_k(p) = (gn_k - sum_j n_j gn_j) / S(p); a_k(q) with q = p + o_k feeds W_k(p), so the
// value is written to grad_guidance[b, k, q] * sign(g) - every in-image (k, q) with an in-image p is
// written exactly once, the rest stays at the zero fill.  Mode OURS: softmax Jacobian in place.
template <typename T>
__global__ void bwd_finish_guidance_kernel(const T* __restrict__ g, int64_t gbs, const float* __restrict__ nw,
                                           const float* __restrict__ ssum, const float* __restrict__ gn,
                                           T* __restrict__ gg, int Cg, int B, int H, int W, int mode, const __grid_constant__ TapTable tt)
{
    const size_t hw = (size_t)H * W;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)B * hw) return;
    const int b = (int)(idx / hw);
    const size_t p = idx - (size_t)b * hw;
    const int y = (int)(p / W), x = (int)(p - (size_t)y * W);
    const float* nb = nw + (size_t)b * tt.n * hw + p;
    const float* gnb = gn + (size_t)b * tt.n * hw + p;
    T* ggb = gg + (size_t)b * Cg * hw;
    float dot = 0.f;
    for (int k = 0; k < tt.n; ++k) dot = fmaf(nb[k * hw], gnb[k * hw], dot);
    if (mode == CSPN_MODE_NEW) {
        const float s = ssum[idx];
        const T* gb = g + (size_t)b * gbs;
        for (int k = 0; k < tt.n; ++k) {
            const int yy = y + tt.dy[k], xx = x + tt.dx[k];
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                const size_t q = k * hw + (size_t)yy * W + xx;
                ggb[q] = from_f32<T>(signf(to_f32(gb[q])) * (gnb[k * hw] - dot) / s);
            }
        }
    } else {
        for (int k = 0; k < tt.n; ++k) ggb[k * hw + p] = from_f32<T>(nb[k * hw] * (gnb[k * hw] - dot));
    }
}

#define CSPN_LAUNCH_CHECK()                                  \
    do {                                                     \
        cudaError_t e_ = cudaGetLastError();                 \
        if (e_ != cudaSuccess) return (int)e_;               \
        ++call_stats().launches;                             \
    } while (0)

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

size_t generic_fwd_workspace(int B, int C, int H, int W, int taps)
{
    const size_t hw = (size_t)H * W;
    return align256((size_t)B * taps * hw * 4) + 2 * align256((size_t)B * C * hw * 4);
}

size_t generic_bwd_workspace(int B, int C, int H, int W, int iters, int taps)
{
    const size_t hw = (size_t)H * W;
    // nw, gn: B*taps planes; ssum: B planes; history r^1..r^{T-1}: (iters-1) * B*C; G ping-pong 2 * B*C; gd0 B*C
    const size_t hist = iters > 1 ? (size_t)(iters - 1) : 0;
    return 2 * align256((size_t)B * taps * hw * 4) + align256((size_t)B * hw * 4) +
           align256(hist * B * C * hw * 4) + 3 * align256((size_t)B * C * hw * 4);
}

template <typename T>
int generic_forward(const FwdArgs<T>& a, const TapTable& tt)
{
    const size_t hw = (size_t)a.H * a.W;
    const size_t npx = (size_t)a.B * a.C * hw;
    if (a.iters == 0) {
        copy_kernel<T><<<blocks_for(npx), kThreads, 0, a.stream>>>(a.depth, a.out, npx);
        CSPN_LAUNCH_CHECK();
        return 0;
    }
    char* ws = (char*)a.ws;
    float* nw = (float*)ws; ws += align256((size_t)a.B * tt.n * hw * 4);
    float* r0 = (float*)ws; ws += align256(npx * 4);
    float* r1 = (float*)ws;
    prep_weights_kernel<T><<<blocks_for((size_t)a.B * hw), kThreads, 0, a.stream>>>(a.guidance, a.gbs, nw, nullptr, a.B, a.H, a.W, a.mode, tt);
    CSPN_LAUNCH_CHECK();
    const unsigned grid = blocks_for(npx);
    if (a.iters == 1) {
        sweep_kernel<T, T, T><<<grid, kThreads, 0, a.stream>>>(nw, a.depth, a.depth, a.sparse, a.sparse_channels, a.out, a.B, a.C, a.H, a.W, tt);
        CSPN_LAUNCH_CHECK();
        return 0;
    }
    sweep_kernel<T, T, float><<<grid, kThreads, 0, a.stream>>>(nw, a.depth, a.depth, a.sparse, a.sparse_channels, r0, a.B, a.C, a.H, a.W, tt);
    CSPN_LAUNCH_CHECK();
    float* cur = r0; float* nxt = r1;
    for (int t = 1; t < a.iters - 1; ++t) {
        sweep_kernel<T, float, float><<<grid, kThreads, 0, a.stream>>>(nw, cur, a.depth, a.sparse, a.sparse_channels, nxt, a.B, a.C, a.H, a.W, tt);
        CSPN_LAUNCH_CHECK();
        float* tmp = cur; cur = nxt; nxt = tmp;
    }
    sweep_kernel<T, float, T><<<grid, kThreads, 0, a.stream>>>(nw, cur, a.depth, a.sparse, a.sparse_channels, a.out, a.B, a.C, a.H, a.W, tt);
    CSPN_LAUNCH_CHECK();
    return 0;
}

template <typename T>
int generic_backward(const BwdArgs<T>& a, const TapTable& tt)
{
    const size_t hw = (size_t)a.H * a.W;
    const size_t npx = (size_t)a.B * a.C * hw;
    const size_t nimg = (size_t)a.B * hw;
    cudaError_t e = cudaMemsetAsync(a.grad_guidance, 0, (size_t)a.B * a.Cg * hw * sizeof(T), a.stream);
    if (e != cudaSuccess) return (int)e;
    if (a.iters == 0) {
        copy_kernel<T><<<blocks_for(npx), kThreads, 0, a.stream>>>(a.grad_out, a.grad_depth, npx);
        CSPN_LAUNCH_CHECK();
        return 0;
    }
    char* ws = (char*)a.ws;
    float* nw = (float*)ws; ws += align256((size_t)a.B * tt.n * hw * 4);
    float* gn = (float*)ws; ws += align256((size_t)a.B * tt.n * hw * 4);
    float* ssum = (float*)ws; ws += align256(nimg * 4);
    float* hist = (float*)ws; ws += align256((a.iters > 1 ? (size_t)(a.iters - 1) : 0) * npx * 4);
    float* G0 = (float*)ws; ws += align256(npx * 4);
    float* G1 = (float*)ws; ws += align256(npx * 4);
    float* gd0 = (float*)ws;
    const unsigned grid = blocks_for(npx), gimg = blocks_for(nimg);

    prep_weights_kernel<T><<<gimg, kThreads, 0, a.stream>>>(a.guidance, a.gbs, nw, ssum, a.B, a.H, a.W, a.mode, tt);
    CSPN_LAUNCH_CHECK();
    // forward recompute keeping r^1 .. r^{T-1} (r^0 is the depth input itself)
    for (int t = 1; t < a.iters; ++t) {
        float* dst = hist + (size_t)(t - 1) * npx;
        if (t == 1) sweep_kernel<T, T, float><<<grid, kThreads, 0, a.stream>>>(nw, a.depth, a.depth, a.sparse, a.sparse_channels, dst, a.B, a.C, a.H, a.W, tt);
        else sweep_kernel<T, float, float><<<grid, kThreads, 0, a.stream>>>(nw, dst - npx, a.depth, a.sparse, a.sparse_channels, dst, a.B, a.C, a.H, a.W, tt);
        CSPN_LAUNCH_CHECK();
    }
    // reverse sweeps t = T-1 .. 0
    float* cur = G0; float* nxt = G1;
    for (int t = a.iters - 1; t >= 0; --t) {
        const int first = (t == a.iters - 1);
        if (first && t == 0)
            bwd_sweep_kernel<T, T, T><<<gimg, kThreads, 0, a.stream>>>(nw, a.depth, a.grad_out, a.sparse, a.sparse_channels, nxt, gn, gd0, a.B, a.C, a.H, a.W, first, tt);
        else if (first)
            bwd_sweep_kernel<T, float, T><<<gimg, kThreads, 0, a.stream>>>(nw, hist + (size_t)(t - 1) * npx, a.grad_out, a.sparse, a.sparse_channels, nxt, gn, gd0, a.B, a.C, a.H, a.W, first, tt);
        else if (t == 0)
            bwd_sweep_kernel<T, T, float><<<gimg, kThreads, 0, a.stream>>>(nw, a.depth, cur, a.sparse, a.sparse_channels, nxt, gn, gd0, a.B, a.C, a.H, a.W, first, tt);
        else
            bwd_sweep_kernel<T, float, float><<<gimg, kThreads, 0, a.stream>>>(nw, hist + (size_t)(t - 1) * npx, cur, a.sparse, a.sparse_channels, nxt, gn, gd0, a.B, a.C, a.H, a.W, first, tt);
        CSPN_LAUNCH_CHECK();
        float* tmp = cur; cur = nxt; nxt = tmp;
    }
    bwd_finish_depth_kernel<T><<<grid, kThreads, 0, a.stream>>>(gd0, cur, a.grad_depth, npx);
    CSPN_LAUNCH_CHECK();
    bwd_finish_guidance_kernel<T><<<gimg, kThreads, 0, a.stream>>>(a.guidance, a.gbs, nw, ssum, gn, a.grad_guidance, a.Cg, a.B, a.H, a.W, a.mode, tt);
    CSPN_LAUNCH_CHECK();
    return 0;
}

template int generic_forward<float>(const FwdArgs<float>&, const TapTable&);
template int generic_forward<__half>(const FwdArgs<__half>&, const TapTable&);
template int generic_backward<float>(const BwdArgs<float>&, const TapTable&);
template int generic_backward<__half>(const BwdArgs<__half>&, const TapTable&);

}  // namespace cspn
