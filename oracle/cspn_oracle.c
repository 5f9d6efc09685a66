/* CPU oracle for the CSPN affinity-propagation hot path - plain C restatement.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/cspn_oracle.py for the rules): the product
 * path never links or calls this file.  It exists so that parity tests can check
 * full-size images in seconds and so that bench.py has a multi-threaded CPU
 * baseline ("port") to time beside the B200 kernel.
 *
 * Pinned against the reference's own outputs through tests/golden/cspn_golden.npz
 * (tests/test_oracle_golden.py).
 *
 * Reference lines restated (paths relative to the reference checkout):
 *   mode 0  network/libs/post_process/CSPN_new.py:26-128   (abs, neighbour-indexed weights,
 *           border renormalisation through the gathered denominator, :121-128)
 *   mode 1  network/libs/post_process/CSPN_ours.py:24-54 + network/libs/base/pac.py:75-121
 *           (softmax over K*K-1 channels, centre tap zero, centre-indexed weights)
 *   both    re-injection r = (1-m)*r + m*d0 with m = sign(sparse)
 *           (CSPN_new.py:77-78,89-90; CSPN_ours.py:43-45,51-53)
 *
 * Layout: NCHW contiguous fp32; guidance may carry extra channels (Cg >= taps) and is
 * addressed through its own batch stride.  Depth may have C channels sharing one
 * affinity; sparse has 1 channel (broadcast) or C channels.
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC -o liboracle.so cspn_oracle.c -lm   (oracle/Makefile)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAX_TAPS 48

static float sgn(float v) { return (float)((v > 0.f) - (v < 0.f)); }

/* Fill the tap offset table.  Returns the tap count, or -1 for an unsupported config. */
static int tap_offsets(int mode, int ksize, int* dy, int* dx)
{
    if (mode == 0) {
        /* CSPN_new.py:43-67 pads + :87 crop, see oracle/cspn_oracle.py MODE_A_OFFSETS */
        static const int ady[8] = { +1, +1, +1, 0, 0, -1, -1, -1 };
        static const int adx[8] = { +1, 0, -1, +1, -1, +1, 0, -1 };
        if (ksize != 3) return -1;
        memcpy(dy, ady, sizeof ady); memcpy(dx, adx, sizeof adx);
        return 8;
    }
    if (ksize < 3 || (ksize & 1) == 0 || ksize * ksize - 1 > MAX_TAPS) return -1;
    int p = ksize / 2, n = 0;
    for (int iy = 0; iy < ksize; ++iy)
        for (int ix = 0; ix < ksize; ++ix) {
            if (iy == p && ix == p) continue;            /* centre tap stays zero, CSPN_ours.py:37-39 */
            dy[n] = iy - p; dx[n] = ix - p; ++n;
        }
    return n;
}

/* Normalised per-pixel tap weights n[k][y][x] for one image (weights located at the centre).
 * mode 0: n_k(p) = |g_k(p+o_k)| / sum_j |g_j(p+o_j)| over in-bounds neighbours (0/0 -> NaN kept).
 * mode 1: softmax over channels at p. */
static void tap_weights(const float* g, int mode, int taps, const int* dy, const int* dx,
                        int H, int W, float* n, float* ssum /* mode 0: S(p), may be NULL */)
{
    const size_t hw = (size_t)H * W;
    if (mode == 0) {
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                float wk[8], s = 0.f;
                for (int k = 0; k < 8; ++k) {
                    int yy = y + dy[k], xx = x + dx[k];
                    wk[k] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? fabsf(g[k * hw + (size_t)yy * W + xx]) : 0.f;
                    s += wk[k];                                        /* sequential, CSPN_new.py:124 */
                }
                for (int k = 0; k < 8; ++k) n[k * hw + (size_t)y * W + x] = wk[k] / s;
                if (ssum) ssum[(size_t)y * W + x] = s;
            }
    } else {
        for (size_t p = 0; p < hw; ++p) {
            float mx = -INFINITY, s = 0.f;
            for (int k = 0; k < taps; ++k) mx = fmaxf(mx, g[k * hw + p]);
            for (int k = 0; k < taps; ++k) { float e = expf(g[k * hw + p] - mx); n[k * hw + p] = e; s += e; }
            for (int k = 0; k < taps; ++k) n[k * hw + p] /= s;
        }
    }
}

/* One propagation sweep of one plane: out(p) = (1-m)*sum_k w_k(p)*r(p+o_k) + m*d0(p). */
static void sweep(const float* n, const float* wraw, const float* ssum, int mode, int taps,
                  const int* dy, const int* dx, int H, int W,
                  const float* r, const float* d0, const float* sparse, float* out)
{
    const size_t hw = (size_t)H * W;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const size_t p = (size_t)y * W + x;
            float acc = 0.f;
            for (int k = 0; k < taps; ++k) {
                int yy = y + dy[k], xx = x + dx[k];
                if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;  /* zero pad: 0*0 */
                acc += (mode == 0 ? wraw[k * hw + p] : n[k * hw + p]) * r[(size_t)yy * W + xx];
            }
            if (mode == 0) acc = acc / ssum[p];                        /* one divide, CSPN_new.py:127 */
            if (sparse) { float m = sgn(sparse[p]); acc = (1.f - m) * acc + m * d0[p]; }
            out[p] = acc;
        }
}

/* Forward. Returns 0 on success, nonzero for bad arguments. */
int cspn_oracle_forward(const float* guidance, int64_t g_batch_stride, const float* depth,
                        const float* sparse, int sparse_channels, float* out,
                        int B, int C, int H, int W, int iters, int ksize, int mode, int threads)
{
    int dy[MAX_TAPS], dx[MAX_TAPS];
    const int taps = tap_offsets(mode, ksize, dy, dx);
    if (taps < 0 || B < 0 || C < 1 || H < 1 || W < 1 || iters < 0) return 1;
    const size_t hw = (size_t)H * W;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int b = 0; b < B; ++b) {
        float* n = (float*)malloc(sizeof(float) * hw * taps);
        float* wraw = NULL; float* ssum = NULL;
        float* r0 = (float*)malloc(sizeof(float) * hw);
        float* r1 = (float*)malloc(sizeof(float) * hw);
        const float* g = guidance + (size_t)b * g_batch_stride;
        if (mode == 0) {
            /* keep raw gathered weights and S: the reference divides the finished sum (CSPN_new.py:125-127) */
            wraw = (float*)malloc(sizeof(float) * hw * 8); ssum = (float*)malloc(sizeof(float) * hw);
            tap_weights(g, 0, 8, dy, dx, H, W, n, ssum);
            for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) for (int k = 0; k < 8; ++k) {
                int yy = y + dy[k], xx = x + dx[k];
                wraw[k * hw + (size_t)y * W + x] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? fabsf(g[k * hw + (size_t)yy * W + xx]) : 0.f;
            }
        } else {
            tap_weights(g, 1, taps, dy, dx, H, W, n, NULL);
        }
        for (int c = 0; c < C; ++c) {
            const float* d0 = depth + ((size_t)b * C + c) * hw;
            const float* sp = sparse ? sparse + ((size_t)b * sparse_channels + (sparse_channels == 1 ? 0 : c)) * hw : NULL;
            float* o = out + ((size_t)b * C + c) * hw;
            const float* cur = d0; float* nxt = r0;
            for (int t = 0; t < iters; ++t) {
                sweep(n, wraw, ssum, mode, taps, dy, dx, H, W, cur, d0, sp, nxt);
                cur = nxt; nxt = (nxt == r0) ? r1 : r0;
            }
            memcpy(o, cur, sizeof(float) * hw);
        }
        free(n); free(wraw); free(ssum); free(r0); free(r1);
    }
    return 0;
}

/* Backward. grad_guidance [B,Cg,H,W] (fully written, channels >= taps
 * zeroed), grad_depth [B,C,H,W].  Accumulates in double for a tight reference. */
int cspn_oracle_backward(const float* guidance, int64_t g_batch_stride, int Cg, const float* depth,
                         const float* sparse, int sparse_channels, const float* grad_out,
                         float* grad_guidance, float* grad_depth,
                         int B, int C, int H, int W, int iters, int ksize, int mode, int threads)
{
    int dy[MAX_TAPS], dx[MAX_TAPS];
    const int taps = tap_offsets(mode, ksize, dy, dx);
    if (taps < 0 || Cg < taps || B < 0 || C < 1 || H < 1 || W < 1 || iters < 0) return 1;
    const size_t hw = (size_t)H * W;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int b = 0; b < B; ++b) {
        const float* g = guidance + (size_t)b * g_batch_stride;
        float* n = (float*)malloc(sizeof(float) * hw * taps);
        float* ssum = (float*)malloc(sizeof(float) * hw);
        double* gn = (double*)calloc(hw * taps, sizeof(double));
        float* hist = (float*)malloc(sizeof(float) * hw * (size_t)(iters + 1));
        double* gt = (double*)malloc(sizeof(double) * hw);
        double* gnext = (double*)malloc(sizeof(double) * hw);
        float* wraw = NULL;
        tap_weights(g, mode, taps, dy, dx, H, W, n, ssum);
        if (mode == 0) {
            wraw = (float*)malloc(sizeof(float) * hw * 8);
            for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) for (int k = 0; k < 8; ++k) {
                int yy = y + dy[k], xx = x + dx[k];
                wraw[k * hw + (size_t)y * W + x] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? fabsf(g[k * hw + (size_t)yy * W + xx]) : 0.f;
            }
        }
        for (int c = 0; c < C; ++c) {
            const float* d0 = depth + ((size_t)b * C + c) * hw;
            const float* sp = sparse ? sparse + ((size_t)b * sparse_channels + (sparse_channels == 1 ? 0 : c)) * hw : NULL;
            const float* go = grad_out + ((size_t)b * C + c) * hw;
            float* gd = grad_depth + ((size_t)b * C + c) * hw;
            memcpy(hist, d0, sizeof(float) * hw);
            for (int t = 0; t < iters; ++t)
                sweep(n, wraw, ssum, mode, taps, dy, dx, H, W, hist + (size_t)t * hw, d0, sp, hist + (size_t)(t + 1) * hw);
            for (size_t p = 0; p < hw; ++p) { gt[p] = go[p]; gd[p] = 0.f; }
            double* gdacc = (double*)calloc(hw, sizeof(double));
            for (int t = iters - 1; t >= 0; --t) {
                const float* r = hist + (size_t)t * hw;
                memset(gnext, 0, sizeof(double) * hw);
                for (int y = 0; y < H; ++y)
                    for (int x = 0; x < W; ++x) {
                        const size_t p = (size_t)y * W + x;
                        double m = sp ? (double)sgn(sp[p]) : 0.0;
                        double u = (1.0 - m) * gt[p];
                        gdacc[p] += m * gt[p];
                        for (int k = 0; k < taps; ++k) {
                            int yy = y + dy[k], xx = x + dx[k];
                            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                            const size_t q = (size_t)yy * W + xx;
                            gn[k * hw + p] += u * r[q];
                            gnext[q] += u * (double)n[k * hw + p];
                        }
                    }
                double* tmp = gt; gt = gnext; gnext = tmp;
            }
            for (size_t p = 0; p < hw; ++p) gd[p] = (float)(gdacc[p] + gt[p]);
            free(gdacc);
        }
        float* gg = grad_guidance + (size_t)b * Cg * hw;
        memset(gg, 0, sizeof(float) * hw * Cg);
        if (mode == 0) {
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    const size_t p = (size_t)y * W + x;
                    double dot = 0.0;
                    for (int k = 0; k < 8; ++k) dot += (double)n[k * hw + p] * gn[k * hw + p];
                    for (int k = 0; k < 8; ++k) {
                        int yy = y + dy[k], xx = x + dx[k];
                        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                        const size_t q = (size_t)yy * W + xx;        /* a_k(q) feeds W_k(p), q = p + o_k */
                        gg[k * hw + q] = (float)(sgn(g[k * hw + q]) * (gn[k * hw + p] - dot) / (double)ssum[p]);
                    }
                }
        } else {
            for (size_t p = 0; p < hw; ++p) {
                double dot = 0.0;
                for (int k = 0; k < taps; ++k) dot += (double)n[k * hw + p] * gn[k * hw + p];
                for (int k = 0; k < taps; ++k) gg[k * hw + p] = (float)((double)n[k * hw + p] * (gn[k * hw + p] - dot));
            }
        }
        free(n); free(ssum); free(gn); free(hist); free(gt); free(gnext); free(wraw);
    }
    return 0;
}
