// The two output heads that feed the CSPN module: Simple_Gudi_UpConv_Block_Last_Layer x 2,
//   network/unet_cspn_nyu.py:195-218 (class), :331-332 (64 -> 1 blur depth, 64 -> 12 guidance), :383-384 (both applied to the same x)
//   network/unet_ours.py:194-202, :278-279 (64 -> 1, 64 -> 8), :331-332
// Reference per head: nearest 2x upsample, crop to (oheight, owidth), multiply by a mask that keeps the even (row, column)
// positions - built by an O(H W) Python loop of one-element assignments, unet_cspn_nyu.py:208-212 - and a dense 3x3 convolution
// over the result, 3/4 of whose inputs are zeros.  Here:
//   out[o, 2i+py, 2j+px] = sum_c sum_{ni <= py, nj <= px} W[o, c, 2ni-py+1, 2nj-px+1] * x[c, i+ni, j+nj]
// i.e. every half-resolution cell (i, j) produces its 2x2 output pixels from the 2x2 cells x[i..i+1, j..j+1]; each of the 9
// taps is used exactly once per cell (9/4 taps per output pixel instead of 9), no upsampled tensor, no mask, and BOTH heads
// come out of one pass over x (their weights are stacked).  fp32 accumulation in CUDA cores: K = 64..256 with N = 9 / 13 is too
// thin for a tensor-core tile to pay without giving up fp32 parity with the reference's fp32 convolution.
//   forward       thread = 2 adjacent cells x all NO outputs (8 x NO accumulators), x chunks and the weights in shared memory
//   backward x    thread = 2 adjacent cells x 32 input channels, grad_out window and weights in shared memory
//   backward W    split-K GEMM  gW[NO*9, Cin] = GO9[NO*9, cells] X[cells, Cin]: per-CTA partial sums in registers, a second
//                 kernel adds the partials in a fixed order (deterministic, no floating-point atomics)
#include "cspn_common.cuh"

namespace cspn {
namespace {

constexpr int kHT = 128;                  // threads per CTA (forward, backward x)
constexpr int kTR = 8, kTC = 32;          // cell tile of the forward / backward-x kernels
constexpr int kCK = 8;                    // input channels per shared-memory chunk (forward), double-buffered
constexpr int kXP = 36;                   // pitch of a staged x row (33 columns)
constexpr int kNOP = 16;                  // outputs padded to 16 in shared memory
constexpr int kMaxCin = 256;

template <typename T>
__device__ __forceinline__ float w_at(const T* w1, const T* w2, int n1, int n2, int Cin, int o, int c, int tap)
{
    if (o < n1) return to_f32(w1[((size_t)o * Cin + c) * 9 + tap]);
    return o < n1 + n2 ? to_f32(w2[((size_t)(o - n1) * Cin + c) * 9 + tap]) : 0.f;        // NO may be padded above n1 + n2
}
template <typename T>
__device__ __forceinline__ float go_at(const T* go1, const T* go2, int n1, int n2, int b, int o, size_t HW, size_t off)
{
    if (o < n1) return to_f32(go1[((size_t)b * n1 + o) * HW + off]);
    return o < n1 + n2 ? to_f32(go2[((size_t)b * n2 + (o - n1)) * HW + off]) : 0.f;
}

// Gather `count` elements into shared memory with U loads per thread in flight: addr(idx) returns the global address of element idx
// or nullptr (element is zero), put(idx, v) stores it.  The loads of a batch are predicated, not branched around, so that they
// issue back to back - a gather loop with a branch per element serialises on the full memory latency of every element.
template <int U, int NT, typename T, typename AddrFn, typename PutFn>
__device__ __forceinline__ void gather_batched(int count, AddrFn addr, PutFn put)
{
    for (int base = threadIdx.x; base < count; base += U * NT) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = base + u * NT;
            const T* q = idx < count ? addr(idx) : nullptr;
            v[u] = q ? to_f32(__ldg(q)) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = base + u * NT;
            if (idx < count) put(idx, v[u]);
        }
    }
}
template <typename T>
__device__ __forceinline__ const T* w_ptr(const T* w1, const T* w2, int n1, int n2, int Cin, int o, int c, int tap)
{
    if (o < n1) return w1 + ((size_t)o * Cin + c) * 9 + tap;
    return o < n1 + n2 ? w2 + ((size_t)(o - n1) * Cin + c) * 9 + tap : nullptr;
}
template <typename T>
__device__ __forceinline__ const T* go_ptr(const T* go1, const T* go2, int n1, int n2, int b, int o, size_t HW, size_t off)
{
    if (o < n1) return go1 + ((size_t)b * n1 + o) * HW + off;
    return o < n1 + n2 ? go2 + ((size_t)b * n2 + (o - n1)) * HW + off : nullptr;
}

// ---------------------------------------------------------------------------------------------------------------- forward
template <typename T, int NO>
__global__ void __launch_bounds__(kHT)
heads_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w1, const T* __restrict__ w2, T* __restrict__ out1, T* __restrict__ out2,
                 int n1, int n2, int Cin, int h, int w, int H, int W)
{
    extern __shared__ __align__(16) float smem[];
    float* wsm = smem;                                    // [Cin][9][kNOP]
    float* xs = smem + (size_t)Cin * 9 * kNOP;            // [2][kCK][kTR + 1][kXP]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int b = blockIdx.z, i0 = blockIdx.y * kTR, j0 = blockIdx.x * kTC;
    const int hs = (H + 1) >> 1, ws = (W + 1) >> 1;      // cells that survive the crop + mask
    for (int idx = tid; idx < kNOP * Cin * 9; idx += kHT) {                // global order (o, c, tap): coalesced reads
        const int o = idx / (Cin * 9), ct = idx - o * (Cin * 9);
        wsm[ct * kNOP + o] = o < NO ? w_at(w1, w2, n1, n2, Cin, o, ct / 9, ct % 9) : 0.f;
    }
    float acc[2][4][NO];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int o = 0; o < NO; ++o) acc[a][p][o] = 0.f;

    const T* xb = x + (size_t)b * Cin * h * w;
    // stage one chunk of x: (kCK channels) x (kTR + 1 rows) x 33 columns; a warp takes whole rows (one coalesced 128-byte
    // request + the 33rd column), fp32 goes straight to shared memory with cp.async (zero-filled outside the kept cells)
    auto stage = [&](int c0, float* dst) {
        const int lane = tid & 31, warp = tid >> 5;
        for (int row = warp; row < kCK * (kTR + 1); row += kHT / 32) {
            const int c = row / (kTR + 1), r = row - c * (kTR + 1);
            const int i = i0 + r;
            const bool row_ok = c0 + c < Cin && i < hs;
            const T* src = xb + ((size_t)(row_ok ? c0 + c : 0) * h + (row_ok ? i : 0)) * w;
            float* d = dst + row * kXP;
#pragma unroll
            for (int rep = 0; rep < 2; ++rep) {
                const int col = rep == 0 ? lane : 32;
                if (rep == 1 && lane != 0) break;
                const bool ok = row_ok && j0 + col < ws;
                if (sizeof(T) == 4) {
                    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(d + col);
                    const void* ga = ok ? (const void*)(src + j0 + col) : (const void*)xb;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sa), "l"(ga), "r"(ok ? 4 : 0) : "memory");
                } else {
                    d[col] = ok ? to_f32(src[j0 + col]) : 0.f;
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int nchunks = (Cin + kCK - 1) / kCK;
    stage(0, xs);
    for (int ck = 0; ck < nchunks; ++ck) {
        const int c0 = ck * kCK;
        float* cur = xs + (ck & 1) * (kCK * (kTR + 1) * kXP);
        if (ck + 1 < nchunks) {
            stage(c0 + kCK, xs + ((ck + 1) & 1) * (kCK * (kTR + 1) * kXP));
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const int nc = Cin - c0 < kCK ? Cin - c0 : kCK;
        for (int c = 0; c < nc; ++c) {
            const float* r0 = cur + (c * (kTR + 1) + ty) * kXP + 2 * tx;
            const float* r1 = r0 + kXP;
            const float xa[2][4] = {{r0[0], r0[1], r1[0], r1[1]}, {r0[1], r0[2], r1[1], r1[2]}};     // {self, right, down, diagonal} of cells A, B
            const float* wc = wsm + (size_t)(c0 + c) * 9 * kNOP;
            float wt[9][NO];
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                for (int o4 = 0; o4 < (NO + 3) / 4; ++o4) {
                    const float4 v = *reinterpret_cast<const float4*>(wc + tap * kNOP + 4 * o4);
                    if (4 * o4 + 0 < NO) wt[tap][4 * o4 + 0] = v.x;
                    if (4 * o4 + 1 < NO) wt[tap][4 * o4 + 1] = v.y;
                    if (4 * o4 + 2 < NO) wt[tap][4 * o4 + 2] = v.z;
                    if (4 * o4 + 3 < NO) wt[tap][4 * o4 + 3] = v.w;
                }
            }
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int o = 0; o < NO; ++o) {
                    // tap = ky * 3 + kx with ky = 2 ni - py + 1, kx = 2 nj - px + 1
                    acc[a][0][o] = fmaf(wt[4][o], xa[a][0], acc[a][0][o]);
                    acc[a][1][o] = fmaf(wt[5][o], xa[a][1], fmaf(wt[3][o], xa[a][0], acc[a][1][o]));
                    acc[a][2][o] = fmaf(wt[7][o], xa[a][2], fmaf(wt[1][o], xa[a][0], acc[a][2][o]));
                    acc[a][3][o] = fmaf(wt[8][o], xa[a][3], fmaf(wt[6][o], xa[a][2], fmaf(wt[2][o], xa[a][1], fmaf(wt[0][o], xa[a][0], acc[a][3][o]))));
                }
        }
        __syncthreads();                                  // this buffer is staged again two chunks later
    }
    // 4 adjacent output pixels per (output channel, parity row): columns 2 (j0 + 2 tx) .. + 3
    const int i = i0 + ty, X0 = 2 * (j0 + 2 * tx);
    const size_t HW = (size_t)H * W;
    const bool vec = (W % 4 == 0) && ((uintptr_t)out1 % (4 * sizeof(T)) == 0) && (n2 == 0 || (uintptr_t)out2 % (4 * sizeof(T)) == 0);
#pragma unroll
    for (int o = 0; o < NO; ++o) {
        if (o >= n1 + n2) break;
        T* plane = o < n1 ? out1 + ((size_t)b * n1 + o) * HW : out2 + ((size_t)b * n2 + (o - n1)) * HW;
#pragma unroll
        for (int py = 0; py < 2; ++py) {
            const int Y = 2 * i + py;
            if (Y >= H || X0 >= W) continue;
            const float v[4] = {acc[0][2 * py][o], acc[0][2 * py + 1][o], acc[1][2 * py][o], acc[1][2 * py + 1][o]};
            T* row = plane + (size_t)Y * W;
            if (vec) {
                if (sizeof(T) == 4) *reinterpret_cast<float4*>(row + X0) = make_float4(v[0], v[1], v[2], v[3]);
                else {
                    uint2 u;
                    *reinterpret_cast<__half2*>(&u.x) = __floats2half2_rn(v[0], v[1]);
                    *reinterpret_cast<__half2*>(&u.y) = __floats2half2_rn(v[2], v[3]);
                    *reinterpret_cast<uint2*>(row + X0) = u;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (X0 + q < W) row[X0 + q] = from_f32<T>(v[q]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------- backward x
// gx[c, i, j] = sum_o sum_{dy,dx in -1..1} W[o, c, 1-dy, 1-dx] * go[o, 2i+dy, 2j+dx]   (cells with 2i >= H or 2j >= W: 0)
constexpr int kBC = 32;                   // input channels per CTA
constexpr int kGP = 72;                   // pitch of a staged grad_out row: index 4 + (X - 2 j0), X = 2 j0 - 1 .. 2 j0 + 64
constexpr int kGR = 2 * kTR + 1;          // rows 2 i0 - 1 .. 2 i0 + 15

template <typename T, int NO>
__global__ void __launch_bounds__(kHT)
heads_bwd_x_kernel(const T* __restrict__ go1, const T* __restrict__ go2, const T* __restrict__ w1, const T* __restrict__ w2, T* __restrict__ gx,
                   int n1, int n2, int Cin, int h, int w, int H, int W)
{
    extern __shared__ __align__(16) float smem[];
    float* wsm = smem;                                    // [NO][9][kBC]
    float* gs = smem + NO * 9 * kBC;                      // [NO][kGR][kGP]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int nchunk = (Cin + kBC - 1) / kBC;
    const int b = blockIdx.z / nchunk, c0 = (blockIdx.z % nchunk) * kBC;
    const int i0 = blockIdx.y * kTR, j0 = blockIdx.x * kTC;
    const size_t HW = (size_t)H * W;
    gather_batched<8, kHT, T>(NO * kBC * 9,                                                   // global order (o, c, tap): coalesced reads
        [&](int idx) -> const T* {
            const int o = idx / (kBC * 9), ct = idx - o * (kBC * 9), c = ct / 9, tap = ct - 9 * c;
            return c0 + c < Cin ? w_ptr(w1, w2, n1, n2, Cin, o, c0 + c, tap) : nullptr;
        },
        [&](int idx, float v) {
            const int o = idx / (kBC * 9), ct = idx - o * (kBC * 9), c = ct / 9, tap = ct - 9 * c;
            wsm[(o * 9 + tap) * kBC + c] = v;
        });
    {
        // grad_out window: the (row, column) pattern of a thread's elements is the same for every output channel, so its image
        // offsets and shared-memory slots are worked out once (9 elements of the 17 x 66 window per thread) and the loop over the
        // outputs only adds plane bases: 9 independent loads per thread and output in flight, ~3 integer instructions per element.
        constexpr int kWin = kGR * 66, kPerT = (kWin + kHT - 1) / kHT;
        int goff[kPerT], soff[kPerT];
#pragma unroll
        for (int u = 0; u < kPerT; ++u) {
            const int e = tid + u * kHT, r = e / 66, col = e - 66 * r;
            const int Y = 2 * i0 - 1 + r, X = 2 * j0 - 1 + col;
            soff[u] = e < kWin ? r * kGP + 3 + col : -1;
            goff[u] = (e < kWin && Y >= 0 && Y < H && X >= 0 && X < W) ? Y * W + X : -1;
        }
#pragma unroll 1
        for (int o = 0; o < NO; ++o) {
            const T* plane = go_ptr(go1, go2, n1, n2, b, o, HW, 0);          // nullptr: padded output row
            float v[kPerT];
#pragma unroll
            for (int u = 0; u < kPerT; ++u) v[u] = (plane && goff[u] >= 0) ? to_f32(__ldg(plane + goff[u])) : 0.f;
#pragma unroll
            for (int u = 0; u < kPerT; ++u)
                if (soff[u] >= 0) gs[o * kGR * kGP + soff[u]] = v[u];
        }
    }
    __syncthreads();
    float acc[2][kBC];
#pragma unroll
    for (int c = 0; c < kBC; ++c) acc[0][c] = acc[1][c] = 0.f;
#pragma unroll 1
    for (int o = 0; o < NO; ++o) {
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const float* row = gs + (o * kGR + 2 * ty + 1 + dy) * kGP + 4 + 4 * tx;      // output column 2 (j0 + 2 tx)
            const float gm = row[-1];
            const float4 g4 = *reinterpret_cast<const float4*>(row);
            const float gv[5] = {gm, g4.x, g4.y, g4.z, g4.w};
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const float ga = gv[1 + dx], gb = gv[3 + dx];
                const float* wv = wsm + (o * 9 + (1 - dy) * 3 + (1 - dx)) * kBC;
#pragma unroll
                for (int c4 = 0; c4 < kBC / 4; ++c4) {
                    const float4 wq = *reinterpret_cast<const float4*>(wv + 4 * c4);
                    acc[0][4 * c4 + 0] = fmaf(wq.x, ga, acc[0][4 * c4 + 0]); acc[1][4 * c4 + 0] = fmaf(wq.x, gb, acc[1][4 * c4 + 0]);
                    acc[0][4 * c4 + 1] = fmaf(wq.y, ga, acc[0][4 * c4 + 1]); acc[1][4 * c4 + 1] = fmaf(wq.y, gb, acc[1][4 * c4 + 1]);
                    acc[0][4 * c4 + 2] = fmaf(wq.z, ga, acc[0][4 * c4 + 2]); acc[1][4 * c4 + 2] = fmaf(wq.z, gb, acc[1][4 * c4 + 2]);
                    acc[0][4 * c4 + 3] = fmaf(wq.w, ga, acc[0][4 * c4 + 3]); acc[1][4 * c4 + 3] = fmaf(wq.w, gb, acc[1][4 * c4 + 3]);
                }
            }
        }
    }
    const int hs = (H + 1) >> 1, ws = (W + 1) >> 1;
    const int i = i0 + ty, j = j0 + 2 * tx;
    if (i >= h) return;
    T* gxb = gx + (size_t)b * Cin * h * w + (size_t)i * w;
#pragma unroll
    for (int c = 0; c < kBC; ++c) {
        if (c0 + c >= Cin) break;
#pragma unroll
        for (int a = 0; a < 2; ++a)
            if (j + a < w) gxb[(size_t)(c0 + c) * h * w + j + a] = from_f32<T>((i < hs && j + a < ws) ? acc[a][c] : 0.f);   // cropped cells: zero gradient
    }
}

// ------------------------------------------------------------------------------------------------------------- backward W
// gW[o, c, 1-dy, 1-dx] = sum_{b,i,j} go[b, o, 2i+dy, 2j+dx] * x[b, c, i, j]: GEMM over the cells, split across CTAs.
constexpr int kKC = 32;                   // cells per shared-memory chunk
constexpr int kOT = 144;                  // (output, tap) rows: ot = o * 9 + tap, 16 outputs x 9 taps
constexpr int kWT = kOT / 8 * 8;          // threads of the backward-W kernel: 18 row groups x 8 channel groups
constexpr int kOTP = kOT + 4;

template <typename T, int NO>
__global__ void __launch_bounds__(kWT, 3)
heads_bwd_w_kernel(const T* __restrict__ x, const T* __restrict__ go1, const T* __restrict__ go2, float* __restrict__ partial,
                   int n1, int n2, int Cin, int h, int w, int H, int W, int B, int c_base)
{
    // one CTA: 64 input channels starting at c_base, all NO * 9 (o, tap) rows; thread tile 8 (o, tap) x 8 channels
    __shared__ __align__(16) float g9[kKC][kOTP];
    __shared__ __align__(16) float xs[kKC][64 + 4];
    __shared__ int cell_b[kKC], cell_i[kKC], cell_j[kKC];                 // (image, row, column) of the chunk's cells; image -1 = past the end
    const int tid = threadIdx.x, cg = tid & 7, og = tid >> 3;
    const int hs = (H + 1) >> 1, ws = (W + 1) >> 1;
    const long cells = (long)B * hs * ws;
    const size_t HW = (size_t)H * W;
    float acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[a][c] = 0.f;
    for (int idx = tid; idx < kKC * kOTP; idx += kWT) (&g9[0][0])[idx] = 0.f;              // rows >= NO * 9 stay zero
    for (long k0 = (long)blockIdx.x * kKC; k0 < cells; k0 += (long)gridDim.x * kKC) {
        __syncthreads();
        if (tid < kKC) {
            const long cell = k0 + tid;
            const int bb = cell < cells ? (int)(cell / ((long)hs * ws)) : -1;
            const int rem = cell < cells ? (int)(cell - (long)bb * hs * ws) : 0;
            cell_b[tid] = bb; cell_i[tid] = rem / ws; cell_j[tid] = rem % ws;
        }
        __syncthreads();
        if (tid < 4 * kKC) {
            // staging role of a thread: cell sk of the chunk, taps sq, sq + 4 (, sq + 8) of every output and channels sq + 4 u.  The
            // cell's image position is worked out once; all 3 NO + 16 loads of the thread are independent and issued together.
            const int sk = tid & (kKC - 1), sq = tid / kKC;
            const int bb = cell_b[sk], ci = cell_i[sk], cj = cell_j[sk];
            int toff[3];                                                                  // in-plane offset of the thread's taps, -1: not loaded
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                const int tap = sq + 4 * m, ky = tap / 3, kx = tap - 3 * ky;
                const int Y = 2 * ci + 1 - ky, X = 2 * cj + 1 - kx;                       // dy = 1 - ky, dx = 1 - kx
                toff[m] = (bb >= 0 && tap < 9 && Y >= 0 && Y < H && X >= 0 && X < W) ? Y * W + X : -1;
            }
            constexpr int kOG = 4;                                                        // outputs per batch: 12 loads in flight
#pragma unroll 1
            for (int o0 = 0; o0 < NO; o0 += kOG) {
                float gv[kOG][3];
#pragma unroll
                for (int oo = 0; oo < kOG; ++oo) {
                    const T* plane = (o0 + oo < NO && bb >= 0) ? go_ptr(go1, go2, n1, n2, bb, o0 + oo, HW, 0) : nullptr;
#pragma unroll
                    for (int m = 0; m < 3; ++m) gv[oo][m] = (plane && toff[m] >= 0) ? to_f32(__ldg(plane + toff[m])) : 0.f;
                }
#pragma unroll
                for (int oo = 0; oo < kOG; ++oo)
#pragma unroll
                    for (int m = 0; m < 3; ++m)
                        if (o0 + oo < NO && sq + 4 * m < 9) g9[sk][(o0 + oo) * 9 + sq + 4 * m] = gv[oo][m];
            }
            const T* xc = bb >= 0 ? x + (((size_t)bb * Cin + c_base) * h + ci) * w + cj : nullptr;
            const size_t cstride = (size_t)h * w;
#pragma unroll 1
            for (int u0 = 0; u0 < 16; u0 += 8) {
                float xv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int c = sq + 4 * (u0 + u);
                    xv[u] = (xc && c_base + c < Cin) ? to_f32(__ldg(xc + c * cstride)) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) xs[sk][sq + 4 * (u0 + u)] = xv[u];
            }
        }
        __syncthreads();
        if (8 * og >= NO * 9) continue;                                    // row groups past the last (o, tap) only help staging
#pragma unroll 4
        for (int k = 0; k < kKC; ++k) {
            const float4 ga = *reinterpret_cast<const float4*>(&g9[k][8 * og]), gb = *reinterpret_cast<const float4*>(&g9[k][8 * og + 4]);
            const float4 xa = *reinterpret_cast<const float4*>(&xs[k][8 * cg]), xb = *reinterpret_cast<const float4*>(&xs[k][8 * cg + 4]);
            const float gv[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
            const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[a][c] = fmaf(gv[a], xv[c], acc[a][c]);
        }
    }
    float* part = partial + (size_t)blockIdx.x * kOT * 64;
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int c = 0; c < 8; ++c) part[(8 * og + a) * 64 + 8 * cg + c] = acc[a][c];
}

// One CTA per (o, tap) row: 64 channels x 4 segments of the partials; every thread adds its segment in launch order, the four
// segment sums are added in fixed order - deterministic, and the loads of a thread are independent of each other.
constexpr int kRedSeg = 4;
template <typename T>
__global__ void __launch_bounds__(64 * kRedSeg)
heads_bwd_w_reduce_kernel(const float* __restrict__ partial, int nparts, T* __restrict__ gw1, T* __restrict__ gw2, int n1, int n2, int Cin, int c_base)
{
    __shared__ float seg_sum[kRedSeg][64];
    const int ot = blockIdx.x, c = threadIdx.x & 63, seg = threadIdx.x >> 6;
    const int per = (nparts + kRedSeg - 1) / kRedSeg, p0 = seg * per, p1 = p0 + per < nparts ? p0 + per : nparts;
    const float* src = partial + (size_t)ot * 64 + c;
    float s = 0.f;
    int p = p0;
    for (; p + 8 <= p1; p += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = src[(size_t)(p + u) * kOT * 64];
#pragma unroll
        for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; p < p1; ++p) s += src[(size_t)p * kOT * 64];
    seg_sum[seg][c] = s;
    __syncthreads();
    if (seg != 0 || c_base + c >= Cin) return;
    float tot = seg_sum[0][c];
#pragma unroll
    for (int g = 1; g < kRedSeg; ++g) tot += seg_sum[g][c];
    const int o = ot / 9, tap = ot - 9 * o;
    if (o < n1) gw1[((size_t)o * Cin + c_base + c) * 9 + tap] = from_f32<T>(tot);
    else gw2[((size_t)(o - n1) * Cin + c_base + c) * 9 + tap] = from_f32<T>(tot);
}

constexpr int kWParts = 444;              // CTAs of the backward-W kernel (3 per SM on a B200: 121 registers x 144 threads)

template <typename T, int NO>
int heads_forward_no(const T* x, const T* w1, const T* w2, T* out1, T* out2, int n1, int n2, int B, int Cin, int h, int w, int H, int W, cudaStream_t stream)
{
    auto kern = heads_fwd_kernel<T, NO>;
    const size_t smem = ((size_t)Cin * 9 * kNOP + (size_t)2 * kCK * (kTR + 1) * kXP) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int hs = (H + 1) / 2, ws = (W + 1) / 2;
    dim3 grid((unsigned)((ws + kTC - 1) / kTC), (unsigned)((hs + kTR - 1) / kTR), (unsigned)B);
    kern<<<grid, kHT, smem, stream>>>(x, w1, w2, out1, out2, n1, n2, Cin, h, w, H, W);
    e = cudaGetLastError();
    if (e == cudaSuccess) ++call_stats().launches;
    return (int)e;
}

template <typename T, int NO>
int heads_backward_no(const T* x, const T* w1, const T* w2, const T* go1, const T* go2, T* gx, T* gw1, T* gw2, int n1, int n2, int B, int Cin, int h, int w,
                      int H, int W, float* ws, cudaStream_t stream)
{
    cudaError_t e;
    if (gx) {
        auto kern = heads_bwd_x_kernel<T, NO>;
        const size_t smem = ((size_t)NO * 9 * kBC + (size_t)NO * kGR * kGP) * sizeof(float);
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        dim3 grid((unsigned)((w + kTC - 1) / kTC), (unsigned)((h + kTR - 1) / kTR), (unsigned)(B * ((Cin + kBC - 1) / kBC)));
        kern<<<grid, kHT, smem, stream>>>(go1, go2, w1, w2, gx, n1, n2, Cin, h, w, H, W);
        e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
        ++call_stats().launches;
    }
    if (gw1) {
        const long cells = (long)B * ((H + 1) / 2) * ((W + 1) / 2);
        const int parts = (int)((cells + kKC - 1) / kKC < kWParts ? (cells + kKC - 1) / kKC : kWParts);
        for (int c_base = 0; c_base < Cin; c_base += 64) {
            heads_bwd_w_kernel<T, NO><<<parts, kWT, 0, stream>>>(x, go1, go2, ws, n1, n2, Cin, h, w, H, W, B, c_base);
            e = cudaGetLastError();
            if (e != cudaSuccess) return (int)e;
            ++call_stats().launches;
            heads_bwd_w_reduce_kernel<T><<<(n1 + n2) * 9, 64 * kRedSeg, 0, stream>>>(ws, parts, gw1, gw2, n1, n2, Cin, c_base);
            e = cudaGetLastError();
            if (e != cudaSuccess) return (int)e;
            ++call_stats().launches;
        }
    }
    return 0;
}

}  // namespace

size_t heads_workspace_bytes() { return (size_t)kWParts * kOT * 64 * sizeof(float); }

bool heads_supported(int n1, int n2, int Cin, int h, int w, int H, int W)
{
    const int no = n1 + n2;
    return n1 >= 1 && n2 >= 0 && no <= 16 && Cin >= 1 && Cin <= kMaxCin && h >= 1 && w >= 1 && H >= 1 && W >= 1 && H <= 2 * h && W <= 2 * w &&
           (long)H * W < (1l << 31);                  // in-plane offsets are ints
}

// kernels are instantiated for the output counts of the reference's models (1 + 8 = 9, unet_ours.py:278-279; 1 + 12 = 13,
// unet_cspn_nyu.py:331-332), a single depth head (1) and the padded general case (16)
#define CSPN_HEADS_DISPATCH(CALL)                                   \
    {                                                               \
        const int no_ = n1 + n2;                                    \
        if (no_ == 1) return CALL(1);                               \
        if (no_ <= 9) return CALL(9);                               \
        if (no_ <= 13) return CALL(13);                             \
        if (no_ <= 16) return CALL(16);                             \
        return CSPN_ERR_BAD_SHAPE;                                  \
    }

template <typename T>
int heads_forward(const T* x, const T* w1, const T* w2, T* out1, T* out2, int n1, int n2, int B, int Cin, int h, int w, int H, int W, cudaStream_t stream)
{
#define CALL(N) heads_forward_no<T, N>(x, w1, w2, out1, out2, n1, n2, B, Cin, h, w, H, W, stream)
    CSPN_HEADS_DISPATCH(CALL)
#undef CALL
}

template <typename T>
int heads_backward(const T* x, const T* w1, const T* w2, const T* go1, const T* go2, T* gx, T* gw1, T* gw2, int n1, int n2, int B, int Cin, int h, int w,
                   int H, int W, void* ws, cudaStream_t stream)
{
#define CALL(N) heads_backward_no<T, N>(x, w1, w2, go1, go2, gx, gw1, gw2, n1, n2, B, Cin, h, w, H, W, (float*)ws, stream)
    CSPN_HEADS_DISPATCH(CALL)
#undef CALL
}

template int heads_forward<float>(const float*, const float*, const float*, float*, float*, int, int, int, int, int, int, int, int, cudaStream_t);
template int heads_forward<__half>(const __half*, const __half*, const __half*, __half*, __half*, int, int, int, int, int, int, int, int, cudaStream_t);
template int heads_backward<float>(const float*, const float*, const float*, const float*, const float*, float*, float*, float*, int, int, int, int, int,
                                   int, int, int, void*, cudaStream_t);
template int heads_backward<__half>(const __half*, const __half*, const __half*, const __half*, const __half*, __half*, __half*, __half*, int, int, int,
                                    int, int, int, int, int, void*, cudaStream_t);

}  // namespace cspn
