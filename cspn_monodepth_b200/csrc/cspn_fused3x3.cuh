// Fused CSPN kernels for 3x3 neighbourhoods (both reference modes): the whole T-step recurrence in ONE
// launch.  Forward (BWD = false) replaces the loops at CSPN_new.py:80-90 and CSPN_ours.py:47-53
// (pac.py:89-94 per step); backward (BWD = true) replaces autograd through those loops (pac.py:96-121).
// This header is included by cspn_fused3x3.cu (forward instantiations) and cspn_fused3x3_bwd.cu (backward).
//
// Design (B200, sm_100a) - see DESIGN.md section 3 for the derivation and the measured numbers behind it:
//  * The recurrence is FMA-bound once fused (8 FMA per pixel per step against 44 B of HBM traffic per
//    pixel in total), and only the register file can feed the FMA pipe: the 8 loop-invariant, pre-
//    normalised weights of every pixel plus the re-injection term stay in REGISTERS for all T steps
//    (the reference recomputes the normalisation every step).  A thread owns a 2-wide x P-tall strip of
//    pixels as packed f32x2 pairs so the inner loop is 8 FFMA2 per pixel pair (fma.rn.f32x2: two FMAs per
//    issue slot - leaves issue slots for the data movement).
//  * Left/right neighbours come from warp shuffles, rows above/below from the thread's own registers;
//    warps of a CTA exchange their edge rows through shared memory once per step.
//  * A register-resident CTA tile is only 64 x (NW*P) pixels, far smaller than the T-pixel dependency cone,
//    so CTAs cooperate as a thread-block CLUSTER (up to 16 CTAs, e.g. 5x3 = one whole 304x228 NYU image):
//    every second step each CTA pushes a 2-pixel-deep halo ring straight into its 8 neighbours' shared
//    memory with st.async (DSMEM) and the neighbour's mbarrier counts the bytes - no cluster-wide barrier,
//    no fence.  Only at the edge of a cluster tile that is not an image border does the classic shrinking
//    (trapezoid) halo of T pixels apply.
//  * The guidance tile (8 channels, 1-pixel apron) is staged once into shared memory by TMA
//    (cp.async.bulk.tensor, one box per channel, out-of-image elements zero-filled by the hardware = the
//    reference's ZeroPad2d), weights are normalised from there into registers; depth/sparse come in with
//    plain coalesced loads that overlap the TMA.
//  * HBM traffic is the algorithmic minimum: guidance/depth/sparse are read once (plus halo overlap that
//    L2 serves), only the final depth is written.
#pragma once
#include <cuda.h>

#include <cstddef>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>

#include "cspn_common.cuh"

namespace cspn {
namespace {

typedef unsigned long long u64;

constexpr int kTileW = 64;            // 32 lanes x 2 pixels
constexpr int kHaloX = 2;             // one lane (= one pixel pair) of halo per side between CTAs of a cluster
constexpr int kHaloY = 2;             // two rows of halo per side between CTAs of a cluster
constexpr int kPeriod = 2;            // halo refresh period in steps (= halo depth)
constexpr int kStepX = kTileW - 2 * kHaloX;   // x spacing of CTA tiles inside a cluster (60)

// ---- small PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ u64 pk(float lo, float hi)
{
    u64 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ float lo_of(u64 v)
{
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return lo;
}
__device__ __forceinline__ float hi_of(u64 v)
{
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return hi;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b)
{
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b)
{
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// 1/s to within 1 ulp without the slow-path call of __frcp_rn: MUFU.RCP + one Newton step, branch free.
// s = 0 gives inf and then NaN (0 * inf in the correction), so an all-zero weight sum still yields NaN weights.
__device__ __forceinline__ float fast_rcp(float s)
{
    float r;
    asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(s));
    const float e = fmaf(-s, r, 1.f);
    return fmaf(r, e, r);
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t cta_rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta_rank));
    return r;
}
// 8-byte store into another CTA's shared memory; the bytes are counted on that CTA's mbarrier.
__device__ __forceinline__ void st_async_b64(uint32_t remote_addr, u64 v, uint32_t remote_bar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
                 :: "r"(remote_addr), "l"(v), "r"(remote_bar) : "memory");
}
// Same, predicated: lanes with pred == false issue nothing (no divergent branch around the store).
__device__ __forceinline__ void st_async_b64_if(bool pred, uint32_t remote_addr, u64 v, uint32_t remote_bar)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t"
                 "@q st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];\n\t}"
                 :: "r"(remote_addr), "l"(v), "r"(remote_bar), "r"((uint32_t)pred) : "memory");
}
// Bulk copy of `bytes` (multiple of 16) from this CTA's shared memory into a cluster neighbour's, counted on the
// neighbour's mbarrier.  Runs on the async copy engine; the issuing thread does not wait.
__device__ __forceinline__ void bulk_s2s(uint32_t remote_dst, uint32_t local_src, uint32_t bytes, uint32_t remote_bar)
{
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(remote_dst), "r"(local_src), "r"(bytes), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
// Default (cta-scope acquire) semantics: in both uses the awaited bytes land in THIS CTA's shared memory and
// are published by the barrier's complete_tx (TMA box / neighbours' st.async), like any TMA consumer.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}"
        :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
// relaxed: the only thing published before it is mbarrier initialisation, which fence.mbarrier_init covers
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ float2 to_f32x2(float2 v) { return v; }
__device__ __forceinline__ float2 to_f32x2(__half2 v) { return __half22float2(v); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

__device__ __forceinline__ uint32_t sm_id()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
    return r;
}
// 16-byte asynchronous global -> shared copy (LDGSTS, L2 only), generic proxy; completion via cp_async_wait_all.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// (x-1,x) and (x+1,x+2) pairs of a row from its (x,x+1) pair: one shuffle + one register move each.
__device__ __forceinline__ void shifted(u64 a, u64& s1, u64& s2)
{
    const float lo = lo_of(a), hi = hi_of(a);
    const float l = __shfl_up_sync(0xffffffffu, hi, 1), r = __shfl_down_sync(0xffffffffu, lo, 1);
    s1 = pk(l, lo);
    s2 = pk(hi, r);
}

// Optional cycle trace (build with -DCSPN_TRACE): lane 0 of every warp stamps clock64() at fixed points.
#ifdef CSPN_TRACE
__device__ long long* g_trace = nullptr;
constexpr int kTraceSlots = 96;
#define TRACE(slot) do { if (g_trace && (threadIdx.x & 31) == 0) g_trace[((size_t)(blockIdx.z * gridDim.y * gridDim.x + blockIdx.y * gridDim.x + blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kTraceSlots + (slot)] = clock64(); } while (0)
#else
#define TRACE(slot) do { } while (0)
#endif

template <typename T>
struct FusedParams {
    const T* g; int64_t gbs;
    const T* depth; const T* sparse; int sparse_channels;
    T* out;
    int C, H, W, iters;
    int cx, cy;           // cluster dims (CTAs)
    int ntx, nty;         // cluster tiles per image
    int stepx, stepy;     // origin spacing of cluster tiles
    int ew, eh;           // extent of one cluster tile
    int margin;           // decaying halo at cluster-tile edges that are not image borders (= iters)
    uint32_t total_tiles; // GLB exchange only: cx * cy * planes tiles, processed by a persistent grid
    uint4* inbox;         // GLB exchange only: one Inbox per CTA tile in global memory
    uint32_t tag_base;    // GLB exchange only: tag of refresh e is tag_base + e
    int* status;          // GLB exchange only: set to 1 when a neighbour never showed up (the results are NaN-filled as well)
    // backward only
    const T* gout;        // dL/d out
    T* gg; T* gd;         // dL/d guidance [B, Cg, H, W], dL/d depth [B, 1, H, W]
    int Cg;
    int Ctot, ch0;        // backward works on depth channel ch0 of Ctot (one launch per channel; the forward takes all planes at once)
    float* hist;          // history scratch: hist_slots tiles of iters x TH x 64 floats
    int hist_slots;       // hist_by_cta: one per CTA of the launch (launches of at most kHistSlots CTAs, every stream-mode launch);
    int hist_by_cta;      // otherwise one per SM id (%smid), so that the scratch stays L2-sized however many CTAs the launch has
};

// Global-memory halo inbox of one CTA (GLB exchange), in uint4 units.  Every message is one 16-byte store
// {lo, tag, hi, tag}: data and "valid" flag travel in the same transaction (the LL idea of collective libraries),
// so the sender needs no fence and the receiver just re-reads the slot until both tags match.
//   col[parity][side][TH]  then  row[parity][side][kHaloY][32]
template <int TH> struct InboxGeom {
    static constexpr uint32_t col_par = 2 * TH, col_side = TH;
    static constexpr uint32_t row_base = 4 * TH, row_par = 2 * kHaloY * 32, row_side = kHaloY * 32;
    static constexpr uint32_t size = 4 * TH + 4 * kHaloY * 32;
};
__device__ __forceinline__ void st_ll(uint4* slot, u64 v, uint32_t tag)
{
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %2};" :: "l"(slot), "r"(__float_as_uint(lo_of(v))), "r"(tag), "r"(__float_as_uint(hi_of(v))) : "memory");
}
__device__ __forceinline__ uint4 ld_ll(const uint4* slot)
{
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(slot) : "memory");
    return v;
}

// Static part of one CTA's shared memory.  rowbuf: per-step edge rows of each warp (intra-CTA, one buffer
// per step parity).  colbox/rowbox: the halo ring received from cluster neighbours, double buffered by
// refresh parity, each with its own transaction-counting mbarrier.
template <int NW, int P>
struct __align__(128) Smem {
    float rowbuf[2][NW][2][kTileW];
    float colbox[2][2][NW * P][2];        // [parity][side: 0 left, 1 right][tile row][2 px]
    float rowbox[2][2][kHaloY][kTileW];   // [parity][side: 0 top, 1 bottom][halo row][tile x]
    u64 colstage[2][2][NW * P];           // [parity][side][tile row]: rim columns on their way to the left / right neighbour
    u64 halo_bar[2];
    u64 tma_bar[8];
};

// Geometry of the TMA staging buffer for the guidance tile.  The innermost start coordinate of a TMA box must
// be 16-byte aligned in global memory (measured: anything else traps as an illegal instruction), so the box
// starts at the aligned column at or left of (tile x - apron) and is wide enough for every sub-offset.
template <typename T, int TH, int MODE>
struct Stage {
    static constexpr int apron = MODE == CSPN_MODE_NEW ? 1 : 0;       // mode NEW gathers weights from the 8 neighbours
    static constexpr int rows = TH + 2 * apron;
    static constexpr int align = 16 / (int)sizeof(T);                // elements per 16 bytes
    static constexpr int cols = ((kTileW + 2 * apron + align - 1 + align - 1) / align) * align;   // fp32: 72 | 68, fp16: 80 | 72
    static constexpr int box_bytes = rows * cols * (int)sizeof(T);   // what one TMA box delivers
    static constexpr int plane = ((box_bytes + 127) / 128) * 128 / (int)sizeof(T);   // elements per channel, 128-byte aligned for TMA
    static constexpr size_t bytes = (size_t)8 * plane * sizeof(T);
    // aligned box start for a tile whose first pixel column is ox (floor division, ox - apron may be negative)
    __host__ __device__ static int box_x(int ox) { const int v = ox - apron; return (v >= 0 ? v / align : -((-v + align - 1) / align)) * align; }
};

// Shared-memory tiles of the backward pass (they alias the TMA staging buffer, which is dead after the prologue):
//   wt [8][TH][64]  transposed weights wt_j(q) = n'_{7-j}(q + o_j): the adjoint sweep is then the same gather as the forward
//   rt [2][TH][64]  r^t tile of the current / next reverse step, fetched from the history scratch with cp.async
//   stash (placed behind both the staging buffer and the tiles, written by the prologue, read by the epilogue; mode NEW):
//   sinv [TH][64]   1 / S(p);   sgn [TH][32]  per pixel pair: bit j / 8+j = raw guidance of tap j negative (lo / hi pixel),
//                                             bit 16+j / 24+j = raw guidance of tap j exactly zero
template <int TH> struct BwdTiles {
    static constexpr size_t wt_floats = (size_t)8 * TH * kTileW, rt_floats = (size_t)TH * kTileW;
    static constexpr size_t bytes = (wt_floats + 2 * rt_floats) * sizeof(float);
    static constexpr size_t stash_bytes = (size_t)TH * kTileW * sizeof(float) + (size_t)TH * 32 * sizeof(uint32_t);
};

// One CTA tile from prologue to epilogue.  (bx, by, bz) = position of the tile in the (gdx, gdy, planes) grid of CTA
// tiles - the launch grid itself for hardware clusters, the tile counter of the persistent loop for the global-memory
// exchange; use = how many tiles this CTA has processed before (phase parity of the re-used TMA barriers).
template <typename T, int P, int NW, int MODE, bool TMA, bool GLB, bool BWD, bool HYB = false>
__device__ __forceinline__ void fused_tile(const FusedParams<T>& p, const CUtensorMap* gmap, unsigned char* smem_raw,
                                           const uint32_t bx, const uint32_t by, const uint32_t bz, const uint32_t gdx, const uint32_t gdy, const uint32_t use)
{
    constexpr int TH = NW * P;
    constexpr int STEPY = TH - 2 * kHaloY;
    using St = Stage<T, TH, MODE>;
    Smem<NW, P>& sm = *reinterpret_cast<Smem<NW, P>*>(smem_raw);
    const T* stage = reinterpret_cast<const T*>(smem_raw + sizeof(Smem<NW, P>));
    constexpr size_t kStashOff = sizeof(Smem<NW, P>) + ((TMA ? St::bytes : 0) > BwdTiles<TH>::bytes ? (TMA ? St::bytes : 0) : BwdTiles<TH>::bytes);
    float* sinv = reinterpret_cast<float*>(smem_raw + kStashOff);                       // backward only
    uint32_t* sgn = reinterpret_cast<uint32_t*>(sinv + TH * kTileW);

    TRACE(0);
    const bool multi = p.cx * p.cy > 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ccx = bx % p.cx, ccy = by % p.cy;
    const int tix = bx / p.cx, tiy = by / p.cy;
    const int plane = bz, b = plane / p.C, ch = BWD ? p.ch0 : plane - b * p.C;
    const size_t dplane = BWD ? (size_t)b * p.Ctot + p.ch0 : (size_t)plane;      // plane of depth / sparse / gradients this tile works on
    const bool has_left = ccx > 0, has_right = ccx < p.cx - 1, has_up = ccy > 0, has_down = ccy < p.cy - 1;
    const int H = p.H, W = p.W;
    const size_t hw = (size_t)H * W;

    // image coordinates of this CTA tile / this thread's strip: columns gx, gx+1, rows gy0 .. gy0+P-1
    const int ox = tix * p.stepx + ccx * kStepX, oy = tiy * p.stepy + ccy * STEPY;
    const int gx = ox + 2 * lane;
    const int gy0 = oy + warp * P;

    // ---- TMA issue (one thread): 8 boxes, one per guidance channel, each on its own barrier --------------------
    if (TMA && threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t bar = smem_u32(&sm.tma_bar[k]);
            mbar_arrive_expect_tx(bar, (uint32_t)St::box_bytes);
            tma_load_4d(smem_u32(stage + (size_t)k * St::plane), gmap, bar, St::box_x(ox), oy - St::apron, k, b);
        }
    }
    TRACE(1);
    const bool hw_cluster = multi && !GLB;
    // HYB (forward only): hardware clusters of (cx, 1) CTAs - one per tile row of an image.  Left / right neighbours talk through
    // DSMEM like any cluster, the tile rows above / below (other clusters) through the global-memory inboxes of the stream
    // transport; the row rims leave straight from the sweep, as soon as they are final, so that their L2 trip overlaps the rest
    // of the step.  Every cluster of the launch must be resident at once (planner: clusters <= co-resident clusters of cx).
    static_assert(!HYB || (!GLB && !BWD), "the hybrid transport is a forward cluster variant");
    constexpr bool GLBX = GLB || HYB;                    // some messages travel through the global inboxes

    const T* db = p.depth + dplane * hw;
    const T* sb = p.sparse ? p.sparse + ((size_t)b * p.sparse_channels + (p.sparse_channels == 1 ? 0 : ch)) * hw : nullptr;

    // ---- prologue: loop-invariant weights n'_j = (1-m) * n_j, re-injection c = m*d0, r^0 = d0 -------------
    // internal tap order j: (dy,dx) row-major without the centre; mode NEW channel k = 7 - j reads the
    // guidance AT THE NEIGHBOUR p + o (CSPN_new.py:43-67), mode OURS channel j reads it at p (CSPN_ours.py:37-41).
    u64 nw[P][8], cc[P], A[P];
    uint32_t sbits[BWD ? P : 1];
    if (BWD) {
#pragma unroll
        for (int i = 0; i < P; ++i) sbits[i] = 0u;
    }
    const bool x_in0 = gx >= 0 && gx < W, x_in1 = gx + 1 >= 0 && gx + 1 < W;
    const bool vec_ok = (W & 1) == 0;        // pairs start at even x: 8-byte (fp32) / 4-byte (fp16) aligned when W is even

    // depth and sparse first: plain coalesced loads, all issued back to back (clamped addresses, no branches
    // between them) so that they are in flight together while the TMA boxes land
    if (vec_ok) {
        typedef typename std::conditional<sizeof(T) == 4, float2, __half2>::type V2;
        const int cgx = min(max(gx, 0), W - 2);
        V2 dv[P], sv[P];
#pragma unroll
        for (int i = 0; i < P; ++i) {
            const size_t off = (size_t)min(max(gy0 + i, 0), H - 1) * W + cgx;
            dv[i] = *reinterpret_cast<const V2*>(db + off);
            if (sb) sv[i] = *reinterpret_cast<const V2*>(sb + off);
        }
#pragma unroll
        for (int i = 0; i < P; ++i) {
            const int gy = gy0 + i;
            const bool in = gy >= 0 && gy < H && x_in0;        // W even and gx even: both pixels of the pair are in or out together
            const float2 d = to_f32x2(dv[i]);
            float2 m = make_float2(0.f, 0.f);
            if (sb) { const float2 sp2 = to_f32x2(sv[i]); m = make_float2(signf(sp2.x), signf(sp2.y)); }
            A[i] = in ? pk(d.x, d.y) : 0ull;
            cc[i] = in ? pk(m.x, m.y) : 0ull;               // holds the mask until the weights are folded below
        }
    } else {
#pragma unroll
        for (int i = 0; i < P; ++i) {
            const int gy = gy0 + i;
            const bool row_in = gy >= 0 && gy < H;
            float d0 = 0.f, d1 = 0.f, m0 = 0.f, m1 = 0.f;
            const size_t off = (size_t)gy * W + gx;
            if (row_in && x_in0) { d0 = to_f32(db[off]); if (sb) m0 = signf(to_f32(sb[off])); }
            if (row_in && x_in1) { d1 = to_f32(db[off + 1]); if (sb) m1 = signf(to_f32(sb[off + 1])); }
            A[i] = pk(d0, d1);
            cc[i] = pk(m0, m1);
        }
    }

    TRACE(2);
    // raw guidance values -> registers
    if (TMA) {
        const int x_off = ox - St::box_x(ox);           // column of the tile's first pixel inside the staged box
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            mbar_wait(smem_u32(&sm.tma_bar[k]), use & 1u);
            TRACE(3 + k);
            const T* sp = stage + (size_t)k * St::plane;
            if (MODE == CSPN_MODE_NEW) {
                const int j = 7 - k, jj = j < 4 ? j : j + 1;
                const int dy = jj / 3 - 1, dx = jj % 3 - 1;
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    const T* src = sp + (warp * P + i + 1 + dy) * St::cols + x_off + 2 * lane + dx;
                    const float v0 = to_f32(src[0]), v1 = to_f32(src[1]);
                    nw[i][j] = pk(fabsf(v0), fabsf(v1));                               // zero-filled outside the image
                    if (BWD) sbits[i] |= (v0 < 0.f ? 1u << j : 0u) | (v1 < 0.f ? 256u << j : 0u) | (v0 == 0.f ? 65536u << j : 0u) | (v1 == 0.f ? 16777216u << j : 0u);
                }
            } else {
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    const T* src = sp + (warp * P + i) * St::cols + x_off + 2 * lane;
                    nw[i][k] = pk(to_f32(src[0]), to_f32(src[1]));
                }
            }
        }
    } else {
        const T* gb = p.g + (size_t)b * p.gbs;
#pragma unroll
        for (int i = 0; i < P; ++i) {
            const int gy = gy0 + i;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int jj = j < 4 ? j : j + 1;
                const int dy = MODE == CSPN_MODE_NEW ? jj / 3 - 1 : 0, dx = MODE == CSPN_MODE_NEW ? jj % 3 - 1 : 0;
                const int k = MODE == CSPN_MODE_NEW ? 7 - j : j, yy = gy + dy, xx = gx + dx;
                const bool rin = yy >= 0 && yy < H;
                const T* src = gb + (size_t)k * hw + (size_t)yy * W + xx;
                float v0 = (rin && xx >= 0 && xx < W) ? to_f32(src[0]) : 0.f;
                float v1 = (rin && xx + 1 >= 0 && xx + 1 < W) ? to_f32(src[1]) : 0.f;
                if (MODE == CSPN_MODE_NEW) {
                    if (BWD) sbits[i] |= (v0 < 0.f ? 1u << j : 0u) | (v1 < 0.f ? 256u << j : 0u) | (v0 == 0.f ? 65536u << j : 0u) | (v1 == 0.f ? 16777216u << j : 0u);
                    v0 = fabsf(v0); v1 = fabsf(v1);
                }
                nw[i][j] = pk(v0, v1);
            }
        }
    }

    TRACE(11);
    // normalise, fold the sparse mask in (packed f32x2 arithmetic, all rows' reductions independent)
    if (MODE == CSPN_MODE_NEW) {
        u64 scale[P];
#pragma unroll
        for (int i = 0; i < P; ++i) {
            u64 sum = nw[i][7];                                       // reference order k = 0..7 (CSPN_new.py:124), k = 7 - j
#pragma unroll
            for (int k = 1; k < 8; ++k) sum = add2(sum, nw[i][7 - k]);
            const int gy = gy0 + i;
            const bool row_in = gy >= 0 && gy < H;
            // n'_j = (1-m) * W_j / S.  S = 0 -> inf -> 0*inf = NaN like the reference's 0/0.  Pixels outside the
            // image are virtual: exactly zero weights and value (the reference's zero padding).
            const float r0 = fast_rcp(lo_of(sum)), r1 = fast_rcp(hi_of(sum));
            if (BWD) {
                *reinterpret_cast<u64*>(sinv + (warp * P + i) * kTileW + 2 * lane) = pk(r0, r1);
                sgn[(warp * P + i) * 32 + lane] = sbits[i];
            }
            const float f0 = (row_in && x_in0) ? (1.f - lo_of(cc[i])) * r0 : 0.f;
            const float f1 = (row_in && x_in1) ? (1.f - hi_of(cc[i])) * r1 : 0.f;
            scale[i] = pk(f0, f1);
        }
#pragma unroll
        for (int i = 0; i < P; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) nw[i][j] = mul2(nw[i][j], scale[i]);
            cc[i] = mul2(cc[i], A[i]);                                // c = m * d0
        }
    } else {
#pragma unroll
        for (int i = 0; i < P; ++i) {
            const int gy = gy0 + i;
            const bool row_in = gy >= 0 && gy < H;
            const bool in0 = row_in && x_in0, in1 = row_in && x_in1;
            float w0[8], w1[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { w0[j] = lo_of(nw[i][j]); w1[j] = hi_of(nw[i][j]); }
            float m0 = w0[0], m1 = w1[0];
#pragma unroll
            for (int j = 1; j < 8; ++j) { m0 = fmaxf(m0, w0[j]); m1 = fmaxf(m1, w1[j]); }
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) { w0[j] = expf(w0[j] - m0); w1[j] = expf(w1[j] - m1); s0 += w0[j]; s1 += w1[j]; }
            const float f0 = in0 ? (1.f - lo_of(cc[i])) * fast_rcp(s0) : 0.f, f1 = in1 ? (1.f - hi_of(cc[i])) * fast_rcp(s1) : 0.f;
            // taps that read the zero padding contribute n_j * 0 (no border renormalisation, pac.py:89): drop
            // their weight instead, so that whatever a tile-edge shuffle delivers for them is multiplied by 0
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int jj = j < 4 ? j : j + 1;
                const int yy = gy + jj / 3 - 1, xx = gx + jj % 3 - 1;
                const bool rin = yy >= 0 && yy < H;
                nw[i][j] = pk((rin && xx >= 0 && xx < W) ? w0[j] * f0 : 0.f, (rin && xx + 1 >= 0 && xx + 1 < W) ? w1[j] * f1 : 0.f);
            }
            cc[i] = mul2(cc[i], A[i]);
        }
    }

    // which lanes / rows of this CTA tile are authoritative (not halo owned by a cluster neighbour)
    const int lx0 = has_left ? 1 : 0, lx1 = has_right ? 30 : 31;
    const int ry0 = has_up ? kHaloY : 0, ry1 = has_down ? TH - 1 - kHaloY : TH - 1;
    const bool lane_auth = lane >= lx0 && lane <= lx1;
    const uint32_t my_rank = HYB ? (uint32_t)ccx : (uint32_t)(ccx + ccy * p.cx);      // == %cluster_ctarank for a (cx, cy, 1) [HYB: (cx, 1, 1)] cluster
    // bytes this CTA receives per refresh: 8 per row from the left / right neighbour, whole rows (all 32 lanes; halo
    // lanes of a row are overridden by the column boxes) from above / below, 2x8 from each diagonal neighbour
    const uint32_t expect_bytes = HYB ? 8u * (uint32_t)(((has_left ? 1 : 0) + (has_right ? 1 : 0)) * (ry1 - ry0 + 1))      // rows and corners arrive through global memory
                                      : 8u * (uint32_t)(((has_left ? 1 : 0) + (has_right ? 1 : 0)) * (ry1 - ry0 + 1) +
                                                        ((has_up ? 1 : 0) + (has_down ? 1 : 0)) * kHaloY * 32 +
                                                        kHaloY * (((has_up && has_left) ? 1 : 0) + ((has_up && has_right) ? 1 : 0) +
                                                                  ((has_down && has_left) ? 1 : 0) + ((has_down && has_right) ? 1 : 0)));
    const uint32_t sm_base = smem_u32(&sm);
    using SmemT = Smem<NW, P>;
    using IG = InboxGeom<TH>;
    constexpr uint32_t kColbox = offsetof(SmemT, colbox), kRowbox = offsetof(SmemT, rowbox), kHaloBar = offsetof(SmemT, halo_bar);
    constexpr uint32_t kColPar = sizeof(float) * 2 * TH * 2, kColSide = sizeof(float) * TH * 2;
    constexpr uint32_t kRowPar = sizeof(float) * 2 * kHaloY * kTileW, kRowSide = sizeof(float) * kHaloY * kTileW;

    // The re-injection term c = m*d0 moves out of the register file: it is read once per row and step, while the 20
    // registers it occupied let the compiler keep the loop invariants of the halo exchange instead of rebuilding them
    // every step.  It lives where the guidance was staged (dead from here on; each thread only re-reads its own words).
    u64* const ctile = reinterpret_cast<u64*>(smem_raw + sizeof(Smem<NW, P>)) + (warp * P) * 32 + lane;
    __syncthreads();                                     // every warp has taken its guidance out of the staging buffer
#pragma unroll
    for (int i = 0; i < P; ++i) ctile[i * 32] = cc[i];

    TRACE(12);
    if (hw_cluster) cluster_wait();
    TRACE(13);

    // ---- halo-exchange bookkeeping ------------------------------------------------------------------------
    const bool push_l = multi && lane == 1 && has_left, push_r = multi && lane == 30 && has_right;
    // Per-lane message of the column push (parity 0 addresses; msg_dst == 0: this lane sends nothing):
    //   lanes 0..P-1      row i of the left rim column  -> left neighbour's right box
    //   lanes P..2P-1     row i of the right rim column -> right neighbour's left box
    //   lanes 2P..2P+3    (top warp)    rows kHaloY.. of the rim columns -> upper-left / upper-right neighbour, bottom box rows
    //   lanes 2P+4..2P+7  (bottom warp) rows TH-2*kHaloY.. of the rim columns -> lower-left / lower-right neighbour, top box rows
    static_assert(2 * P + 4 * kHaloY <= 32, "column push needs one lane per message");
    uint32_t msg_dst = 0u, msg_src = 0u, msg_bar = 0u;
    if (multi) {
        const int ty0 = warp * P;
        int side = -1, srow = 0, drow = 0, dcy = 0;               // side: 0 = my left rim column, 1 = my right rim column
        if (lane < 2 * P) {
            side = lane / P; srow = drow = ty0 + lane % P;
            if (srow < ry0 || srow > ry1) side = -1;
        } else if (HYB) {
            // corner messages leave with the row rims (ship_row)
        } else if (lane < 2 * P + 2 * kHaloY) {
            if (warp == 0 && has_up) { side = (lane - 2 * P) / kHaloY; const int h = (lane - 2 * P) % kHaloY; srow = kHaloY + h; drow = TH - kHaloY + h; dcy = -1; }
        } else if (lane < 2 * P + 4 * kHaloY) {
            if (warp == NW - 1 && has_down) { side = (lane - 2 * P - 2 * kHaloY) / kHaloY; const int h = (lane - 2 * P - 2 * kHaloY) % kHaloY; srow = TH - 2 * kHaloY + h; drow = h; dcy = 1; }
        }
        if (side == 0 && !has_left) side = -1;
        if (side == 1 && !has_right) side = -1;
        if (side >= 0) {
            msg_src = (uint32_t)(side * TH + srow) * 8u;
            if (GLB) {
                const uint32_t nblk = (bz * gdy + by + dcy) * gdx + bx + (side == 0 ? -1 : 1);
                msg_dst = nblk * IG::size + (side == 0 ? IG::col_side : 0u) + (uint32_t)drow;   // uint4 index, parity 0; never 0
            } else {
                const uint32_t nb = mapa(sm_base, my_rank + dcy * p.cx + (side == 0 ? -1 : 1));
                msg_dst = nb + kColbox + (side == 0 ? kColSide : 0u) + 8u * drow;      // my left rim lands in the neighbour's RIGHT box
                msg_bar = nb + kHaloBar;
            }
        }
    }
    const uint32_t my_blk = (bz * gdy + by) * gdx + bx;

    auto exchange_rows = [&](int par, u64& top, u64& bot) {
        // publish this warp's edge rows for the warps above / below (same CTA), fetch theirs
        *reinterpret_cast<u64*>(&sm.rowbuf[par][warp][0][2 * lane]) = A[0];
        *reinterpret_cast<u64*>(&sm.rowbuf[par][warp][1][2 * lane]) = A[P - 1];
        __syncthreads();
        top = 0ull; bot = 0ull;   // rows -1 and P of this strip (zero above/below the CTA tile)
        if (warp > 0) top = *reinterpret_cast<const u64*>(&sm.rowbuf[par][warp - 1][1][2 * lane]);
        if (warp < NW - 1) bot = *reinterpret_cast<const u64*>(&sm.rowbuf[par][warp + 1][0][2 * lane]);
    };

    // GLB exchange: this warp's slots of the global inbox.  Lanes 0..P+1: left column rows ty0-1..ty0+P, lanes
    // 16..16+P+1: right column; top / bottom warp: both halo rows, one slot per lane.
    struct Polled { uint4 c, r0, r1; };
    const int poll_sd = lane >> 4, poll_row = warp * P - 1 + (lane & 15);
    // HYB: the column boxes are filled through DSMEM except for their halo rows (corner messages from the clusters above / below)
    const bool poll_want = (lane & 15) < P + 2 && poll_row >= 0 && poll_row < TH && (poll_sd == 0 ? has_left : has_right) &&
                           (!HYB || poll_row < ry0 || poll_row > ry1);
    const bool poll_r0 = has_up && warp == 0, poll_r1 = has_down && warp == NW - 1;
    // One read of the slots of refresh epoch e.  Issued BEFORE the intra-CTA row exchange of the step so that the L2
    // round trip overlaps that barrier; take_refresh only re-reads what was not current yet.
    auto poll = [&](int e) {
        const uint32_t tag = p.tag_base + (uint32_t)e;
        const int rpar = e & 1;
        Polled q;
        q.c = make_uint4(0, tag, 0, tag); q.r0 = q.c; q.r1 = q.c;
        if (GLBX) {
            const uint4* box = p.inbox + (size_t)my_blk * IG::size;
            if (poll_want) q.c = ld_ll(box + rpar * IG::col_par + poll_sd * IG::col_side + poll_row);
            if (poll_r0 || poll_r1) {
                const uint4* rslot = box + IG::row_base + rpar * IG::row_par + (poll_r1 ? IG::row_side : 0u) + lane;
                q.r0 = ld_ll(rslot); q.r1 = ld_ll(rslot + 32);
            }
        }
        return q;
    };
    // Ship the halo ring of refresh parity rpar to the neighbours: the warp's rim columns were staged into
    // sm.colstage by lanes 1 / 30 during the sweep; now lane m sends message m (one 8-byte st.async each, a single
    // warp-wide instruction), rim rows go out directly from all 32 lanes of the top / bottom warp.
    auto ship = [&](int rpar, uint32_t tag) {
        (void)tag;
        const uint32_t bar_off = kHaloBar + 8u * rpar;
        TRACE(90);
        __syncwarp();
        if (msg_dst != 0u) {
            const u64 v = *reinterpret_cast<const u64*>(reinterpret_cast<const unsigned char*>(&sm.colstage[rpar][0][0]) + msg_src);
            if (GLB) st_ll(p.inbox + msg_dst + rpar * IG::col_par, v, tag);
            else st_async_b64(msg_dst + rpar * kColPar, v, msg_bar + 8u * rpar);
        }
        TRACE(91);
        if (HYB) return;                                                             // the row rims left during the sweep (ship_row)
        if (has_up && warp == 0) {                                                   // warp-uniform
            // my tile rows kHaloY .. 2*kHaloY-1 are the upper neighbour's bottom halo rows
            if (GLB) {
                uint4* d = p.inbox + (size_t)(my_blk - gdx) * IG::size + IG::row_base + rpar * IG::row_par + IG::row_side + lane;
#pragma unroll
                for (int h = 0; h < kHaloY; ++h) st_ll(d + h * 32, A[kHaloY + h], tag);
            } else {
                const uint32_t d = mapa(sm_base, my_rank - p.cx);
#pragma unroll
                for (int h = 0; h < kHaloY; ++h)
                    st_async_b64(d + kRowbox + rpar * kRowPar + kRowSide + 4u * (h * kTileW + 2 * lane), A[kHaloY + h], d + bar_off);
            }
        }
        if (has_down && warp == NW - 1) {
            if (GLB) {
                uint4* d = p.inbox + (size_t)(my_blk + gdx) * IG::size + IG::row_base + rpar * IG::row_par + lane;
#pragma unroll
                for (int h = 0; h < kHaloY; ++h) st_ll(d + h * 32, A[P - 2 * kHaloY + h], tag);
            } else {
                const uint32_t d = mapa(sm_base, my_rank + p.cx);
#pragma unroll
                for (int h = 0; h < kHaloY; ++h)
                    st_async_b64(d + kRowbox + rpar * kRowPar + 4u * (h * kTileW + 2 * lane), A[P - 2 * kHaloY + h], d + bar_off);
            }
        }
        TRACE(92);
    };

    // HYB: tile row i of this warp just became final inside the sweep; if it is a rim row of the tile, it goes to the inbox of the
    // tile above / below right away (i is a compile-time constant at every call site, the warp tests are uniform).
    const bool ship_up = HYB && has_up && warp == 0, ship_down = HYB && has_down && warp == NW - 1;
    static_assert(!HYB || NW >= 2, "top and bottom rim rows must live in different warps");
    // destinations fixed for the whole loop (parity 0): this lane's slot of the row box of the tile above / below, and - lanes 1 / 30 -
    // of the column box of the diagonal neighbour (the rim columns of the same rows are the corners of its ring)
    uint4* row_dst = nullptr;
    uint4* cor_dst = nullptr;
    if (ship_up || ship_down) {
        const uint32_t nb = ship_up ? my_blk - gdx : my_blk + gdx;
        row_dst = p.inbox + (size_t)nb * IG::size + IG::row_base + (ship_up ? IG::row_side : 0u) + lane;
        if (push_l) cor_dst = p.inbox + (size_t)(nb - 1) * IG::size + IG::col_side + (ship_up ? TH - kHaloY : 0);
        if (push_r) cor_dst = p.inbox + (size_t)(nb + 1) * IG::size + (ship_up ? TH - kHaloY : 0);
    }
    auto ship_row = [&](int i, u64 a, int rpar, uint32_t tag) {
        const bool top = i >= kHaloY && i < 2 * kHaloY, btm = i >= P - 2 * kHaloY && i < P - kHaloY;      // compile-time at every call site
        if (!top && !btm) return;
        if (top && btm) {                                      // short strips: the same row index is a top rim row in warp 0 and a bottom one in warp NW-1
            const int h = ship_up ? i - kHaloY : i - (P - 2 * kHaloY);
            if (row_dst) st_ll(row_dst + rpar * IG::row_par + h * 32, a, tag);
            if (cor_dst) st_ll(cor_dst + rpar * IG::col_par + h, a, tag);
        } else if (top ? ship_up : ship_down) {
            const int h = top ? i - kHaloY : i - (P - 2 * kHaloY);
            st_ll(row_dst + rpar * IG::row_par + h * 32, a, tag);
            if (cor_dst) st_ll(cor_dst + rpar * IG::col_par + h, a, tag);
        }
    };

    // One step, r'(p) = c(p) + sum_j n'_j(p) * r(p + o_j), organised by SOURCE row: row r's three pixel pairs
    // (x-1,x) / (x,x+1) / (x+1,x+2) are formed once and scattered into the accumulators of output rows
    // r+1 (taps 0-2), r (taps 3,4) and r-1 (taps 5-7).  Three independent FMA chains are in flight, only one
    // row of shifted pairs is live, and output row r-1 completes exactly when old row r-1 is dead, so the
    // update is in place (no second copy of the strip).  PUSH: also send finished rim values to the neighbours.
    auto compute_step = [&](auto push_tag, u64 top, u64 bot, int rpar, uint32_t tag) {
        constexpr bool PUSH = decltype(push_tag)::value;
#ifndef CSPN_PACKED_SWEEP
        // Forward kernel: scalar-FMA form of the same sweep (bit-identical: every lane of an FFMA2 is an IEEE FMA).
        // FFMA2 needs its operands in aligned register pairs, and building the shifted pairs and moving finished rows
        // into place costs ~0.9 MOV per FFMA2; scalar FMAs read the halves where they are and run six independent
        // chains instead of three.  Measured: 2 % faster for the forward kernel, 5-10 % slower inside the backward
        // kernel (both its recompute phase and its reverse step), which therefore keeps the packed form below.
        if constexpr (!BWD) {
        float a0[P], a1[P];
#pragma unroll
        for (int r = -1; r <= P; ++r) {
            const u64 src = r < 0 ? top : (r < P ? A[r < 0 ? 0 : (r < P ? r : 0)] : bot);
            const float lo = lo_of(src), hi = hi_of(src);
            const float l = __shfl_up_sync(0xffffffffu, hi, 1), rr = __shfl_down_sync(0xffffffffu, lo, 1);
            if (r + 1 < P) {
                const int i = r + 1;
                const u64 ci = ctile[i * 32];
                float x0 = lo_of(ci), x1 = hi_of(ci);
                x0 = fmaf(lo_of(nw[i][0]), l, x0);   x1 = fmaf(hi_of(nw[i][0]), lo, x1);
                x0 = fmaf(lo_of(nw[i][1]), lo, x0);  x1 = fmaf(hi_of(nw[i][1]), hi, x1);
                a0[i] = fmaf(lo_of(nw[i][2]), hi, x0); a1[i] = fmaf(hi_of(nw[i][2]), rr, x1);
            }
            if (r >= 0 && r < P) {
                const int i = r < 0 ? 0 : (r < P ? r : 0);
                a0[i] = fmaf(lo_of(nw[i][4]), hi, fmaf(lo_of(nw[i][3]), l, a0[i]));
                a1[i] = fmaf(hi_of(nw[i][4]), rr, fmaf(hi_of(nw[i][3]), lo, a1[i]));
            }
            if (r >= 1) {
                const int i = r - 1;
                float x0 = a0[i], x1 = a1[i];
                x0 = fmaf(lo_of(nw[i][5]), l, x0);   x1 = fmaf(hi_of(nw[i][5]), lo, x1);
                x0 = fmaf(lo_of(nw[i][6]), lo, x0);  x1 = fmaf(hi_of(nw[i][6]), hi, x1);
                x0 = fmaf(lo_of(nw[i][7]), hi, x0);  x1 = fmaf(hi_of(nw[i][7]), rr, x1);
                const u64 a = pk(x0, x1);
                A[i] = a;
                if (PUSH && HYB) ship_row(i, a, rpar, tag);
                if (PUSH) {
                    const int ty = warp * P + i;
                    if (push_l) sm.colstage[rpar][0][ty] = a;
                    if (push_r) sm.colstage[rpar][1][ty] = a;
                }
            }
        }
        if (PUSH) ship(rpar, tag);
        return;
        }
#endif
        u64 acc[P];
#pragma unroll
        for (int r = -1; r <= P; ++r) {
            const u64 src = r < 0 ? top : (r < P ? A[r < 0 ? 0 : (r < P ? r : 0)] : bot);
            u64 s1, s2;
            shifted(src, s1, s2);
            if (r + 1 < P) {
                u64 a = ctile[(r + 1) * 32];
                a = fma2(nw[r + 1][0], s1, a);
                a = fma2(nw[r + 1][1], src, a);
                acc[r + 1] = fma2(nw[r + 1][2], s2, a);
            }
            if (r >= 0 && r < P) {
                u64 a = acc[r];
                a = fma2(nw[r][3], s1, a);
                acc[r] = fma2(nw[r][4], s2, a);
            }
            if (r >= 1) {
                const int i = r - 1;
                u64 a = acc[i];
                a = fma2(nw[i][5], s1, a);
                a = fma2(nw[i][6], src, a);
                a = fma2(nw[i][7], s2, a);
                A[i] = a;
                if (PUSH && HYB) ship_row(i, a, rpar, tag);
                if (PUSH) {
                    // stage the rim values locally (predicated stores, no branches); shipped in bulk after the sweep
                    const int ty = warp * P + i;
                    if (push_l) sm.colstage[rpar][0][ty] = a;
                    if (push_r) sm.colstage[rpar][1][ty] = a;
                }
            }
        }
        if (PUSH) ship(rpar, tag);
    };

    // Take the refreshed halo ring of refresh epoch e (pushed by the neighbours during their previous step).
    bool poisoned = false;
#ifdef CSPN_TRACE
    int repolls = 0;
#endif
    auto take_refresh = [&](int e, Polled q, u64& top, u64& bot) {
        const int rpar = e & 1;
        if (GLBX) {
            // wait until every tag is current, then drop the payload into the same shared-memory boxes the DSMEM path fills
            const uint32_t tag = p.tag_base + (uint32_t)e;
            const int sd = poll_sd, row = poll_row;
            const bool want = poll_want, wr0 = poll_r0, wr1 = poll_r1;
            for (int spin = 0; !poisoned; ++spin) {
                const bool ok = q.c.y == tag && q.c.w == tag && q.r0.y == tag && q.r0.w == tag && q.r1.y == tag && q.r1.w == tag;
                if (__all_sync(0xffffffffu, ok)) break;
#ifdef CSPN_TRACE
                ++repolls;
#endif
                if (spin > (1 << 21)) { poisoned = true; break; }     // neighbours never showed up (grid not co-resident?): fail loudly, do not hang
                q = poll(e);
            }
            const uint4 c = q.c, r0 = q.r0, r1 = q.r1;
            // Note for racecheck: adjacent warps both fetch the two rows on their common boundary (each needs the row above
            // and below its strip) and both store the same 8 bytes to the same column-box word; a warp only reads back what
            // it stored itself (after the __syncwarp below).  compute-sanitizer flags this write-write overlap; three
            // overlap-free variants (private slots, shuffles, edge lanes polling their own rows) were racecheck-clean but
            // measured 5 % slower end to end (30.2-30.7 vs 28.8 us on the headline), so the overlap stays.
            if (want) *reinterpret_cast<uint2*>(&sm.colbox[rpar][sd][row][0]) = make_uint2(c.x, c.z);
            if (wr0 || wr1) {
                *reinterpret_cast<uint2*>(&sm.rowbox[rpar][wr1 ? 1 : 0][0][2 * lane]) = make_uint2(r0.x, r0.z);
                *reinterpret_cast<uint2*>(&sm.rowbox[rpar][wr1 ? 1 : 0][1][2 * lane]) = make_uint2(r1.x, r1.z);
            }
            __syncwarp();
        }
        TRACE(93);
        if (!GLB) mbar_wait(sm_base + kHaloBar + 8u * rpar, (uint32_t)(((e - 1) >> 1) & 1));
        TRACE(94);
        if (has_up && warp == 0) {
#pragma unroll
            for (int h = 0; h < kHaloY; ++h) A[h] = *reinterpret_cast<const u64*>(&sm.rowbox[rpar][0][h][2 * lane]);
        }
        if (has_down && warp == NW - 1) {
#pragma unroll
            for (int h = 0; h < kHaloY; ++h) A[P - kHaloY + h] = *reinterpret_cast<const u64*>(&sm.rowbox[rpar][1][h][2 * lane]);
        }
        const bool edge = (lane == 0 && has_left) || (lane == 31 && has_right);
        if (edge) {
            const int side = lane == 0 ? 0 : 1;
            const int ty0 = warp * P;
#pragma unroll
            for (int i = 0; i < P; ++i) A[i] = *reinterpret_cast<const u64*>(&sm.colbox[rpar][side][ty0 + i][0]);
            top = ty0 > 0 ? *reinterpret_cast<const u64*>(&sm.colbox[rpar][side][ty0 - 1][0]) : 0ull;
            bot = ty0 + P < TH ? *reinterpret_cast<const u64*>(&sm.colbox[rpar][side][ty0 + P][0]) : 0ull;
        }
    };

    // Backward only: every r^t (t = 0 .. T-1) of this CTA's tile goes to a scratch tile in global memory that is
    // private to the SM this CTA runs on (one CTA per SM), so the scratch stays resident in L2 however large the
    // problem is.  The reverse phase reads the tiles back with cp.async.
    float* hist_cta = nullptr;
    if (BWD) {
        const uint32_t slot = p.hist_by_cta ? (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x : sm_id();
        if (slot >= (uint32_t)p.hist_slots) poisoned = true;          // unknown SM numbering: fail loudly (NaN gradients)
        hist_cta = p.hist + (size_t)(slot % (uint32_t)p.hist_slots) * (size_t)p.iters * (TH * kTileW);
    }
    auto dump_hist = [&](int t) {
        if (BWD) {
            float* h = hist_cta + (size_t)t * (TH * kTileW) + (warp * P) * kTileW + 2 * lane;
#pragma unroll
            for (int i = 0; i < P; ++i) *reinterpret_cast<u64*>(h + i * kTileW) = A[i];
        }
    };

    // ---- T propagation steps ---------------------------------------------------------------------------
    // Steps come in pairs.  The halo ring received from cluster neighbours is 2 pixels deep, so it is refreshed
    // at the start of every even step t >= 2; the values it needs (r^t on the rim of each CTA's authoritative
    // region) are final during the odd step t-1 and are pushed at the end of that step.
    // Backward: the same loop recomputes r^1 .. r^{T-1} (r^T is not needed) and records every r^t.
    for (int t = 0; t < p.iters; t += 2) {
        // ===== even step t =====
        const int e = t / kPeriod;
        u64 top, bot;
        if (t < 24) TRACE(16 + 3 * t);
        Polled pq{};
        if (GLBX && multi && t > 0) pq = poll(e);
        exchange_rows(0, top, bot);
        if (t < 24) TRACE(17 + 3 * t);
        if (multi && t > 0) take_refresh(e, pq, top, bot);
        if (t < 24) TRACE(18 + 3 * t);
        dump_hist(t);
        if (BWD && t + 1 >= p.iters) break;
        compute_step(std::false_type{}, top, bot, 0, 0u);
        if (t + 1 >= p.iters) break;
        // ===== odd step t+1: its results feed the refresh at step t+2 =====
        if (t < 23) TRACE(16 + 3 * (t + 1));
        exchange_rows(1, top, bot);
        if (t < 23) { TRACE(17 + 3 * (t + 1)); TRACE(18 + 3 * (t + 1)); }
        dump_hist(t + 1);
        if (BWD && t + 2 >= p.iters) break;
        if (multi && t + 2 < p.iters) {
            const int rnext = (e + 1) & 1;
            if (!GLB && threadIdx.x == 0) mbar_arrive_expect_tx(sm_base + kHaloBar + 8u * rnext, expect_bytes);
            TRACE(89);
            compute_step(std::true_type{}, top, bot, rnext, p.tag_base + (uint32_t)(e + 1));
        } else {
            compute_step(std::false_type{}, top, bot, 0, 0u);
        }
    }
    TRACE(14);
#ifdef CSPN_TRACE
    if (g_trace && lane == 0) g_trace[((size_t)(blockIdx.z * gridDim.y * gridDim.x + blockIdx.y * gridDim.x + blockIdx.x) * (blockDim.x >> 5) + warp) * kTraceSlots + 95] = repolls;
#endif

    // region of this cluster tile whose results are exact (not inside the decaying margin of a tile edge that is
    // not an image border); every image pixel is inside exactly one such region
    const int vx0 = tix > 0 ? tix * p.stepx + p.margin : 0;
    const int vx1 = tix == p.ntx - 1 ? W : tix * p.stepx + p.ew - p.margin;
    const int vy0 = tiy > 0 ? tiy * p.stepy + p.margin : 0;
    const int vy1 = tiy == p.nty - 1 ? H : tiy * p.stepy + p.eh - p.margin;

    if constexpr (!BWD) {
        if (GLBX && multi) {
            // every message addressed to this CTA has been consumed: leave the inbox clean for the next launch / graph replay
            __syncthreads();
            uint4* box = p.inbox + (size_t)my_blk * IG::size;
            for (uint32_t i = threadIdx.x; i < IG::size; i += NW * 32) box[i] = make_uint4(0, 0, 0, 0);
        }
        // ---- epilogue: only the final depth goes back to HBM, and only from the pixels this CTA is authoritative for
        T* ob = p.out + (size_t)plane * hw;
        if (poisoned) {
#pragma unroll
            for (int i = 0; i < P; ++i) A[i] = pk(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000));
        }
        if (lane_auth) {
            const bool ok0 = gx >= vx0 && gx < vx1 && gx < W, ok1 = gx + 1 >= vx0 && gx + 1 < vx1 && gx + 1 < W;
#pragma unroll
            for (int i = 0; i < P; ++i) {
                const int ty = warp * P + i, gy = gy0 + i;
                if (ty < ry0 || ty > ry1 || gy < vy0 || gy >= vy1 || gy >= H) continue;
                const size_t off = (size_t)gy * W + gx;
                if (ok0 && ok1 && vec_ok) {
                    if (sizeof(T) == 4) *reinterpret_cast<float2*>(ob + off) = make_float2(lo_of(A[i]), hi_of(A[i]));
                    else *reinterpret_cast<__half2*>(ob + off) = __floats2half2_rn(lo_of(A[i]), hi_of(A[i]));
                } else {
                    if (ok0) ob[off] = from_f32<T>(lo_of(A[i]));
                    if (ok1) ob[off + 1] = from_f32<T>(hi_of(A[i]));
                }
            }
        }
    } else {
        // ================================ reverse phase ================================
        // With the mask folded into the weights (n' = (1-m) n) the adjoint recurrence is, for t = T-1 .. 0,
        //     gn'_j(p) += G^{t+1}(p) * r^t(p + o_j)            (dL/dn_j = (1-m) * gn'_j)
        //     G^t(q)    = sum_j n'_j(q - o_j) * G^{t+1}(q - o_j) = sum_j wt_j(q) * G^{t+1}(q + o_j),  wt_j(q) = n'_{7-j}(q + o_j)
        //     dL/dd0(p) = m(p) * sum_{t=1..T} G^t(p) + G^0(p)
        // i.e. the same 8-tap gather as the forward with TRANSPOSED weights, which live in shared memory (the
        // registers that held n' now hold the 8 gradient accumulators per pixel).  G needs the same halo
        // treatment as r; the r^t tile (own pixels + the ring it needs, all inside this CTA's tile) comes from the
        // history scratch.
        float* wt = reinterpret_cast<float*>(smem_raw + sizeof(Smem<NW, P>));
        float* rt = wt + BwdTiles<TH>::wt_floats;
        const int T_ = p.iters;
        const int EF = (T_ - 1) / 2;                       // refresh epochs used by the forward phase
        __syncthreads();                                   // history complete and visible CTA-wide; staging buffer idle
        {
            float4* z = reinterpret_cast<float4*>(wt);
            for (uint32_t i = threadIdx.x; i < (uint32_t)(BwdTiles<TH>::wt_floats / 4); i += NW * 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < P; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int jj = j < 4 ? j : j + 1;
                const int dy = jj / 3 - 1, dx = jj % 3 - 1;
                const int ty = warp * P + i + dy, tx = 2 * lane + dx;             // tx in [-1, 63]
                if (ty >= 0 && ty < TH) {
                    float* d = wt + ((size_t)(7 - j) * TH + ty) * kTileW;
                    if (tx >= 0) d[tx] = lo_of(nw[i][j]);
                    if (tx + 1 < kTileW) d[tx + 1] = hi_of(nw[i][j]);
                }
            }
        }
        auto fetch_rt = [&](int t, int buf) {
            const float* src = hist_cta + (size_t)t * (TH * kTileW);
            const uint32_t dst = smem_u32(rt + (size_t)buf * BwdTiles<TH>::rt_floats);
            for (uint32_t c = threadIdx.x; c < (uint32_t)(TH * kTileW / 4); c += NW * 32) cp_async16(dst + 16u * c, src + 4 * c);
        };
        fetch_rt(T_ - 1, 0);

        u64 gn[P][8], sumG[P];
#pragma unroll
        for (int i = 0; i < P; ++i) {
            sumG[i] = 0ull;
#pragma unroll
            for (int j = 0; j < 8; ++j) gn[i][j] = 0ull;
        }
        {   // G^T = dL/d out on the whole tile (zero outside the image)
            const T* gob = p.gout + dplane * hw;
            if (vec_ok) {
                typedef typename std::conditional<sizeof(T) == 4, float2, __half2>::type V2;
                const int cgx = min(max(gx, 0), W - 2);
                V2 gv[P];
#pragma unroll
                for (int i = 0; i < P; ++i) gv[i] = *reinterpret_cast<const V2*>(gob + (size_t)min(max(gy0 + i, 0), H - 1) * W + cgx);
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    const int gy = gy0 + i;
                    const float2 v = to_f32x2(gv[i]);
                    A[i] = (gy >= 0 && gy < H && x_in0) ? pk(v.x, v.y) : 0ull;
                }
            } else {
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    const int gy = gy0 + i;
                    const bool row_in = gy >= 0 && gy < H;
                    const size_t off = (size_t)gy * W + gx;
                    A[i] = pk((row_in && x_in0) ? to_f32(gob[off]) : 0.f, (row_in && x_in1) ? to_f32(gob[off + 1]) : 0.f);
                }
            }
        }

        // One reverse step (same source-row organisation as compute_step, in place): r-field row r feeds the
        // accumulators of rows r+1 / r / r-1 with the OLD G of those rows, G-field row r feeds the new G.
        auto bwd_step = [&](auto push_tag, u64 top, u64 bot, const float* rtile, int rpar, uint32_t tag) {
            constexpr bool PUSH = decltype(push_tag)::value;
            // (packed FFMA2 here: the scalar form that helps compute_step measured 5 % slower for this step)
            u64 acc[P];
#pragma unroll
            for (int r = -1; r <= P; ++r) {
                const u64 src = r < 0 ? top : (r < P ? A[r < 0 ? 0 : (r < P ? r : 0)] : bot);
                u64 s1, s2;
                shifted(src, s1, s2);
                const int ty = warp * P + r;
                u64 rs = 0ull;                                            // r^t row (zero outside the CTA tile)
                if (ty >= 0 && ty < TH) rs = *reinterpret_cast<const u64*>(rtile + ty * kTileW + 2 * lane);
                u64 q1, q2;
                shifted(rs, q1, q2);
                if (r + 1 < P) {
                    const int i = r + 1;
                    const float* w = wt + (size_t)(warp * P + i) * kTileW + 2 * lane;
                    u64 a = mul2(*reinterpret_cast<const u64*>(w + 0 * TH * kTileW), s1);
                    a = fma2(*reinterpret_cast<const u64*>(w + 1 * TH * kTileW), src, a);
                    acc[i] = fma2(*reinterpret_cast<const u64*>(w + 2 * TH * kTileW), s2, a);
                    gn[i][0] = fma2(A[i], q1, gn[i][0]);
                    gn[i][1] = fma2(A[i], rs, gn[i][1]);
                    gn[i][2] = fma2(A[i], q2, gn[i][2]);
                }
                if (r >= 0 && r < P) {
                    const int i = r < 0 ? 0 : (r < P ? r : 0);
                    const float* w = wt + (size_t)(warp * P + i) * kTileW + 2 * lane;
                    u64 a = acc[i];
                    a = fma2(*reinterpret_cast<const u64*>(w + 3 * TH * kTileW), s1, a);
                    acc[i] = fma2(*reinterpret_cast<const u64*>(w + 4 * TH * kTileW), s2, a);
                    gn[i][3] = fma2(A[i], q1, gn[i][3]);
                    gn[i][4] = fma2(A[i], q2, gn[i][4]);
                }
                if (r >= 1) {
                    const int i = r - 1;
                    const float* w = wt + (size_t)(warp * P + i) * kTileW + 2 * lane;
                    u64 a = acc[i];
                    a = fma2(*reinterpret_cast<const u64*>(w + 5 * TH * kTileW), s1, a);
                    a = fma2(*reinterpret_cast<const u64*>(w + 6 * TH * kTileW), src, a);
                    a = fma2(*reinterpret_cast<const u64*>(w + 7 * TH * kTileW), s2, a);
                    gn[i][5] = fma2(A[i], q1, gn[i][5]);
                    gn[i][6] = fma2(A[i], rs, gn[i][6]);
                    gn[i][7] = fma2(A[i], q2, gn[i][7]);
                    sumG[i] = add2(sumG[i], A[i]);
                    A[i] = a;                                             // old G of row i is dead from here on
                    if (PUSH) {
                        const int tyo = warp * P + i;
                        if (push_l) sm.colstage[rpar][0][tyo] = a;
                        if (push_r) sm.colstage[rpar][1][tyo] = a;
                    }
                }
            }
            if (PUSH) ship(rpar, tag);
        };

        for (int s = 0; s < T_; s += 2) {
            // ===== even reverse step s (t = T-1-s) =====
            const int e = EF + s / kPeriod;
            u64 top, bot;
            Polled pq{};
            if (GLB && multi && s > 0) pq = poll(e);
            cp_async_wait_all();                              // my share of r^t for this step has landed ...
            exchange_rows(0, top, bot);                       // ... and after this barrier everybody's has
            if (s + 1 < T_) fetch_rt(T_ - 2 - s, 1);          // next step's tile (its buffer was last read in step s-1)
            if (multi && s > 0) take_refresh(e, pq, top, bot);
            bwd_step(std::false_type{}, top, bot, rt, 0, 0u);
            if (s + 1 >= T_) break;
            // ===== odd reverse step s+1 =====
            cp_async_wait_all();
            exchange_rows(1, top, bot);
            if (s + 2 < T_) fetch_rt(T_ - 3 - s, 0);
            if (multi && s + 2 < T_) {
                const int rnext = (e + 1) & 1;
                if (!GLB && threadIdx.x == 0) mbar_arrive_expect_tx(sm_base + kHaloBar + 8u * rnext, expect_bytes);
                bwd_step(std::true_type{}, top, bot, rt + BwdTiles<TH>::rt_floats, rnext, p.tag_base + (uint32_t)(e + 1));
            } else {
                bwd_step(std::false_type{}, top, bot, rt + BwdTiles<TH>::rt_floats, 0, 0u);
            }
        }

        if (GLB && multi) {
            __syncthreads();
            uint4* box = p.inbox + (size_t)my_blk * IG::size;
            for (uint32_t i = threadIdx.x; i < IG::size; i += NW * 32) box[i] = make_uint4(0, 0, 0, 0);
        }

        // ---- epilogue: dL/d depth and dL/d guidance for the pixels this CTA is authoritative for; A holds G^0.
        // Mode NEW needs nothing from HBM but the sparse mask: n'_j(p) is read back from the transposed weight tile,
        // 1/S and the sign of the raw guidance come from the stash the prologue left in shared memory.
        // Mode OURS reads its 8 guidance values again to redo the softmax (the forward zeroed border taps).
        T* ggb = p.gg + (size_t)b * p.Cg * hw;
        T* gdb = p.gd + dplane * hw;
        const float qnan = __int_as_float(0x7fc00000);
        const u64 qnan2 = pk(qnan, qnan), one2 = pk(1.f, 1.f), mone2 = pk(-1.f, -1.f);
        if (lane_auth) {
            const bool ok0 = gx >= vx0 && gx < vx1 && gx < W, ok1 = gx + 1 >= vx0 && gx + 1 < vx1 && gx + 1 < W;
#pragma unroll
            for (int i = 0; i < P; ++i) {
                const int ty = warp * P + i, gy = gy0 + i;
                if (ty < ry0 || ty > ry1 || gy < vy0 || gy >= vy1 || gy >= H || (!ok0 && !ok1)) continue;
                const size_t off = (size_t)gy * W + gx;
                float m0 = 0.f, m1 = 0.f;
                if (sb) {
                    if (ok0) m0 = signf(to_f32(sb[off]));
                    if (ok1) m1 = signf(to_f32(sb[off + 1]));
                }
                const u64 m2 = pk(m0, m1);
                const u64 om = fma2(m2, mone2, one2);                                  // 1 - m
                u64 gdv = fma2(m2, sumG[i], A[i]);                                     // m * sum_t G^t + G^0
                if (poisoned) gdv = qnan2;
                if (ok0) gdb[off] = from_f32<T>(lo_of(gdv));
                if (ok1) gdb[off + 1] = from_f32<T>(hi_of(gdv));
                if (MODE == CSPN_MODE_NEW) {
                    // dL/dW_j(p) = ((1-m) gn'_j - sum_i n'_i gn'_i) / S; it belongs to guidance element (k = 7-j, p + o_j) times
                    // the sign of that element.  Elements (k, q) whose p = q - o_k is outside the image get 0: that is
                    // channel j at THIS pixel whenever p + o_j is outside.
                    const u64 inv = *reinterpret_cast<const u64*>(sinv + ty * kTileW + 2 * lane);
                    const uint32_t bits = sgn[ty * 32 + lane];
                    u64 D = 0ull;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int jj = j < 4 ? j : j + 1;
                        const int tyn = ty + jj / 3 - 1, tx = 2 * lane + jj % 3 - 1;
                        float n0 = 0.f, n1 = 0.f;
                        if (tyn >= 0 && tyn < TH) {
                            const float* w = wt + ((size_t)(7 - j) * TH + tyn) * kTileW;       // wt_{7-j}(p + o_j) = n'_j(p)
                            if (tx >= 0) n0 = w[tx];
                            if (tx + 1 < kTileW) n1 = w[tx + 1];
                        }
                        D = fma2(pk(n0, n1), gn[i][j], D);
                    }
                    const u64 negD = D ^ 0x8000000080000000ull;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int jj = j < 4 ? j : j + 1;
                        const int yy = gy + jj / 3 - 1, xx = gx + jj % 3 - 1;
                        u64 gw = mul2(fma2(om, gn[i][j], negD), inv);
                        if (poisoned) gw = qnan2;
                        float v0 = lo_of(gw), v1 = hi_of(gw);
                        v0 = (bits >> j) & 1u ? -v0 : v0;
                        v0 = (bits >> (16 + j)) & 1u ? 0.f : v0;
                        v1 = (bits >> (8 + j)) & 1u ? -v1 : v1;
                        v1 = (bits >> (24 + j)) & 1u ? 0.f : v1;
                        const bool rin = yy >= 0 && yy < H;
                        T* dst = ggb + (size_t)(7 - j) * hw + (size_t)yy * W + xx;
                        T* zer = ggb + (size_t)j * hw + off;
                        if (ok0) { if (rin && xx >= 0 && xx < W) dst[0] = from_f32<T>(v0); else zer[0] = from_f32<T>(0.f); }
                        if (ok1) { if (rin && xx + 1 >= 0 && xx + 1 < W) dst[1] = from_f32<T>(v1); else zer[1] = from_f32<T>(0.f); }
                    }
                } else {
                    const T* gb = p.g + (size_t)b * p.gbs + off;
                    float z0[8], z1[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        z0[j] = ok0 ? to_f32(gb[(size_t)j * hw]) : 0.f;
                        z1[j] = ok1 ? to_f32(gb[(size_t)j * hw + 1]) : 0.f;
                    }
                    float mx0 = z0[0], mx1 = z1[0];
#pragma unroll
                    for (int j = 1; j < 8; ++j) { mx0 = fmaxf(mx0, z0[j]); mx1 = fmaxf(mx1, z1[j]); }
                    float s0 = 0.f, s1 = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) { z0[j] = expf(z0[j] - mx0); z1[j] = expf(z1[j] - mx1); s0 += z0[j]; s1 += z1[j]; }
                    const u64 rs2 = pk(fast_rcp(s0), fast_rcp(s1));
                    u64 sj[8], gq[8], dot = 0ull;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int jj = j < 4 ? j : j + 1;
                        const int yy = gy + jj / 3 - 1, xx = gx + jj % 3 - 1;
                        const bool rin = yy >= 0 && yy < H;
                        sj[j] = mul2(pk(z0[j], z1[j]), rs2);
                        const u64 q = mul2(om, gn[i][j]);
                        // taps on the zero padding carry no gradient
                        gq[j] = pk((rin && xx >= 0 && xx < W) ? lo_of(q) : 0.f, (rin && xx + 1 >= 0 && xx + 1 < W) ? hi_of(q) : 0.f);
                        dot = fma2(sj[j], gq[j], dot);
                    }
                    const u64 negdot = dot ^ 0x8000000080000000ull;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        u64 o = mul2(sj[j], add2(gq[j], negdot));
                        if (poisoned) o = qnan2;
                        T* dst = ggb + (size_t)j * hw + off;
                        if (ok0) dst[0] = from_f32<T>(lo_of(o));
                        if (ok1) dst[1] = from_f32<T>(hi_of(o));
                    }
                }
            }
        }
    }
    if (GLBX && poisoned && p.status && lane == 0) *p.status = 1;
    TRACE(15);
}

// The kernel.  Hardware clusters (GLB = false): one CTA tile per CTA, grid = (cluster tiles x cluster shape, planes).
// Global-memory exchange (GLB = true): a persistent 1-D grid of at most one CTA per SM slot under a cooperative launch
// (all co-resident); CTA c works through tiles c, c + grid, c + 2 grid, ... of the linear order (plane, tile row, tile
// column).  Every image is ONE virtual cluster of cx x cy tiles - no hardware limit of 16, no decaying margins - and
// tiles only ever wait for tiles of their own image at most cx + 1 positions away: with more CTAs than that, the
// CTA owning a larger-numbered neighbour is at worst busy with a tile that precedes every tile waiting for it, so
// the wavefront always advances (the spin in take_refresh is bounded regardless).
template <typename T, int P, int NW, int MODE, bool TMA, bool GLB, bool BWD, bool HYB = false>
__global__ void __launch_bounds__(NW * 32, (NW <= 5 ? 2 : 1))
fused3x3_kernel(const __grid_constant__ FusedParams<T> p, const __grid_constant__ CUtensorMap gmap)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem<NW, P>& sm = *reinterpret_cast<Smem<NW, P>*>(smem_raw);
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&sm.halo_bar[0]), 1);
        mbar_init(smem_u32(&sm.halo_bar[1]), 1);
#pragma unroll
        for (int k = 0; k < 8; ++k) mbar_init(smem_u32(&sm.tma_bar[k]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if ((GLB || HYB) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0 && p.status) *p.status = 0;     // a timeout (2^21 polls later at the earliest) sets it to 1
    if (!GLB) {
        // "my barriers exist": neighbours may only push into this CTA after everyone passed the matching wait
        if (p.cx * p.cy > 1) cluster_arrive();
        fused_tile<T, P, NW, MODE, TMA, GLB, BWD, HYB>(p, &gmap, smem_raw, blockIdx.x, blockIdx.y, blockIdx.z, gridDim.x, gridDim.y, 0u);
    } else {
        const uint32_t per_image = (uint32_t)(p.cx * p.cy);
        uint32_t use = 0u;
        for (uint32_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++use) {
            if (use) __syncthreads();                     // the previous tile's shared memory is dead in every warp
            const uint32_t in_image = tile % per_image;
            fused_tile<T, P, NW, MODE, TMA, GLB, BWD>(p, &gmap, smem_raw, in_image % (uint32_t)p.cx, in_image / (uint32_t)p.cx, tile / per_image,
                                                      (uint32_t)p.cx, (uint32_t)p.cy, use);
        }
    }
}

// ---- host side: pick the cluster shape and the tiling ----------------------------------------------------
#ifndef CSPN_FWD_WARPS
#define CSPN_FWD_WARPS 8
#define CSPN_FWD_ROWS 10
#endif
constexpr int kNW = CSPN_FWD_WARPS;           // forward: warps per CTA (one CTA per SM) ...
constexpr int kPFwd = CSPN_FWD_ROWS;          // ... x rows per warp: 64 x 80 pixel register tile per CTA
#ifndef CSPN_BWD_WARPS
#define CSPN_BWD_WARPS 8
#endif
constexpr int kNWBwd = CSPN_BWD_WARPS;        // backward: 64 x 64 tile (twice the per-pixel register state of the forward);
constexpr int kPBwd = 64 / kNWBwd;            //           8 warps x 8 rows (16 warps x 4 rows at 128 registers measured the same)
constexpr int kHistSlots = 192;               // history scratch tiles, indexed by %smid (B200: 148 SMs; guarded in the kernel)


inline int tiles_needed(int extent, int size, int margin, int* step)
{
    if (extent >= size) { *step = extent; return 1; }     // one tile reaches both image borders
    const int s = extent - 2 * margin;
    *step = s;
    if (s <= 0) return -1;
    // tile i covers [i*s, i*s + extent); the last one must reach the image border
    return (size - extent + s - 1) / s + 1;
}

// What the GPU can hold at once (1 CTA per SM kernels): SM count and, per hardware cluster size, how many clusters
// are co-resident.  Defaults = B200 as measured with tools/microbench/clusters.cu (GPCs of different sizes make
// large clusters expensive: 7 clusters of 11-16 CTAs, 15 of 7-9, 26 of 5); refined per device at run time.
inline Capacity default_capacity(int ctas_per_sm = 1)
{
    Capacity c{148, {0, 148, 74, 45, 33, 26, 22, 15, 15, 15, 11, 7, 7, 7, 7, 7, 7}};
    c.sms *= ctas_per_sm;
    for (int i = 1; i <= 16; ++i) c.clusters[i] *= ctas_per_sm;
    return c;
}

inline int exchange_override()
{
    // debugging knob: dsmem = hardware clusters only, global = stream wherever it is possible, hybrid = row clusters wherever possible
    static const int force = [] {
        const char* v = getenv("CSPN_EXCHANGE");
        return !v ? 0 : (!strcmp(v, "dsmem") ? 1 : (!strcmp(v, "global") ? 2 : (!strcmp(v, "hybrid") ? 3 : 0)));
    }();
    return force;
}

// th = rows of one CTA tile (NW * P); planes = independent images (B * C).  Two ways to run a problem:
//  * hardware clusters (DSMEM halo exchange, ~0.75 of the tile time of the other way): cluster tiles of cx x cy <= 16
//    CTAs with decaying margins between them; time ~ ceil(clusters / co-resident clusters of that size);
//  * stream (tl.stream): every image is one virtual cluster of cx x cy tiles without margins, halo exchange through
//    global memory, a persistent grid of one CTA per SM slot walks the tiles as a wavefront; time ~ tiles / slots.
//  * hybrid (tl.hyb, forward only): the stream tiling, but every tile ROW of an image is a hardware cluster of (cx, 1) CTAs:
//    left / right rims through DSMEM, the rows above / below through the global inboxes, shipped from inside the sweep.  One CTA
//    per tile, every cluster resident at once (cy * planes <= co-resident clusters of cx CTAs).
inline Tiling choose_tiling(int H, int W, int iters, int th, long planes, const Capacity& cap, int hyb_policy = 0)
{
    const int step_y = th - 2 * kHaloY;
    Tiling best{}; best.ok = false; best.ctas = 0; best.stream = false;
    double best_cost = 0.0;
    auto consider = [&](const Tiling& t, double cost) {
        // cheapest wins; ties go to fewer CTAs, then to the smaller cluster (cheaper to place)
        if (!best.ok || cost < best_cost - 1e-9 ||
            (cost < best_cost + 1e-9 && (t.ctas < best.ctas || (t.ctas == best.ctas && t.cx * t.cy < best.cx * best.cy)))) {
            best = t; best_cost = cost;
        }
    };
    {
        for (int cx = 1; cx <= 16; ++cx)
            for (int cy = 1; cx * cy <= 16; ++cy) {
                Tiling t{}; t.cx = cx; t.cy = cy; t.stream = false;
                t.ew = kStepX * (cx - 1) + kTileW; t.eh = step_y * (cy - 1) + th;
                t.ntx = tiles_needed(t.ew, W, iters, &t.stepx);
                t.nty = tiles_needed(t.eh, H, iters, &t.stepy);
                if (t.ntx < 0 || t.nty < 0) continue;
                if ((t.stepx & 1) != 0) continue;             // keep pixel pairs at even x
                t.ctas = (long)t.ntx * t.nty * cx * cy; t.ok = true;
                const long held = cap.clusters[cx * cy];
                if (held <= 0) continue;
                const long clusters = (long)t.ntx * t.nty * planes;
                consider(t, 0.75 * (double)((clusters + held - 1) / held));
            }
    }
    if (exchange_override() != 1 && iters <= 60) {
        Tiling t{}; t.stream = true; t.ntx = t.nty = 1;
        t.cx = W <= kTileW ? 1 : (W - kTileW + kStepX - 1) / kStepX + 1;
        t.cy = H <= th ? 1 : (H - th + step_y - 1) / step_y + 1;
        t.ew = kStepX * (t.cx - 1) + kTileW; t.eh = step_y * (t.cy - 1) + th;
        t.stepx = t.ew; t.stepy = t.eh;
        t.ctas = (long)t.cx * t.cy; t.ok = true;
        const long total = t.ctas * planes;
        const long grid = total < cap.sms ? total : cap.sms;
        // All tiles of an image advance in lockstep (every refresh waits for all four neighbours), so a tile e refreshes
        // into the loop needs the tile e hops away to have STARTED: an image with more tiles than the persistent grid has
        // CTAs would wait on tiles that only start when earlier ones finish - a circular wait.  Stream mode is therefore
        // only offered when a whole image is resident at once; larger images run as hardware clusters with margins (or,
        // forward, through the dual-slot kernel, which cuts them into resident units).
        if (t.cx * t.cy > 1 && t.ctas <= cap.sms && total <= kMaxGlobalExchangeCtas) {
            // the grid effectively works on floor(slots / tiles per image) images at a time (measured: 32 KITTI images of
            // 105 tiles take 32 tile times on 148 SMs)
            const long groups = cap.sms / t.ctas;
            double cost = (double)((planes + groups - 1) / groups);
            if (exchange_override() == 2) cost = 0.0;
            consider(t, cost);
        }
        // hybrid: one round by construction.  hyb_policy 0: never (backward), 1: only when forced (CSPN_EXCHANGE=hybrid), 2: by cost.
        // Measured 27.6-28.1 vs 28.3 us on the headline (mode NEW; inside the build-to-build spread, fp16 25.9 vs
        // 25.7 us: stays with the stream transport) and 30.0 vs 32.3 us in mode OURS (takes the hybrid transport where it fits).
        const bool hyb_forced = exchange_override() == 3;
        if (hyb_policy > 0 && (hyb_forced || (hyb_policy == 2 && exchange_override() == 0)) && t.cx >= 2 && t.cx <= 16 && t.cy >= 2 &&
            planes <= 65535 && t.cy * planes <= (long)cap.clusters[t.cx] && total <= kMaxGlobalExchangeCtas) {
            Tiling h = t; h.stream = false; h.hyb = true;
            consider(h, hyb_forced ? 0.0 : 0.93);
        }
    }
    return best;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
        return (EncodeTiledFn)f;
    }();
    return fn;
}

// 4-D view (W, H, 8 channels, B) of the guidance tensor; box = one channel plane of the staging buffer.
template <typename T, int TH, int MODE>
bool make_guidance_map(const T* guidance, int64_t gbs, int B, int H, int W, CUtensorMap* map)
{
    using St = Stage<T, TH, MODE>;
    const size_t es = sizeof(T);
    if (!encode_tiled()) return false;
    if (((uintptr_t)guidance & 15) || ((size_t)W * es) % 16 || ((size_t)gbs * es) % 16) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, 8, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)W * es, (cuuint64_t)H * W * es, (cuuint64_t)gbs * es};
    const cuuint32_t box[4] = {(cuuint32_t)St::cols, (cuuint32_t)St::rows, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = encode_tiled()(map, es == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                                      const_cast<T*>(guidance), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <int TH> constexpr size_t inbox_bytes() { return (size_t)InboxGeom<TH>::size * sizeof(uint4); }
constexpr size_t kStatusBytes = 256;     // stream-mode scratch starts with the status word (rest of the 256 bytes unused)

template <typename T, int P, int NW, int MODE, bool TMA, bool BWD>
constexpr size_t fused_smem_bytes()
{
    constexpr size_t stage = TMA ? Stage<T, NW * P, MODE>::bytes : 0;
    constexpr size_t tiles = BWD ? BwdTiles<NW * P>::bytes : 0;
    constexpr size_t ctile = (size_t)NW * P * 32 * sizeof(u64);      // re-injection tile (aliases the staging buffer)
    constexpr size_t dyn = stage > tiles ? (stage > ctile ? stage : ctile) : (tiles > ctile ? tiles : ctile);
    return sizeof(Smem<NW, P>) + dyn + (BWD ? BwdTiles<NW * P>::stash_bytes : 0);
}

template <typename T, int P, int NW, int MODE, bool TMA, bool GLB, bool BWD, bool HYB = false>
int launch_variant(const FusedParams<T>& p, const CUtensorMap& map, const Tiling& tl, int planes, long grid_ctas, cudaStream_t stream)
{
    auto kern = fused3x3_kernel<T, P, NW, MODE, TMA, GLB, BWD, HYB>;
    constexpr size_t smem = fused_smem_bytes<T, P, NW, MODE, TMA, BWD>();
    static_assert(smem * (NW <= 5 ? 2 : 1) <= 227 * 1024, "shared memory budget of one SM exceeded");
    // the two opt-ins are per device and last for the life of the context: set them once per device, not on every launch
    static std::atomic<uint64_t> configured{0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    const uint64_t bit = 1ull << (dev & 63);
    if (!(configured.load(std::memory_order_acquire) & bit)) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured.fetch_or(bit, std::memory_order_release);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = GLB ? dim3((unsigned)grid_ctas) : dim3((unsigned)(tl.ntx * tl.cx), (unsigned)(tl.nty * tl.cy), (unsigned)planes);   // HYB: ntx = nty = 1
    cfg.blockDim = dim3(NW * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    cfg.attrs = at; cfg.numAttrs = 1;
    if (GLB) {
        // neighbours talk through global memory and spin on it: every CTA of the grid must be resident
        at[0].id = cudaLaunchAttributeCooperative;
        at[0].val.cooperative = 1;
    } else {
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)tl.cx; at[0].val.clusterDim.y = HYB ? 1u : (unsigned)tl.cy; at[0].val.clusterDim.z = 1;
        if (HYB) {
            // the clusters of an image spin on each other through global memory: the launch is cooperative as well.  Measured
            // (tools/microbench/coopcluster.cu, driver 580 / CUDA 12.9): cluster + cooperative launches are accepted and are
            // refused with cudaErrorCooperativeLaunchTooLarge one cluster above cudaOccupancyMaxActiveClusters - the same
            // residency guarantee the stream transport has
            at[1].id = cudaLaunchAttributeCooperative;
            at[1].val.cooperative = 1;
            cfg.numAttrs = 2;
        }
    }
    e = cudaLaunchKernelEx(&cfg, kern, p, map);
    if (e != cudaSuccess) return (int)e;
    ++call_stats().launches;
    return 0;
}

// What the current device holds at once for one kernel configuration: CTA slots and co-resident clusters per cluster
// size (occupancy queries, cached per device).  The B200 defaults when there is no usable device (workspace queries
// on a machine without a GPU).
template <int P, int NW, bool BWD>
Capacity capacity()
{
    constexpr int per_sm = NW <= 5 ? 2 : 1;
    static Capacity cache[64];
    static bool valid[64] = {};
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return default_capacity(per_sm); }
    std::lock_guard<std::mutex> lock(mu);
    if (valid[dev & 63]) return cache[dev & 63];
    auto kern = fused3x3_kernel<float, P, NW, CSPN_MODE_NEW, true, false, BWD>;
    constexpr size_t smem = fused_smem_bytes<float, P, NW, CSPN_MODE_NEW, true, BWD>();
    int sms = 0, blocks = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0 ||
        cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kern, NW * 32, smem) != cudaSuccess || blocks <= 0) {
        cudaGetLastError();
        return default_capacity(per_sm);
    }
    Capacity c{}; c.sms = sms * blocks; c.clusters[0] = 0;
    for (int size = 1; size <= 16; ++size) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)size, 1, 64); cfg.blockDim = dim3(NW * 32); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)size; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); n = default_capacity(per_sm).clusters[size]; }
        c.clusters[size] = n;
    }
    cache[dev & 63] = c; valid[dev & 63] = true;
    return c;
}

template <typename T>
FusedParams<T> forward_params(const FwdArgs<T>& a)
{
    FusedParams<T> p{};
    p.g = a.guidance; p.gbs = a.gbs; p.depth = a.depth; p.sparse = a.sparse; p.sparse_channels = a.sparse_channels; p.out = a.out;
    p.C = a.C; p.H = a.H; p.W = a.W; p.iters = a.iters;
    return p;
}

inline std::atomic<uint32_t>& exchange_epoch()
{
    static std::atomic<uint32_t> e{0x5a17u};
    return e;
}

// Common launcher: p carries the problem (pointers, sizes, tiling); inbox / inbox_bytes = optional scratch for the
// global-memory exchange.
template <typename T, int P, int NW, int MODE, bool BWD>
int launch(FusedParams<T> p, const Tiling& tl, int B, void* inbox, size_t inbox_avail, cudaStream_t stream)
{
    constexpr int TH = NW * P;
    p.cx = tl.cx; p.cy = tl.cy; p.ntx = tl.ntx; p.nty = tl.nty; p.stepx = tl.stepx; p.stepy = tl.stepy; p.ew = tl.ew; p.eh = tl.eh;
    p.margin = p.iters;
    const int planes = B * p.C;
    // Exchange transport follows the tiling: hardware clusters talk through DSMEM, a streamed problem through inboxes
    // in global memory (one per tile), walked by a persistent grid of at most one CTA per SM slot.
    const bool glb = tl.stream;
    long grid_ctas = 0;
    if (glb || tl.hyb) {
        const long total = tl.ctas * planes;
        if (!inbox || inbox_avail < kStatusBytes + (size_t)total * inbox_bytes<TH>()) return CSPN_ERR_WORKSPACE;
        p.status = (int*)inbox;                                       // first word of the scratch: exchange-timeout flag
        inbox = (char*)inbox + kStatusBytes;
        const Capacity cap = capacity<P, NW, BWD>();
        grid_ctas = total < cap.sms ? total : cap.sms;
        p.total_tiles = (uint32_t)total;
        p.inbox = (uint4*)inbox;
        p.tag_base = exchange_epoch().fetch_add(1, std::memory_order_relaxed) << 7;      // + refresh index (1..120) is never 0
    }
    alignas(64) CUtensorMap map;
    memset(&map, 0, sizeof map);
    const bool tma = make_guidance_map<T, TH, MODE>(p.g, p.gbs, B, p.H, p.W, &map);      // false: unaligned guidance, plain-load prologue
    if constexpr (!BWD) {
        if (tl.hyb) return tma ? launch_variant<T, P, NW, MODE, true, false, false, true>(p, map, tl, planes, 0, stream) : launch_variant<T, P, NW, MODE, false, false, false, true>(p, map, tl, planes, 0, stream);
    }
    if (glb) return tma ? launch_variant<T, P, NW, MODE, true, true, BWD>(p, map, tl, planes, grid_ctas, stream) : launch_variant<T, P, NW, MODE, false, true, BWD>(p, map, tl, planes, grid_ctas, stream);
    return tma ? launch_variant<T, P, NW, MODE, true, false, BWD>(p, map, tl, planes, 0, stream) : launch_variant<T, P, NW, MODE, false, false, BWD>(p, map, tl, planes, 0, stream);
}

}  // namespace
}  // namespace cspn
