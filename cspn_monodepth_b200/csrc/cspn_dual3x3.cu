// Dual-slot streamed forward kernel for the 3x3 CSPN recurrence (both reference modes): the whole T-step loop of
// CSPN_new.py:80-90 / CSPN_ours.py:47-53 in ONE launch, like cspn_fused3x3.cuh, but organised so that the halo
// exchange between CTAs never sits on the critical path.
//
// Why a second organisation (round-2 profile of the 64 x 80 single-tile kernel, profiles/r01_ncu_fused3x3_nyu_b8.txt):
// the sweeps themselves issue at ~85 % of the scheduler rate, but 57 % of the step loop is the halo refresh - the
// L2 round trip of the messages (every tile of an image advances in lockstep, so each refresh exposes one full
// store -> L2 -> load latency), the skew it leaves at the next CTA barrier and the poll/apply code.
//
// Here a CTA owns TWO register tiles ("slots") of 64 x 8P pixels that belong to different, independent units (images or
// margin-separated sub-images) and works on them alternately, one refresh period (two steps) at a time:
//     slot A: refresh, step, step, ship rim | slot B: refresh, step, step, ship rim | slot A: ...
// so the rim of slot A travels through L2 while slot B computes, and vice versa.  The incoming halo of the idle slot is
// prefetched into shared memory with cp.async (LDGSTS, L2 only) half a period ahead; when its tags are current - the
// normal case - the refresh costs a few shared-memory loads and no global round trip.  Messages are 16-byte
// {lo, tag, hi, tag} stores (data and flag in one transaction, no fence) straight from the registers that hold the rim.
//
// Work decomposition: an image plane is cut into ntx x nty units with decaying margins of T pixels (one unit = the
// whole plane whenever it fits); a unit is cx x cy tiles and must fit the GPU (cx * cy <= SMs).  A round processes up
// to 2 * nA units (nA = SMs / tiles per unit): CTA w holds tile (w mod tiles) of unit (w / tiles) of class A and the
// same tile of the corresponding unit of class B.  All tiles of a unit are resident in the same round, so no tile ever
// waits for a tile that has not started (the deadlock of the old persistent stream for images larger than the GPU
// cannot occur); larger batches take several rounds inside the same launch.
#include "cspn_fused3x3.cuh"

namespace cspn {
namespace {

constexpr int kDNW = 8;                      // warps per CTA; the tile is 64 x (8 * P) pixels per slot
constexpr uint32_t kNone = 0xffffffffu;
constexpr int kSpinLimit = 1 << 20;

template <typename T>
struct DualParams {
    const T* g; int64_t gbs;
    const T* depth; const T* sparse; int sparse_channels;
    T* out;
    int C, H, W, iters;
    int cx, cy;              // tiles per unit
    int ntx, nty;            // units per image plane
    int stepx, stepy;        // origin spacing of units
    int ew, eh;              // extent of one unit
    int margin;              // decaying halo at unit edges that are not image borders (= iters)
    int per_unit;            // cx * cy
    int nA;                  // units per class and round
    int total_units;         // planes * ntx * nty
    int rounds;
    uint4* inbox;            // [2 round parities][grid][2 slots] inboxes of InboxGeom<TH>::size uint4
    uint32_t tag_base;       // tag of refresh e in round r is tag_base + 64 r + e (never 0)
    int* status;             // set to 1 when a neighbour never showed up (output is NaN-filled as well)
};

template <int P>
struct __align__(128) DualSm {
    float rowbuf[2][2][kDNW][2][kTileW];     // [slot][step parity][warp][first / last row of the warp's strip][x]
    uint4 colz[2][2][kDNW * P];              // landing zone of the halo columns: [slot][side: left, right][tile row]
    uint4 rowz[2][2][kHaloY][32];            // landing zone of the halo rows:    [slot][side: top, bottom][row][lane]
    u64 tma_bar[2][8];
};

__device__ __forceinline__ u64 pk_bits(uint32_t lo, uint32_t hi)
{
    u64 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
    return d;
}

template <typename T, int P, int MODE>
__global__ void __launch_bounds__(kDNW * 32, 1)
dual3x3_kernel(const __grid_constant__ DualParams<T> p, const __grid_constant__ CUtensorMap gmap)
{
    constexpr int NW = kDNW, TH = NW * P, STEPY = TH - 2 * kHaloY;
    static_assert(P >= 4, "rim rows 2,3 / P-4,P-3 must live in the first / last warp");
    using St = Stage<T, TH, MODE>;
    using IB = InboxGeom<TH>;
    typedef typename std::conditional<sizeof(T) == 4, float2, __half2>::type V2;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    DualSm<P>& sm = *reinterpret_cast<DualSm<P>*>(smem_raw);
    constexpr size_t kStageOff = (sizeof(DualSm<P>) + 127) & ~(size_t)127;
    T* const stage_base = reinterpret_cast<T*>(smem_raw + kStageOff);                 // slot s: + s * 8 * St::plane
    u64* const ctile_base = reinterpret_cast<u64*>(smem_raw + kStageOff);            // aliases the staging buffer of its slot
    constexpr size_t kSlotStageBytes = St::bytes;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int k = 0; k < 8; ++k) mbar_init(smem_u32(&sm.tma_bar[s][k]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ---- position of this CTA's tiles inside their units (the same for both slots) ------------------------------
    const int w = blockIdx.x;
    const int uidx = w / p.per_unit, within = w - uidx * p.per_unit;
    const int ccy = within / p.cx, ccx = within - ccy * p.cx;
    const bool hasL = ccx > 0, hasR = ccx < p.cx - 1, hasU = ccy > 0, hasD = ccy < p.cy - 1;
    const bool multi = p.per_unit > 1;
    const int H = p.H, W = p.W;
    const size_t hw = (size_t)H * W;

    // Outgoing messages of this lane, as uint4 offsets from (inbox set base + slot * IB::size), refresh parity 0:
    //   col_dst  lanes 1 / 30: my rim column, all P rows of the warp -> left neighbour's right box / right neighbour's left box
    //   row_dst  first / last warp: my rim rows (2,3 / TH-4,TH-3) -> upper neighbour's bottom box / lower neighbour's top box,
    //            except the lanes whose pixels I do not own (lane 0 / 31 next to a left / right neighbour)
    //   diag_dst first / last warp, lanes 1 / 30: the same rim rows -> the corner lane (31 / 0) of the DIAGONAL neighbour's box
    uint32_t col_dst = kNone, row_dst = kNone, diag_dst = kNone;
    if (multi) {
        const uint32_t slot2 = 2u * IB::size;
        if (lane == 1 && hasL) col_dst = (uint32_t)(w - 1) * slot2 + IB::col_side + (uint32_t)(warp * P);
        if (lane == 30 && hasR) col_dst = (uint32_t)(w + 1) * slot2 + (uint32_t)(warp * P);
        const bool own = !(lane == 0 && hasL) && !(lane == 31 && hasR);
        if (warp == 0 && hasU) {
            const uint32_t up = (uint32_t)(w - p.cx) * slot2 + IB::row_base + IB::row_side;
            if (own) row_dst = up + (uint32_t)lane;
            if (lane == 1 && hasL) diag_dst = up - slot2 + 31u;
            if (lane == 30 && hasR) diag_dst = up + slot2;
        }
        if (warp == NW - 1 && hasD) {
            const uint32_t dn = (uint32_t)(w + p.cx) * slot2 + IB::row_base;
            if (own) row_dst = dn + (uint32_t)lane;
            if (lane == 1 && hasL) diag_dst = dn - slot2 + 31u;
            if (lane == 30 && hasR) diag_dst = dn + slot2;
        }
    }
    // Incoming: what this lane prefetches from the CTA's own inbox (exactly what its warp consumes)
    const int pf_sd = lane >> 4, pf_row = warp * P + (lane & 15);
    const bool pf_col = multi && (lane & 15) < P && (pf_sd == 0 ? hasL : hasR);
    const bool rows_top = multi && warp == 0 && hasU, rows_bot = multi && warp == NW - 1 && hasD;
    const bool edge = multi && ((lane == 0 && hasL) || (lane == 31 && hasR));
    const int edge_sd = lane == 31 ? 1 : 0;

    bool poisoned = false;
    if (blockIdx.x == 0 && threadIdx.x == 0 && p.status) *p.status = 0;     // a timeout (2^20 polls later at the earliest) sets it to 1

    for (int r = 0; r < p.rounds; ++r) {
        // ---- which units this CTA serves in round r ------------------------------------------------------------
        const int first = r * 2 * p.nA;
        const int n = min(p.total_units - first, 2 * p.nA);
        const int nAr = (n + 1) >> 1;
        if (uidx >= nAr) break;                             // later rounds are never larger
        const bool haveB = uidx < n - nAr;
        const int unit_a = first + uidx, unit_b = first + nAr + uidx;
        const uint32_t set_off = (uint32_t)(r & 1) * gridDim.x * 2u * IB::size;
        const uint32_t round_tag = p.tag_base + ((uint32_t)r << 6);
        uint4* const my_box = p.inbox + set_off + (size_t)w * 2u * IB::size;        // slot s: + s * IB::size
        uint4* const out_base = p.inbox + set_off;

        struct Geo { int plane, b, ch, tix, tiy, ox, oy; };
        auto geo_of = [&](int unit) {
            Geo g;
            const int upp = p.ntx * p.nty;
            g.plane = unit / upp;
            const int sub = unit - g.plane * upp;
            g.tiy = sub / p.ntx; g.tix = sub - g.tiy * p.ntx;
            g.b = g.plane / p.C; g.ch = g.plane - g.b * p.C;
            g.ox = g.tix * p.stepx + ccx * kStepX;
            g.oy = g.tiy * p.stepy + ccy * STEPY;
            return g;
        };

        if (r) __syncthreads();                             // every warp is done with the previous round's shared memory

        u64 nwA[P][8], nwB[P][8], A0[P], A1[P];
        {
        const Geo ga = geo_of(unit_a), gb = geo_of(haveB ? unit_b : unit_a);

        // ---- TMA: 8 boxes per slot, one per guidance channel, each on its own barrier ---------------------------
        if (threadIdx.x == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (s == 1 && !haveB) break;
                const Geo& g = s ? gb : ga;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t bar = smem_u32(&sm.tma_bar[s][k]);
                    mbar_arrive_expect_tx(bar, (uint32_t)St::box_bytes);
                    tma_load_4d(smem_u32(stage_base + (size_t)s * 8 * St::plane + (size_t)k * St::plane), &gmap, bar, St::box_x(g.ox), g.oy - St::apron, k, g.b);
                }
            }
        }

        // ---- prologue: loop-invariant weights n'_j = (1-m) n_j in registers, c = m d0 in shared memory, r^0 = d0 ----
        u64 ccA[P], ccB[P];
        V2 dvA[P], svA[P], dvB[P], svB[P];
        auto load_ds = [&](const Geo& g, V2 (&dv)[P], V2 (&sv)[P]) {
            const T* db = p.depth + (size_t)g.plane * hw;
            const T* sb = p.sparse ? p.sparse + ((size_t)g.b * p.sparse_channels + (p.sparse_channels == 1 ? 0 : g.ch)) * hw : nullptr;
            const int cgx = min(max(g.ox + 2 * lane, 0), W - 2);
#pragma unroll
            for (int i = 0; i < P; ++i) {
                const size_t off = (size_t)min(max(g.oy + warp * P + i, 0), H - 1) * W + cgx;
                dv[i] = *reinterpret_cast<const V2*>(db + off);
                if (sb) sv[i] = *reinterpret_cast<const V2*>(sb + off);
            }
        };
        load_ds(ga, dvA, svA);
        if (haveB) load_ds(gb, dvB, svB);

        auto weights = [&](int s, const Geo& g, const V2 (&dv)[P], const V2 (&sv)[P], u64 (&nw)[P][8], u64 (&A)[P], u64 (&cc)[P]) {
            const int gx = g.ox + 2 * lane, gy0 = g.oy + warp * P;
            const bool x_in = gx >= 0 && gx < W;           // W is even and gx is even: both pixels of the pair are in or out together
            const bool has_sparse = p.sparse != nullptr;
#pragma unroll
            for (int i = 0; i < P; ++i) {
                const int gy = gy0 + i;
                const bool in = gy >= 0 && gy < H && x_in;
                const float2 d = to_f32x2(dv[i]);
                float2 m = make_float2(0.f, 0.f);
                if (has_sparse) { const float2 sp2 = to_f32x2(sv[i]); m = make_float2(signf(sp2.x), signf(sp2.y)); }
                A[i] = in ? pk(d.x, d.y) : 0ull;
                cc[i] = in ? pk(m.x, m.y) : 0ull;           // the mask, until it is folded into the weights below
            }
            const T* stage = stage_base + (size_t)s * 8 * St::plane;
            const int x_off = g.ox - St::box_x(g.ox);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                mbar_wait(smem_u32(&sm.tma_bar[s][k]), (uint32_t)(r & 1));
                const T* sp = stage + (size_t)k * St::plane;
                if (MODE == CSPN_MODE_NEW) {
                    // channel k = 7 - j is read AT THE NEIGHBOUR p + o_j (CSPN_new.py:43-67)
                    const int j = 7 - k, jj = j < 4 ? j : j + 1;
                    const int dy = jj / 3 - 1, dx = jj % 3 - 1;
#pragma unroll
                    for (int i = 0; i < P; ++i) {
                        const T* src = sp + (warp * P + i + 1 + dy) * St::cols + x_off + 2 * lane + dx;
                        nw[i][j] = pk(fabsf(to_f32(src[0])), fabsf(to_f32(src[1])));       // zero-filled outside the image
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < P; ++i) {
                        const T* src = sp + (warp * P + i) * St::cols + x_off + 2 * lane;
                        nw[i][k] = pk(to_f32(src[0]), to_f32(src[1]));
                    }
                }
            }
            if (MODE == CSPN_MODE_NEW) {
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    u64 sum = nw[i][7];                     // reference order k = 0..7 (CSPN_new.py:124), k = 7 - j
#pragma unroll
                    for (int k = 1; k < 8; ++k) sum = add2(sum, nw[i][7 - k]);
                    const int gy = gy0 + i;
                    const bool in = gy >= 0 && gy < H && x_in;
                    // S = 0 -> inf -> 0 * inf = NaN like the reference's 0/0; pixels outside the image are virtual zeros
                    const float f0 = in ? (1.f - lo_of(cc[i])) * fast_rcp(lo_of(sum)) : 0.f;
                    const float f1 = in ? (1.f - hi_of(cc[i])) * fast_rcp(hi_of(sum)) : 0.f;
                    const u64 scale = pk(f0, f1);
#pragma unroll
                    for (int j = 0; j < 8; ++j) nw[i][j] = mul2(nw[i][j], scale);
                    cc[i] = mul2(cc[i], A[i]);              // c = m * d0
                }
            } else {
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    const int gy = gy0 + i;
                    const bool in = gy >= 0 && gy < H && x_in;
                    float w0[8], w1[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) { w0[j] = lo_of(nw[i][j]); w1[j] = hi_of(nw[i][j]); }
                    float m0 = w0[0], m1 = w1[0];
#pragma unroll
                    for (int j = 1; j < 8; ++j) { m0 = fmaxf(m0, w0[j]); m1 = fmaxf(m1, w1[j]); }
                    float s0 = 0.f, s1 = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) { w0[j] = expf(w0[j] - m0); w1[j] = expf(w1[j] - m1); s0 += w0[j]; s1 += w1[j]; }
                    const float f0 = in ? (1.f - lo_of(cc[i])) * fast_rcp(s0) : 0.f, f1 = in ? (1.f - hi_of(cc[i])) * fast_rcp(s1) : 0.f;
                    // taps on the zero padding contribute n_j * 0 (no border renormalisation, pac.py:89): drop their weight
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
        };
        weights(0, ga, dvA, svA, nwA, A0, ccA);
        if (haveB) weights(1, gb, dvB, svB, nwB, A1, ccB);

        // The re-injection term lives where the guidance was staged (read once per row and step; every thread only
        // re-reads its own words).
        __syncthreads();                                    // every warp has taken its guidance out of the staging buffers
        {
            u64* const cA = ctile_base + (warp * P) * 32 + lane;
            u64* const cB = reinterpret_cast<u64*>(smem_raw + kStageOff + kSlotStageBytes) + (warp * P) * 32 + lane;
#pragma unroll
            for (int i = 0; i < P; ++i) cA[i * 32] = ccA[i];
            if (haveB) {
#pragma unroll
                for (int i = 0; i < P; ++i) cB[i * 32] = ccB[i];
            }
        }
        }
        u64* const ctA = ctile_base + (warp * P) * 32 + lane;
        u64* const ctB = reinterpret_cast<u64*>(smem_raw + kStageOff + kSlotStageBytes) + (warp * P) * 32 + lane;

        // ---- building blocks of the step loop --------------------------------------------------------------------
        auto exchange_rows = [&](int s, int par, const u64 (&A)[P], u64& top, u64& bot) {
            *reinterpret_cast<u64*>(&sm.rowbuf[s][par][warp][0][2 * lane]) = A[0];
            *reinterpret_cast<u64*>(&sm.rowbuf[s][par][warp][1][2 * lane]) = A[P - 1];
            __syncthreads();
            top = 0ull; bot = 0ull;
            if (warp > 0) top = *reinterpret_cast<const u64*>(&sm.rowbuf[s][par][warp - 1][1][2 * lane]);
            if (warp < NW - 1) bot = *reinterpret_cast<const u64*>(&sm.rowbuf[s][par][warp + 1][0][2 * lane]);
        };

        // One step r'(p) = c(p) + sum_j n'_j(p) r(p + o_j), organised by SOURCE row (scalar FMAs on the halves of the
        // packed pairs; see cspn_fused3x3.cuh compute_step for the derivation), in place.
        auto sweep = [&](const u64 (&nw)[P][8], u64 (&A)[P], const u64* ctile, u64 top, u64 bot) {
            float a0[P], a1[P];
#pragma unroll
            for (int rr = -1; rr <= P; ++rr) {
                const u64 src = rr < 0 ? top : (rr < P ? A[rr < 0 ? 0 : (rr < P ? rr : 0)] : bot);
                const float lo = lo_of(src), hi = hi_of(src);
                const float l = __shfl_up_sync(0xffffffffu, hi, 1), rt = __shfl_down_sync(0xffffffffu, lo, 1);
                if (rr + 1 < P) {
                    const int i = rr + 1;
                    const u64 ci = ctile[i * 32];
                    float x0 = lo_of(ci), x1 = hi_of(ci);
                    x0 = fmaf(lo_of(nw[i][0]), l, x0);   x1 = fmaf(hi_of(nw[i][0]), lo, x1);
                    x0 = fmaf(lo_of(nw[i][1]), lo, x0);  x1 = fmaf(hi_of(nw[i][1]), hi, x1);
                    a0[i] = fmaf(lo_of(nw[i][2]), hi, x0); a1[i] = fmaf(hi_of(nw[i][2]), rt, x1);
                }
                if (rr >= 0 && rr < P) {
                    const int i = rr < 0 ? 0 : (rr < P ? rr : 0);
                    a0[i] = fmaf(lo_of(nw[i][4]), hi, fmaf(lo_of(nw[i][3]), l, a0[i]));
                    a1[i] = fmaf(hi_of(nw[i][4]), rt, fmaf(hi_of(nw[i][3]), lo, a1[i]));
                }
                if (rr >= 1) {
                    const int i = rr - 1;
                    float x0 = a0[i], x1 = a1[i];
                    x0 = fmaf(lo_of(nw[i][5]), l, x0);   x1 = fmaf(hi_of(nw[i][5]), lo, x1);
                    x0 = fmaf(lo_of(nw[i][6]), lo, x0);  x1 = fmaf(hi_of(nw[i][6]), hi, x1);
                    x0 = fmaf(lo_of(nw[i][7]), hi, x0);  x1 = fmaf(hi_of(nw[i][7]), rt, x1);
                    A[i] = pk(x0, x1);
                }
            }
        };

        // Ship the rim of slot s for refresh epoch e1: straight from the registers, data + tag in one 16-byte store.
        auto ship = [&](int s, const u64 (&A)[P], int e1) {
            const uint32_t tag = round_tag + (uint32_t)e1;
            uint4* const base = out_base + (size_t)s * IB::size;
            if (col_dst != kNone) {
                uint4* d = base + col_dst + (e1 & 1) * IB::col_par;
#pragma unroll
                for (int i = 0; i < P; ++i) st_ll(d + i, A[i], tag);
            }
            if (rows_top || rows_bot) {                                          // warp-uniform
                const u64 v0 = warp == 0 ? A[kHaloY] : A[P - 2 * kHaloY], v1 = warp == 0 ? A[kHaloY + 1] : A[P - 2 * kHaloY + 1];
                if (row_dst != kNone) {
                    uint4* d = base + row_dst + (e1 & 1) * IB::row_par;
                    st_ll(d, v0, tag); st_ll(d + 32, v1, tag);
                }
                if (diag_dst != kNone) {
                    uint4* d = base + diag_dst + (e1 & 1) * IB::row_par;
                    st_ll(d, v0, tag); st_ll(d + 32, v1, tag);
                }
            }
        };

        // Start fetching the halo of slot s, refresh epoch e, from this CTA's inbox into the landing zone.
        auto prefetch = [&](int s, int e) {
            const uint4* box = my_box + (size_t)s * IB::size;
            if (pf_col) cp_async16(smem_u32(&sm.colz[s][pf_sd][pf_row]), box + (e & 1) * IB::col_par + pf_sd * IB::col_side + pf_row);
            if (rows_top || rows_bot) {
                const int side = rows_bot ? 1 : 0;
                const uint4* src = box + IB::row_base + (e & 1) * IB::row_par + side * IB::row_side + lane;
#pragma unroll
                for (int h = 0; h < kHaloY; ++h) cp_async16(smem_u32(&sm.rowz[s][side][h][lane]), src + h * 32);
            }
        };

        // Take the halo ring of refresh epoch e.  Normal case: the prefetch issued half a period ago has landed with
        // current tags.  Otherwise fetch again until the neighbours have delivered (bounded: fail loudly, never hang).
        auto apply = [&](int s, u64 (&A)[P], int e, bool prefetched) {
            const uint32_t tag = round_tag + (uint32_t)e;
            if (!prefetched) prefetch(s, e);
            for (int spin = 0;; ++spin) {
                cp_async_wait_all();
                __syncwarp();
                bool ok = true;
                if (edge) {
#pragma unroll
                    for (int i = 0; i < P; ++i) {
                        const uint4 q = sm.colz[s][edge_sd][warp * P + i];
                        ok = ok && q.y == tag && q.w == tag;
                        A[i] = pk_bits(q.x, q.z);
                    }
                }
                if (rows_top) {                                                   // after the columns: the corner comes with the rows
#pragma unroll
                    for (int h = 0; h < kHaloY; ++h) {
                        const uint4 q = sm.rowz[s][0][h][lane];
                        ok = ok && q.y == tag && q.w == tag;
                        A[h] = pk_bits(q.x, q.z);
                    }
                }
                if (rows_bot) {
#pragma unroll
                    for (int h = 0; h < kHaloY; ++h) {
                        const uint4 q = sm.rowz[s][1][h][lane];
                        ok = ok && q.y == tag && q.w == tag;
                        A[P - kHaloY + h] = pk_bits(q.x, q.z);
                    }
                }
                if (__all_sync(0xffffffffu, ok)) break;
                if (spin > kSpinLimit || poisoned) { poisoned = true; break; }
                __syncwarp();                                                     // everybody has read the landing zone
                prefetch(s, e);
            }
        };

        // ---- T steps, two at a time per slot ----------------------------------------------------------------------
        const int T_ = p.iters;
        bool pfA = false, pfB = false;                      // a prefetch for the slot's next refresh is in flight
        for (int t = 0; t < T_; t += 2) {
            const int e = t >> 1;
            const bool odd = t + 1 < T_, more = t + 2 < T_;
            u64 top, bot;
            // ===== slot A =====
            if (multi && e > 0) { apply(0, A0, e, pfA); pfA = false; }
            exchange_rows(0, 0, A0, top, bot);
            sweep(nwA, A0, ctA, top, bot);
            if (odd) {
                exchange_rows(0, 1, A0, top, bot);
                if (multi && haveB && e > 0) { prefetch(1, e); pfB = true; }      // B's rim messages left its neighbours a period ago
                sweep(nwA, A0, ctA, top, bot);
                if (multi && more) ship(0, A0, e + 1);
            }
            // ===== slot B =====
            if (haveB) {
                if (multi && e > 0) { apply(1, A1, e, pfB); pfB = false; }
                exchange_rows(1, 0, A1, top, bot);
                sweep(nwB, A1, ctB, top, bot);
                if (odd) {
                    exchange_rows(1, 1, A1, top, bot);
                    if (multi && more) { prefetch(0, e + 1); pfA = true; }
                    sweep(nwB, A1, ctB, top, bot);
                    if (multi && more) ship(1, A1, e + 1);
                }
            }
        }

        // ---- epilogue: only the final depth goes back to HBM, from the pixels this tile is authoritative for ----------
        auto store_out = [&](const Geo& g, u64 (&A)[P]) {
            // region of this unit whose results are exact (outside the decaying margin of unit edges inside the image)
            const int vx0 = g.tix > 0 ? g.tix * p.stepx + p.margin : 0;
            const int vx1 = g.tix == p.ntx - 1 ? W : g.tix * p.stepx + p.ew - p.margin;
            const int vy0 = g.tiy > 0 ? g.tiy * p.stepy + p.margin : 0;
            const int vy1 = g.tiy == p.nty - 1 ? H : g.tiy * p.stepy + p.eh - p.margin;
            const int ry0 = hasU ? kHaloY : 0, ry1 = hasD ? TH - 1 - kHaloY : TH - 1;
            const int gx = g.ox + 2 * lane;
            T* ob = p.out + (size_t)g.plane * hw;
            if (poisoned) {
#pragma unroll
                for (int i = 0; i < P; ++i) A[i] = pk(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000));
            }
            if (lane >= (hasL ? 1 : 0) && lane <= (hasR ? 30 : 31)) {
                const bool ok0 = gx >= vx0 && gx < vx1 && gx < W, ok1 = gx + 1 >= vx0 && gx + 1 < vx1 && gx + 1 < W;
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    const int ty = warp * P + i, gy = g.oy + ty;
                    if (ty < ry0 || ty > ry1 || gy < vy0 || gy >= vy1 || gy >= H) continue;
                    const size_t off = (size_t)gy * W + gx;
                    if (ok0 && ok1) {
                        if (sizeof(T) == 4) *reinterpret_cast<float2*>(ob + off) = make_float2(lo_of(A[i]), hi_of(A[i]));
                        else *reinterpret_cast<__half2*>(ob + off) = __floats2half2_rn(lo_of(A[i]), hi_of(A[i]));
                    } else {
                        if (ok0) ob[off] = from_f32<T>(lo_of(A[i]));
                        if (ok1) ob[off + 1] = from_f32<T>(hi_of(A[i]));
                    }
                }
            }
        };
        store_out(geo_of(unit_a), A0);
        if (haveB) store_out(geo_of(unit_b), A1);

        if (multi) {
            // Every message addressed to this CTA in this round has been consumed (the last refresh is followed by at
            // least one CTA barrier): leave the inboxes clean for the round after next / the next launch / graph replay.
            const uint32_t nclean = (haveB ? 2u : 1u) * IB::size;
            for (uint32_t i = threadIdx.x; i < nclean; i += NW * 32) my_box[i] = make_uint4(0, 0, 0, 0);
            __threadfence();
        }
    }
    if (poisoned && p.status) *p.status = 1;
}

// ---- host side -------------------------------------------------------------------------------------------------------
struct DualPlan {
    bool ok;
    int P, cx, cy, ntx, nty, stepx, stepy, ew, eh, per_unit, nA, rounds, grid;
    long units;
    double cost;
};

inline int dual_force_p()
{
    static const int v = [] { const char* e = getenv("CSPN_DUAL_P"); return e ? atoi(e) : 0; }();   // tuning knob: 4 or 5
    return v;
}

// Cost model in SM cycles (calibrated on B200, profiles/r02_*): a round = prologue + epilogue + T/2 periods; a period
// of a two-slot CTA is four half sweeps (56 P cycles each) plus exchange overhead, a one-slot CTA has two half sweeps
// but waits for the message round trip.
DualPlan dual_plan(int H, int W, int iters, long planes, int sms)
{
    DualPlan best{}; best.ok = false;
    if (planes < 1 || iters < 1 || iters > 120 || (W & 1)) return best;
    for (int P = 4; P <= 5; ++P) {
        if (dual_force_p() && dual_force_p() != P) continue;
        const int th = kDNW * P, step_y = th - 2 * kHaloY;
        const int cx_full = W <= kTileW ? 1 : (W - kTileW + kStepX - 1) / kStepX + 1;
        const int cy_full = H <= th ? 1 : (H - th + step_y - 1) / step_y + 1;
        for (int cx = 1; cx <= cx_full; ++cx)
            for (int cy = 1; cy <= cy_full; ++cy) {
                if ((long)cx * cy > sms) continue;
                DualPlan d{}; d.P = P; d.cx = cx; d.cy = cy;
                d.ew = kStepX * (cx - 1) + kTileW; d.eh = step_y * (cy - 1) + th;
                d.ntx = tiles_needed(d.ew, W, iters, &d.stepx);
                d.nty = tiles_needed(d.eh, H, iters, &d.stepy);
                if (d.ntx < 0 || d.nty < 0) continue;
                if (d.ntx == 1 && cx != cx_full) continue;              // a unit that reaches both borders uses the minimal tiling
                if (d.nty == 1 && cy != cy_full) continue;
                if (d.ntx > 1 && (d.stepx & 1)) continue;               // pixel pairs stay at even x
                d.per_unit = cx * cy; d.nA = sms / d.per_unit;
                d.units = planes * d.ntx * d.nty;
                if (d.units > (1l << 30)) continue;
                const long upr = 2l * d.nA;
                const long rounds = (d.units + upr - 1) / upr;
                if (rounds > 16384) continue;                           // round index lives in bits 6..19 of the tag
                d.rounds = (int)rounds;
                const long n0 = d.units < upr ? d.units : upr;
                d.grid = (int)((n0 + 1) / 2) * d.per_unit;
                const double hs = 56.0 * P;
                const double period = d.units >= 2 ? 4 * hs + 600 : 2 * hs + 1100;
                d.cost = (double)rounds * (11000.0 + ((iters + 1) / 2) * period);
                d.ok = true;
                if (!best.ok || d.cost < best.cost - 1e-6 || (d.cost < best.cost + 1e-6 && (long)d.grid * d.rounds < (long)best.grid * best.rounds)) best = d;
            }
    }
    return best;
}

template <int P> constexpr size_t dual_inbox_bytes(int grid) { return (size_t)2 * grid * 2 * InboxGeom<kDNW * P>::size * sizeof(uint4); }
constexpr size_t kDualStatusBytes = 256;

inline size_t dual_ws_bytes(const DualPlan& d)
{
    if (d.per_unit <= 1) return kDualStatusBytes;
    return kDualStatusBytes + (d.P == 4 ? dual_inbox_bytes<4>(d.grid) : dual_inbox_bytes<5>(d.grid));
}

template <typename T, int P, int MODE>
int dual_launch(const FwdArgs<T>& a, const DualPlan& d, const CUtensorMap& map)
{
    constexpr int TH = kDNW * P;
    using St = Stage<T, TH, MODE>;
    constexpr size_t smem = ((sizeof(DualSm<P>) + 127) & ~(size_t)127) + 2 * St::bytes;
    static_assert(smem <= 227 * 1024, "shared memory budget of one SM exceeded");
    static_assert((size_t)TH * 32 * sizeof(u64) <= St::bytes, "re-injection tile must fit its staging buffer");
    auto kern = dual3x3_kernel<T, P, MODE>;
    // the shared-memory opt-in is per device and survives for the life of the context: set it once per device
    static std::atomic<uint64_t> configured{0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return (int)cudaGetLastError();
    const uint64_t bit = 1ull << (dev & 63);
    if (!(configured.load(std::memory_order_acquire) & bit)) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured.fetch_or(bit, std::memory_order_release);
    }
    DualParams<T> p{};
    p.g = a.guidance; p.gbs = a.gbs; p.depth = a.depth; p.sparse = a.sparse; p.sparse_channels = a.sparse_channels; p.out = a.out;
    p.C = a.C; p.H = a.H; p.W = a.W; p.iters = a.iters;
    p.cx = d.cx; p.cy = d.cy; p.ntx = d.ntx; p.nty = d.nty; p.stepx = d.stepx; p.stepy = d.stepy; p.ew = d.ew; p.eh = d.eh;
    p.margin = a.iters; p.per_unit = d.per_unit; p.nA = d.nA; p.total_units = (int)d.units; p.rounds = d.rounds;
    p.status = (int*)a.ws;
    p.inbox = (uint4*)((char*)a.ws + kDualStatusBytes);
    p.tag_base = exchange_epoch().fetch_add(1, std::memory_order_relaxed) << 20;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)d.grid);
    cfg.blockDim = dim3(kDNW * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = a.stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;       // neighbours spin on each other's messages: all CTAs must be resident
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p, map);
    if (e != cudaSuccess) return (int)e;
    ++call_stats().launches;
    return 0;
}

int device_sms()
{
    static int cache[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 148; }
    int& c = cache[dev & 63];
    if (c > 0) return c;
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) { cudaGetLastError(); return 148; }
    c = sms;
    return sms;
}

DualPlan plan_for(int B, int C, int H, int W, int iters) { return dual_plan(H, W, iters, (long)B * C, device_sms()); }

}  // namespace

bool dual_supported(int B, int C, int H, int W, int iters, int ksize, int mode)
{
    (void)mode;
    static const bool off = [] { const char* e = getenv("CSPN_FWD_KERNEL"); return e && !strcmp(e, "single"); }();   // A/B knob
    if (off || ksize != 3 || B < 1) return false;
    if ((long)H * W > (1l << 30)) return false;
    if (((size_t)W * 2) % 16) return false;                 // TMA rows must be 16-byte multiples for both element sizes
    return plan_for(B, C, H, W, iters).ok;
}

// rows per warp, tiles per unit (cx, cy), units per plane (ntx, nty), CTAs, rounds, units per class and round, units
void dual_describe(int B, int C, int H, int W, int iters, int* out9)
{
    const DualPlan d = plan_for(B, C, H, W, iters);
    const int v[9] = {d.P, d.cx, d.cy, d.ntx, d.nty, d.grid, d.rounds, d.nA, (int)d.units};
    for (int i = 0; i < 9; ++i) out9[i] = d.ok ? v[i] : 0;
}

size_t dual_workspace(int B, int C, int H, int W, int iters)
{
    const DualPlan d = plan_for(B, C, H, W, iters);
    return d.ok ? dual_ws_bytes(d) : 0;
}

// CSPN_ERR_UNALIGNED_FALLBACK (internal): the guidance cannot be described by a TMA tensor map; the caller takes the
// single-tile kernel, which has a plain-load prologue.
template <typename T>
int dual_forward(const FwdArgs<T>& a)
{
    const DualPlan d = plan_for(a.B, a.C, a.H, a.W, a.iters);
    if (!d.ok) return kDualFallback;
    if (!a.ws || a.ws_bytes < dual_ws_bytes(d)) return CSPN_ERR_WORKSPACE;
    alignas(64) CUtensorMap map;
    memset(&map, 0, sizeof map);
    bool tma;
    if (a.mode == CSPN_MODE_NEW) tma = d.P == 4 ? make_guidance_map<T, 32, CSPN_MODE_NEW>(a.guidance, a.gbs, a.B, a.H, a.W, &map) : make_guidance_map<T, 40, CSPN_MODE_NEW>(a.guidance, a.gbs, a.B, a.H, a.W, &map);
    else tma = d.P == 4 ? make_guidance_map<T, 32, CSPN_MODE_OURS>(a.guidance, a.gbs, a.B, a.H, a.W, &map) : make_guidance_map<T, 40, CSPN_MODE_OURS>(a.guidance, a.gbs, a.B, a.H, a.W, &map);
    if (!tma) return kDualFallback;
    if (a.mode == CSPN_MODE_NEW) return d.P == 4 ? dual_launch<T, 4, CSPN_MODE_NEW>(a, d, map) : dual_launch<T, 5, CSPN_MODE_NEW>(a, d, map);
    return d.P == 4 ? dual_launch<T, 4, CSPN_MODE_OURS>(a, d, map) : dual_launch<T, 5, CSPN_MODE_OURS>(a, d, map);
}

template int dual_forward<float>(const FwdArgs<float>&);
template int dual_forward<__half>(const FwdArgs<__half>&);

}  // namespace cspn
