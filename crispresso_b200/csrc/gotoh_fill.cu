// gotoh_fill.cu -- k_gotoh_fill<G,K>: batched Gotoh fill (needle semantics) for sm_100a.
//
// Replaces the DP of EMBOSS needle's embAlignPathCalcWithEndGapPenalties as CRISPResso runs it
// (CRISPResso/CRISPRessoCORE.py:1791-1806; algorithm: SURVEY.md App. A.1-A.3, exact integer
// form A.6).
//
// Work decomposition (DESIGN.md "k_gotoh_fill"):
//   * one GROUP of G lanes owns one PAIR of reads (equal length, packed in the two 16-bit
//     halves of every register) and sweeps the read columns x = 0..Lb-1;
//   * lane t of the group owns K consecutive amplicon rows (virtual rows r = t*K .. t*K+K-1;
//     the amplicon is padded on TOP with P = G*K - La rows scoring 0 against everything, which
//     reproduces needle's free-end-gap boundary exactly), whose state H3 = max(m,ix,iy)[y,x-1] and
//     IX = ix[y,x-1] lives in 2K registers;
//   * the lanes form a systolic pipeline: at step s lane t works on column x = s - t and hands
//     (max3, iy, m) of its bottom row to lane t+1 with one warp shuffle each;
//   * the substitution scores of a column come from a pair profile table in shared memory
//     (25 read-code pairs x padded rows, one LDS.128 per 4 rows), loaded once per CTA with a
//     TMA bulk copy (cp.async.bulk -> UBLKCP);
//   * per cell 5 "not equal" flag bits (crgpu_common.cuh) are packed into one byte and written
//     with 16-byte vector stores; needle's start-cell scan over the last amplicon row and the last
//     read column is folded into the sweep (running first-maximum per lane), so only 12 bytes per
//     pair and per lane leave the kernel for it.
// The kernel is persistent: grid = SMs x resident CTAs, groups stride over the batch's pairs.
//
// Banded two-pass variant (FillArgs.band_B > 0; DESIGN.md "Band"): 16 of the ~22 instructions per cell
// pair produce the traceback flags, but a traceback only ever reads the cells on its path, and the path
// of an amplicon read hugs the main diagonal.  k_gotoh_score2<G,K> (gotoh_score2.cu) therefore evaluates every cell
// WITHOUT flags (scores, start-cell scan) and saves, per lane, what it received from the lane above at
// the columns of a diagonal band and its register state at the band's left edge; k_gotoh_band<G,K>
// re-evaluates only the band columns of every lane with flags -- lanes are independent there, the top
// boundary comes from memory.  The walker raises an escape flag for a read whose path leaves the band
// and the host re-aligns those reads with the full single-pass kernel: results never depend on the band.
#include "gotoh_tile.cuh"
#include <cstdio>
#include <cstdlib>
#include <type_traits>

namespace crgpu {

// One read column for the K rows of this lane.
//
// Recurrences in the form used here, valid because gapopen >= gapextend (checked
// on the host): the "open" source max(m,iy)[y,x-1] of ix may be replaced by max3[y,x-1] -- if ix is
// the strict maximum then ix-open <= ix-ext, so the result is unchanged -- and likewise for iy.
// That leaves per cell: m = S + max3[y-1,x-1]; ix = max(max3[y,x-1]-open, ix[y,x-1]-ext);
// iy = max(max3[y-1,x]-open, iy[y-1,x]-ext); max3 = VIMNMX3(m,ix,iy).
//
//  EDGE = false: interior columns, all penalties are immediates.
//  EDGE = true : some lane of the warp is on its first (x == 0) or last (x == Lb-1) column, whose
//  iy rule / FY flag differ (App. A.2/A.3: the last column opens from m only with zero penalties;
//  A.4: gey = 0 on both).  The whole warp then runs this body with per-lane sources/penalties, so
//  the special columns cost no divergence.
template <int K, bool EDGE, bool FLAGS, bool SCAN>
__device__ __forceinline__ void column_step(Strip<K> &st, const int32_t *__restrict__ prow,
                                            uint32_t upH3, uint32_t upIY, uint32_t upM, uint32_t hd,
                                            const uint32_t nopen16, const uint32_t ext32,
                                            const uint32_t nopen16_last, const uint32_t ext32_last,
                                            const bool lastLane, const bool isFirstCol, const bool isLastCol, const uint32_t one,
                                            uint32_t *__restrict__ tbw, const int firstRealSlot,
                                            uint32_t &colBest, int &colPosLo, int &colPosHi,
                                            uint32_t &botH3, uint32_t &botIY, uint32_t &botM)
{
    uint32_t words[K / 2];
    uint32_t ceven = 0;
    int32_t S4[4];
    // the profile row is read 4 rows ahead (LDS latency ~30 cycles would otherwise sit on the m chain)
    int4 Snext = *reinterpret_cast<const int4 *>(prow);
    // per-lane vertical-gap parameters of this column (EDGE only)
    const uint32_t nopen16_v = (EDGE && isLastCol) ? 0u : nopen16;
    const uint32_t ext32_v = (EDGE && isLastCol) ? 0u : ext32;
    const uint32_t ext32_f = (EDGE && (isLastCol || isFirstCol)) ? 0u : ext32;   // gey of the FY flag
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if ((k & 3) == 0) {
            S4[0] = Snext.x; S4[1] = Snext.y; S4[2] = Snext.z; S4[3] = Snext.w;
            if (k + 4 < K) Snext = *reinterpret_cast<const int4 *>(prow + k + 4);
        }
        const uint32_t S = (uint32_t)S4[k & 3];
        const uint32_t h0 = st.H3[k], ix0 = st.IX[k];
        const uint32_t m = hd + S;                       // m = sub + max3[row-1, x-1]
        uint32_t t_, ix;
        if (k == K - 1) {                                // the only row that can be amplicon row La-1
            const uint32_t src = lastLane ? st.mlast : h0;
            t_ = ix0 - ext32_last;
            ix = vaddmax2(src, nopen16_last, t_);
        } else {
#if defined(FILL_FMA_ADDS) && FILL_FMA_ADDS >= 2
            t_ = fma_add(ix0, 0u - ext32, one);
#else
            t_ = ix0 - ext32;
#endif
            ix = vaddmax2(h0, nopen16, t_);
        }
        uint32_t iy, nFY = 0;
        if (EDGE) {
            const uint32_t src = isLastCol ? upM : upH3;
            iy = vaddmax2(src, nopen16_v, upIY - ext32_v);
            if (FLAGS) nFY = vmin2(iy ^ (upIY - ext32_f), ONE2);
        } else {
#ifdef FILL_FMA_ADDS
            const uint32_t u_ = fma_add(upIY, 0u - ext32, one);
#else
            const uint32_t u_ = upIY - ext32;
#endif
            iy = vaddmax2(upH3, nopen16, u_);
            if (FLAGS) nFY = vmin2(iy - u_, ONE2);
        }
        const uint32_t h3 = __vimax3_s16x2(m, ix, iy);
        if (FLAGS) {
            const uint32_t nM = vmin2(h3 - m, ONE2);
            const uint32_t nX = vmin2(h3 - ix, ONE2);
            const uint32_t nY = vmin2(h3 - iy, ONE2);
            const uint32_t nFX = vmin2(ix - t_, ONE2);
            const uint32_t c = nM + 2u * nX + 4u * nY + 8u * nFX + 16u * nFY;
            if (k & 1) words[k >> 1] = ceven + (c << 8);
            else ceven = c;
        }
        if (EDGE && SCAN) {
            // start-cell scan down the last read column (App. A.4): first row whose max(m,ix,iy) is
            // strictly greater than everything above it; padded rows are not part of the matrix
            if (isLastCol && k >= firstRealSlot) {
                const uint32_t nb = vmax2(colBest, h3);
                const uint32_t d = nb ^ colBest;
                if (d & 0xffffu) colPosLo = k;
                if (d >> 16) colPosHi = k;
                colBest = nb;
            }
        }
        st.H3[k] = h3;
        st.IX[k] = ix;
        if (k == K - 1) st.mlast = m;
        upH3 = h3; upIY = iy; upM = m; hd = h0;
    }
    botH3 = upH3; botIY = upIY; botM = upM;
    if constexpr (!FLAGS) return;
    if constexpr ((K % 8) == 0) {
#pragma unroll
        for (int j = 0; j < K / 8; ++j)
            reinterpret_cast<uint4 *>(tbw)[j] = make_uint4(words[4 * j], words[4 * j + 1], words[4 * j + 2], words[4 * j + 3]);
    } else {
#pragma unroll
        for (int j = 0; j < K / 4; ++j)
            reinterpret_cast<uint2 *>(tbw)[j] = make_uint2(words[2 * j], words[2 * j + 1]);
    }
}

template <int G, int K>
__global__ void __maxnreg__(fill_maxnreg<K>()) k_gotoh_fill(const FillArgs a)
{
    static_assert(K % 4 == 0 && (32 % G) == 0, "bad tile");
    constexpr int PS = prof_stride(G, K);
    constexpr int GK = G * K;
    constexpr bool FLAGS = true;
    extern __shared__ __align__(128) int32_t sprof[];
    __shared__ __align__(8) uint64_t mbar;
    stage_profile(sprof, &mbar, a.prof, NPAIR * PS * 4);

    const int lane = threadIdx.x & 31;
    const int t = lane % G;
    const int gl = lane / G;
    constexpr int GPW = 32 / G;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const bool lastLane = (t == G - 1);

    const uint32_t Z = BIAS2;                                             // stored 0
    const uint32_t NOPEN_ST = BIAS2 - (uint32_t)a.open * 0x10001u;        // stored -open
    const uint32_t nopen16 = ((uint32_t)(-a.open) & 0xffffu) * 0x10001u;  // per-half two's complement
    const uint32_t ext32 = (uint32_t)a.ext * 0x10001u;
    const uint32_t nopen16_last = lastLane ? 0u : nopen16;                // amplicon row La-1: zero end-gap penalties
    const uint32_t ext32_last = lastLane ? 0u : ext32;
    const uint32_t one = (uint32_t)a.one;

    for (int base = a.p0 + warp_global * GPW; base < a.p1; base += nwarps * GPW) {
        const int p = base + gl;
        const bool valid = p < a.p1;
        const int Lb = valid ? a.plen[p] : 0;
        const int steps = __reduce_max_sync(0xffffffffu, Lb) + G - 1;
        const int64_t pco = valid ? a.pc_off[p] : 0;
        const int64_t pco_rel = valid ? pco - a.pc_off[a.p0] : 0;
        const uint8_t *pcp = a.pc + pco;
        uint32_t *tbp = FLAGS ? a.tb + pco_rel * (GK / 2) + t * (K / 2) : nullptr;
        uint32_t *lrp = a.lastrow + (int64_t)(p - a.p0) * 3;             // (best, x_lo, x_hi) of amplicon row La-1
        uint32_t *lcp = a.lastcol + ((int64_t)(p - a.p0) * G + t) * 3;   // (best, slot_lo, slot_hi) of this lane's rows, column Lb-1
        const int firstRealSlot = (GK - a.La) - t * K;                   // slots below it are padding rows

        Strip<K> st;
#pragma unroll
        for (int k = 0; k < K; ++k) { st.H3[k] = Z; st.IX[k] = NOPEN_ST; }
        st.mlast = Z;
        uint32_t botH3 = Z, botIY = NOPEN_ST, botM = Z;
        uint32_t hd0 = Z;
        uint32_t rowBest = 0, colBest = 0;                               // stored scores are > 0: 0 is -infinity
        int rowPosLo = 0, rowPosHi = 0, colPosLo = 0, colPosHi = 0;
        int cp_next = (t == 0 && Lb > 0) ? pcp[0] : 0;
        // top boundary from the pass that owns the rows above (shared DP prefix), read one column ahead
        const int64_t tcol = top_base_col(pco_rel, p - a.p0);
        const uint4 *tin = (a.top_in && valid) ? reinterpret_cast<const uint4 *>(a.top_in) + tcol : nullptr;
        uint4 *tout = (a.top_out && valid && t == a.top_out_lane) ? reinterpret_cast<uint4 *>(a.top_out) + tcol : nullptr;
        uint4 tn = make_uint4(Z, NOPEN_ST, Z, 0u);
        if (tin && t == 0 && Lb > 0) tn = tin[0];
        // One systolic step.  STEADY = every lane of the warp is on an interior column of its read (no lane
        // idle, none on a first or last column): the votes, the activity branch and the edge body drop out.
        auto step = [&](auto steady_tag, const int s) {
            constexpr bool STEADY = decltype(steady_tag)::value;
            const int x = s - t;
            uint32_t rH3 = __shfl_up_sync(0xffffffffu, botH3, 1, G);
            uint32_t rIY = __shfl_up_sync(0xffffffffu, botIY, 1, G);
            uint32_t rM = __shfl_up_sync(0xffffffffu, botM, 1, G);
            if (t == 0) { rH3 = tn.x; rIY = tn.y; rM = tn.z; }        // free boundary above the padded top, or the saved row
            if (tin && t == 0 && x + 1 >= 0 && x + 1 < Lb) tn = tin[x + 1];
            const bool active = STEADY || ((x >= 0) && (x < Lb));
            const bool firstCol = !STEADY && active && x == 0, lastCol = !STEADY && active && x == Lb - 1;
            const bool edge = !STEADY && __any_sync(0xffffffffu, firstCol || lastCol);   // warp-uniform
            const int cp = cp_next;
            if (STEADY || (x + 1 >= 0 && x + 1 < Lb)) cp_next = pcp[x + 1];
            if (active) {
                const int32_t *prow = sprof + cp * PS + t * strip_stride(K);
                uint32_t *tbw = FLAGS ? tbp + (int64_t)x * (GK / 2) : nullptr;
                if (edge)
                    column_step<K, true, FLAGS, true>(st, prow, rH3, rIY, rM, hd0, nopen16, ext32, nopen16_last, ext32_last,
                                                      lastLane, firstCol, lastCol, one, tbw, firstRealSlot, colBest, colPosLo, colPosHi,
                                                      botH3, botIY, botM);
                else
                    column_step<K, false, FLAGS, true>(st, prow, rH3, rIY, rM, hd0, nopen16, ext32, nopen16_last, ext32_last,
                                                       lastLane, false, false, one, tbw, firstRealSlot, colBest, colPosLo, colPosHi,
                                                       botH3, botIY, botM);
                hd0 = rH3;                                            // max3[row above, x] for column x+1
                // start-cell scan along the last amplicon row (meaningful in the last lane only): first
                // column whose max(m,ix,iy) is strictly greater than all columns before it
                {
                    const uint32_t nb = vmax2(rowBest, botH3);
                    const uint32_t d = nb ^ rowBest;
                    if (d & 0xffffu) rowPosLo = x;
                    if (d >> 16) rowPosHi = x;
                    rowBest = nb;
                }
                if (tout) tout[x] = make_uint4(botH3, botIY, botM, 0u);
                if (lastCol) {
                    lcp[0] = colBest; lcp[1] = (uint32_t)colPosLo; lcp[2] = (uint32_t)colPosHi;
                    if (lastLane) { lrp[0] = rowBest; lrp[1] = (uint32_t)rowPosLo; lrp[2] = (uint32_t)rowPosHi; }
                }
            }
        };
        // steps G .. Lmin-2 are steady for the whole warp (lane t is on column s - t)
        const int Lmin = __reduce_min_sync(0xffffffffu, Lb);
        const int steady_end = max(min(Lmin - 1, steps), min(G, steps));      // first non-steady step after the steady run
        int s = 0;
        for (; s < min(G, steps); ++s) step(std::false_type{}, s);
        for (; s < steady_end; ++s) step(std::true_type{}, s);
        for (; s < steps; ++s) step(std::false_type{}, s);
    }
}

// Second pass of the banded fill: one thread re-evaluates, with flags, columns xlo .. xlo+W-1 of one SUB-STRIP of
// Kb = K / NSUB rows (sub-strip u = t*NSUB + hh of lane t's strip).  Sub-strips are independent (top boundary from
// band_tops / top_in / the free boundary, left edge from band_left), so there are no shuffles, threads map to
// (pair, sub-strip) linearly and every thread of a warp runs the same W iterations.  Halving the strip height
// (NSUB = 2) shrinks the rectangle that covers the diagonal band from K + 2B + 1 to K/2 + 2B + 1 columns.
template <int G, int K, int NSUB>
__global__ void __maxnreg__(fill_maxnreg<K / NSUB>()) k_gotoh_band(const FillArgs a)
{
    constexpr int PS = prof_stride(G, K);
    constexpr int Kb = K / NSUB, G2 = G * NSUB;
    static_assert(Kb % 8 == 0 || NSUB == 1, "sub-strips store 16-byte flag pieces");
    extern __shared__ __align__(128) int32_t sprof[];
    __shared__ __align__(8) uint64_t mbar;
    stage_profile(sprof, &mbar, a.prof, NPAIR * PS * 4);

    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t Z = BIAS2;
    const uint32_t NOPEN_ST = BIAS2 - (uint32_t)a.open * 0x10001u;
    const uint32_t nopen16 = ((uint32_t)(-a.open) & 0xffffu) * 0x10001u;
    const uint32_t ext32 = (uint32_t)a.ext * 0x10001u;
    const uint32_t one = (uint32_t)a.one;
    const int W = a.band_W;
    const uint4 FREE = make_uint4(Z, NOPEN_ST, Z, 0u);
    // (diagonal shortcut: only the pairs that still need a traceback, FillArgs.pair_list)
    // (... and, with sub_n > 0, only sub-strips sub_lo .. sub_lo + sub_n - 1 of those pairs)
    const int G2run = a.sub_n > 0 ? a.sub_n : G2;
    const int64_t total = (int64_t)(a.pair_list ? *a.pair_list_n : a.p1 - a.p0) * G2run;

    for (int64_t base = (int64_t)warp_global * 32; base < total; base += (int64_t)nwarps * 32) {
        const int64_t slot_id = base + lane;                              // (position in the list) * G2 + u
        const bool valid = slot_id < total;
        const int pj = valid ? (int)(slot_id / G2run) : 0;
        const int u = valid ? a.sub_lo + (int)(slot_id - (int64_t)pj * G2run) : 0;
        const int pr = (valid && a.pair_list) ? a.pair_list[pj] : pj;     // pair - p0
        const int64_t sub_id = (int64_t)pr * G2 + u;
        const int t = u / NSUB, hh = u - t * NSUB;
        const bool lastLane = (u == G2 - 1);                              // owns amplicon row La-1 in its last slot
        const uint32_t nopen16_last = lastLane ? 0u : nopen16;
        const uint32_t ext32_last = lastLane ? 0u : ext32;
        const int p = a.p0 + pr;
        const int Lb = valid ? a.plen[p] : 0;
        const int64_t pco = valid ? a.pc_off[p] : 0;
        const int64_t pco_rel = valid ? pco - a.pc_off[a.p0] : 0;
        const uint8_t *pcp = a.pc + pco;
        const int xlo = a.band_row0 + u * Kb - a.band_B;
        const int x0 = max(xlo, 0), x1 = min(xlo + W - 1, Lb - 1);       // this sub-strip's columns; empty when x0 > x1
        // source of the top boundary: the row above (saved by the score pass), the pass that owns the rows above the
        // sub-tile, or the free boundary.  Indexed so that src[x] is column x.
        // Both come from the score pass in its drift coordinates: value + ext * (srow + x), srow = the row's index in the
        // frame of the pass that wrote it (u*Kb - 1 here; a.top_in_row for the shared-prefix row).
        const uint4 *src = nullptr;
        int srow = 0;
        if (u > 0) { src = reinterpret_cast<const uint4 *>(a.band_tops) + (valid ? sub_id : 0) * band_topw(W) - (xlo - 2); srow = u * Kb - 1; }
        else if (a.top_in) { src = reinterpret_cast<const uint4 *>(a.top_in) + top_base_col(pco_rel, pr); srow = a.top_in_row; }
        auto plain = [&](uint4 v, int x) { const uint32_t d = ext32 * (uint32_t)(srow + x); return make_uint4(v.x - d, v.y - d, v.z - d, 0u); };
        uint32_t *tbl = a.band_tb + (valid ? sub_id : 0) * W * (Kb / 2);
        const int32_t *pbase = sprof + t * strip_stride(K) + hh * Kb;

        Strip<Kb> st;
        uint32_t hd0 = Z;
        if (x0 > 0 && x0 <= x1) {
            const uint4 *lp = reinterpret_cast<const uint4 *>(a.band_left + sub_id * band_leftw(Kb));
#pragma unroll
            for (int j = 0; j < Kb / 4; ++j) {
                const uint4 v = lp[j], w = lp[Kb / 4 + j];
                st.H3[4 * j] = v.x; st.H3[4 * j + 1] = v.y; st.H3[4 * j + 2] = v.z; st.H3[4 * j + 3] = v.w;
                st.IX[4 * j] = w.x; st.IX[4 * j + 1] = w.y; st.IX[4 * j + 2] = w.z; st.IX[4 * j + 3] = w.w;
            }
            st.mlast = lp[Kb / 2].x;
            if (src) hd0 = src[x0 - 1].x - ext32 * (uint32_t)(srow + x0 - 1);
        } else {
#pragma unroll
            for (int k = 0; k < Kb; ++k) { st.H3[k] = Z; st.IX[k] = NOPEN_ST; }
            st.mlast = Z;
        }
        uint32_t botH3, botIY, botM, colBest = 0;
        int colPosLo = 0, colPosHi = 0;
        uint4 rn = FREE;
        int cp_next = 0;
        if (x0 <= x1) { if (src) rn = load_top3(src + x0); cp_next = pcp[x0]; }

        for (int i = 0; i < W; ++i) {
            const int x = xlo + i;
            const bool active = x >= x0 && x <= x1;
            const uint4 r = src ? plain(rn, x) : rn;             // (un-drifted at use, one column after the load)
            const int cp = cp_next;
            if (x + 1 >= x0 && x + 1 <= x1) { if (src) rn = load_top3(src + x + 1); cp_next = pcp[x + 1]; }
            const bool firstCol = active && x == 0, lastCol = active && x == Lb - 1;
            const bool edge = __any_sync(0xffffffffu, firstCol || lastCol);   // warp-uniform
            if (active) {
                const int32_t *prow = pbase + cp * PS;
                uint32_t *tbw = tbl + (int64_t)i * (Kb / 2);
                if (edge)
                    column_step<Kb, true, true, false>(st, prow, r.x, r.y, r.z, hd0, nopen16, ext32, nopen16_last, ext32_last,
                                                       lastLane, firstCol, lastCol, one, tbw, 0, colBest, colPosLo, colPosHi,
                                                       botH3, botIY, botM);
                else
                    column_step<Kb, false, true, false>(st, prow, r.x, r.y, r.z, hd0, nopen16, ext32, nopen16_last, ext32_last,
                                                        lastLane, false, false, one, tbw, 0, colBest, colPosLo, colPosHi,
                                                        botH3, botIY, botM);
                hd0 = r.x;
            }
        }
    }
}

// ---- host-side dispatch ----------------------------------------------------------------
struct Tile { int G, K; };

// (attribute + occupancy per call: both are per DEVICE, a process may hold contexts on several devices and drive them
// from several threads, and the two runtime calls cost microseconds)
template <typename Kern>
static cudaError_t launch_persistent(Kern kern, size_t smem, int64_t blocks_wanted, const FillArgs &a, int num_sms, cudaStream_t stream)
{
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int blocks_per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, 128, smem);
    if (e != cudaSuccess) return e;
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    int64_t grid = blocks_wanted;
    const int cap = num_sms * blocks_per_sm;            // persistent: a multiple of the SM count
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    kern<<<(int)grid, 128, smem, stream>>>(a);
    return cudaGetLastError();
}

// kind 0: single pass with flags; kind 2: band pass on sub-strips of a.band_K rows (K, or K/2 when K % 16 == 0)
template <int G, int K>
static cudaError_t launch_tile(const FillArgs &a, int num_sms, cudaStream_t stream, int kind)
{
    const size_t smem = (size_t)NPAIR * prof_stride(G, K) * 4;
    const int npairs = a.p1 - a.p0;
    if (kind == 0) {
        const int groups_per_block = 4 * (32 / G);
        return launch_persistent(k_gotoh_fill<G, K>, smem, (npairs + groups_per_block - 1) / groups_per_block, a, num_sms, stream);
    }
    if (a.band_K == K) return launch_persistent(k_gotoh_band<G, K, 1>, smem, ((int64_t)npairs * G + 127) / 128, a, num_sms, stream);
    if constexpr (K % 16 == 0) {
        if (2 * a.band_K == K) return launch_persistent(k_gotoh_band<G, K, 2>, smem, ((int64_t)npairs * G * 2 + 127) / 128, a, num_sms, stream);
    }
    return cudaErrorInvalidValue;
}

// Smallest padded tile G*K >= La from the compiled menu.
bool choose_tile(int La, int *G, int *K)
{
    // strip heights with K % 8 == 0 only: the 8-byte-store variants (K = 20, 28, 36) measured 2.2x
    // slower per cell on B200 (profiles/r01_notes.md)
    static const Tile menu[] = {
        {4, 16}, {4, 24}, {4, 32}, {4, 40}, {4, 48}, {8, 32}, {8, 40}, {8, 48}, {16, 32}, {16, 40}, {32, 24}, {32, 32},
    };
    if (getenv("CRGPU_TILE")) { int g, k; if (sscanf(getenv("CRGPU_TILE"), "%d,%d", &g, &k) == 2 && g * k >= La) { *G = g; *K = k; return true; } }
    int best = -1, bestgk = 1 << 30;
    for (unsigned i = 0; i < sizeof(menu) / sizeof(menu[0]); ++i) {
        const int gk = menu[i].G * menu[i].K;
        if (gk >= La && gk < bestgk) { bestgk = gk; best = (int)i; }
    }
    if (best < 0) return false;
    *G = menu[best].G; *K = menu[best].K;
    return true;
}

// is there a compiled kernel for this tile?
bool tile_available(int G, int K)
{
    static const Tile all[] = {{4, 16}, {4, 24}, {4, 32}, {4, 40}, {4, 48}, {8, 16}, {8, 24}, {8, 32}, {8, 40}, {8, 48},
                               {16, 16}, {16, 24}, {16, 32}, {16, 40}, {16, 48}, {32, 24}, {32, 32}};
    for (const Tile &t : all) if (t.G == G && t.K == K) return true;
    return false;
}

// kind: 0 = full single pass (flags for every cell), 1 = score pass of the banded fill, 2 = band pass
cudaError_t launch_fill(int G, int K, const FillArgs &a, int num_sms, cudaStream_t stream, int kind)
{
    if (kind < 0 || kind > 2) return cudaErrorInvalidValue;
    if (kind == 1) return launch_score2(G, K, a, num_sms, stream);
#define CASE(g, k) if (G == g && K == k) return launch_tile<g, k>(a, num_sms, stream, kind);
    CASE(4, 16) CASE(4, 24) CASE(4, 32) CASE(4, 40) CASE(8, 16) CASE(8, 24) CASE(8, 32) CASE(8, 40)
    CASE(16, 16) CASE(16, 24) CASE(16, 32) CASE(16, 40) CASE(32, 24) CASE(32, 32) CASE(4, 48) CASE(8, 48) CASE(16, 48)
#undef CASE
    return cudaErrorInvalidValue;
}

}  // namespace crgpu
