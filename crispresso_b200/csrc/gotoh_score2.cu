// gotoh_score2.cu -- k_gotoh_score2<G,K,NSUB>: score pass of the banded two-pass fill, TWO read columns per
// systolic step (DESIGN.md "Score pass").
//
// Same DP, same drift coordinates, same outputs as the one-column kernel it replaces (needle's
// embAlignPathCalcWithEndGapPenalties as CRISPResso runs it, CRISPResso/CRISPRessoCORE.py:1791-1806; SURVEY.md
// App. A.1-A.3, exact integer form A.6):
//
//     v'[r,x] = v[r,x] + ext * (r + x)
//     m'  = (S + 2 ext) + max3'[r-1,x-1]
//     ix' = max(max3'[r,x-1] + (ext - open), ix'[r,x-1])       VIADDMNMX.S16x2
//     iy' = max(max3'[r-1,x] + (ext - open), iy'[r-1,x])       VIADDMNMX.S16x2
//     max3' = VIMNMX3.S16x2(m', ix', iy')
//
// The one-column kernel was bound by the dependent chain iy' -> max3' -> iy' down the K rows of a strip: two
// instructions of 4 cycles each per row, with three warps per scheduler to hide them (0.42 of the integer-ALU
// ceiling, ncu: 31 % of the stall samples fixed-latency waits).  Here lane t works on read columns 2j and 2j + 1 in
// the same step, the second column one row behind the first:
//
//     iteration k:   A = cell (row k, column 2j)        B = cell (row k - 1, column 2j + 1)
//
// A and B are independent (B needs A's results of iterations k - 1 and k - 2), so every thread carries two chains
// and the eight instructions of an iteration fill the eight cycles of its critical path.  Column 2j's values live
// in a three-row window of registers; the state arrays H3 / IX go from column 2j - 1 straight to column 2j + 1, in
// place.  The per-step overhead (shuffles, addresses, boundary loads, scans) is paid once per two columns.
//
// What leaves the kernel is unchanged: start-cell summaries (last amplicon row / last read column), the rows
// handed to the band pass (band_tops), the registers at the band's left edge (band_left), the shared-prefix row
// (top_out) -- all converted back to plain values.  The band's left edge is always an odd column (the host makes
// P + B even, run_plan_band), i.e. the second column of a step, so the registers to save are the state arrays.
#include "gotoh_tile.cuh"
#include <algorithm>
#include <type_traits>

namespace crgpu {

// 16 bytes of the profile table by 32-bit shared address (the table is read-only after staging: not volatile)
__device__ __forceinline__ int4 lds128(uint32_t addr)
{
    int4 v;
    asm("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// Two read columns (x0 = 2j, x1 = 2j + 1) for the K rows of this lane, drift coordinates.
//  TAIL = false: both columns are interior columns of the read for every lane of the warp.
//  TAIL = true : per-lane flags -- last0 / last1: that column is the read's LAST column (iy opens from m only, zero
//  penalties: SURVEY App. A.2/A.3); do1 = false: column x1 is past the read's end (odd length), the state then keeps
//  column x0's values.  The whole warp runs this body with per-lane parameters, no divergence.
template <int K, bool TAIL, int NSUB>
__device__ __forceinline__ void score_columns2(Strip<K> &st, const uint32_t prow0, const uint32_t prow1,   // shared-memory addresses
                                               uint32_t aH, uint32_t aY, uint32_t aM,      // row above, column x0
                                               uint32_t bH, uint32_t bY, uint32_t bM,      // row above, column x1
                                               const uint32_t hd,                          // max3'[row above, x0 - 1]
                                               const uint32_t cOpen, const uint32_t cA_last, const uint32_t cB_last,
                                               const uint32_t e32, const bool lastLane,
                                               const bool last0, const bool last1, const bool do1,
                                               uint32_t (&bot0)[3], uint32_t (&bot1)[3], uint32_t (&mid0)[3], uint32_t (&mid1)[3])
{
    const uint32_t cV0 = (TAIL && last0) ? e32 : cOpen, bV0 = (TAIL && last0) ? e32 : 0u;
    const uint32_t cV1 = (TAIL && last1) ? e32 : cOpen, bV1 = (TAIL && last1) ? e32 : 0u;
    uint32_t dg1 = aH;                       // max3'[row j - 1, x0]: diagonal source of column x1's row j
    uint32_t h0p = 0, ix0p = 0, m0p = 0;     // column x0, row k - 1
    int4 S0 = make_int4(0, 0, 0, 0), S1 = make_int4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k <= K; ++k) {
        uint32_t h0 = 0, ix0 = 0, m0 = 0, iy0 = 0;
        if (k < K) {                                                     // A: row k of column x0
            if ((k & 3) == 0) S0 = lds128(prow0 + k * 4);
            const int32_t s = (k & 3) == 0 ? S0.x : (k & 3) == 1 ? S0.y : (k & 3) == 2 ? S0.z : S0.w;
            m0 = (k == 0 ? hd : st.H3[k - 1]) + (uint32_t)s;
            if (k == K - 1) {                                            // the only slot that can be amplicon row La-1
                const uint32_t src = lastLane ? st.mlast : st.H3[k];
                ix0 = vaddmax2(src, cA_last, st.IX[k] + cB_last);
            } else {
                ix0 = vaddmax2(st.H3[k], cOpen, st.IX[k]);
            }
            if (TAIL) iy0 = vaddmax2(last0 ? aM : aH, cV0, aY + bV0);
            else iy0 = vaddmax2(aH, cOpen, aY);
            h0 = __vimax3_s16x2(m0, ix0, iy0);
        }
        if (k >= 1) {                                                    // B: row j = k - 1 of column x1
            const int j = k - 1;
            if ((j & 3) == 0) S1 = lds128(prow1 + j * 4);
            const int32_t s = (j & 3) == 0 ? S1.x : (j & 3) == 1 ? S1.y : (j & 3) == 2 ? S1.z : S1.w;
            const uint32_t m1 = dg1 + (uint32_t)s;
            uint32_t ix1;
            if (j == K - 1) {
                const uint32_t src = lastLane ? m0p : h0p;
                ix1 = vaddmax2(src, cA_last, ix0p + cB_last);
            } else {
                ix1 = vaddmax2(h0p, cOpen, ix0p);
            }
            uint32_t iy1;
            if (TAIL) iy1 = vaddmax2(last1 ? bM : bH, cV1, bY + bV1);
            else iy1 = vaddmax2(bH, cOpen, bY);
            const uint32_t h1 = __vimax3_s16x2(m1, ix1, iy1);
            if (TAIL) {
                st.H3[j] = do1 ? h1 : h0p;
                st.IX[j] = do1 ? ix1 : ix0p;
                if (j == K - 1) st.mlast = do1 ? m1 : m0p;
            } else {
                st.H3[j] = h1;
                st.IX[j] = ix1;
                if (j == K - 1) st.mlast = m1;
            }
            bH = h1; bY = iy1; bM = m1;
            if (NSUB == 2 && j == K / 2 - 1) { mid1[0] = h1; mid1[1] = iy1; mid1[2] = m1; }
            dg1 = h0p;
        }
        if (k < K) {
            aH = h0; aY = iy0; aM = m0;
            if (NSUB == 2 && k == K / 2 - 1) { mid0[0] = h0; mid0[1] = iy0; mid0[2] = m0; }
            h0p = h0; ix0p = ix0; m0p = m0;
        }
    }
    bot0[0] = aH; bot0[1] = aY; bot0[2] = aM;
    bot1[0] = bH; bot1[1] = bY; bot1[2] = bM;
}

#ifndef SCORE2_MAXNREG
#define SCORE2_MAXNREG 152
#endif
template <int K> constexpr int score2_maxnreg() { return K <= 32 ? SCORE2_MAXNREG : fill_maxnreg<K>(); }

// k_gotoh_score2 leaves everything it saves per column in DRIFT coordinates (no subtraction in the step):
//   band_tops  (row above sub-strip u, column x)  : + ext * (u*Kb - 1 + x)      un-drifted by k_gotoh_band
//   top_out    (bottom row of lane top_out_lane)  : + ext * (row + x)           consumers get FillArgs.top_in_row
//   lastrow_vals (last tile row, one word / column): + ext * (G*K - 1 + x)      un-drifted by k_lastrow_scan
// band_left (once per pair and sub-strip) and lastcol (once per pair and lane) hold plain values.
template <int G, int K, int NSUB>
__global__ void __maxnreg__(score2_maxnreg<K>()) k_gotoh_score2(const FillArgs a)
{
    static_assert(K % 4 == 0 && (32 % G) == 0 && (NSUB == 1 || (NSUB == 2 && K % 16 == 0)), "bad tile");
    constexpr int Kb = K / NSUB;                                          // rows per sub-strip of the band pass
    constexpr int PS = prof_stride(G, K);
    extern __shared__ __align__(128) int32_t sprof[];
    __shared__ __align__(8) uint64_t mbar;
    stage_profile(sprof, &mbar, a.prof, NPAIR * PS * 4);       // a.prof: the drifted table (S + 2 ext)

    const int lane = threadIdx.x & 31;
    const int t = lane % G;
    const int gl = lane / G;
    constexpr int GPW = 32 / G;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const bool lastLane = (t == G - 1);

    const uint32_t Z = BIAS2;                                             // stored 0
    const uint32_t NOPEN_ST = BIAS2 - (uint32_t)a.open * 0x10001u;        // stored -open
    const uint32_t e32 = a.d_e;                                           // ext in both halves
    const uint32_t cOpen = a.d_copen;                                     // per-half two's complement of ext - open
    const uint32_t cA_last = lastLane ? e32 : cOpen;                      // amplicon row La-1: zero end-gap penalties
    const uint32_t cB_last = lastLane ? e32 : 0u;
    // Band columns of this lane's upper sub-strip: xlo1 .. xlo1+W (xlo1 = first band column - 1); of its lower one
    // (NSUB == 2): shifted by Kb.  xlo1 is odd (host: P + B even) and W is odd, so the steps that touch the window
    // are those with x0 in [xlo1 - 1, xlo1 + W]: step s - sBand in [0, (W+1)/2]; both columns of such a step are
    // stored (band_topw has room for the two extra columns).  The registers at the left edge are saved in the first
    // of those steps (its second column is xlo1).
    const int xlo1 = a.band_row0 + t * K - a.band_B - 1;
    const int sBand = a.band_tops ? t + ((xlo1 - 1) >> 1) : (1 << 29);    // (no band: never)
    const unsigned nBand = (unsigned)((a.band_W + 1) >> 1);
    const uint32_t sprof_t = smem_u32(sprof) + (uint32_t)(t * strip_stride(K) * 4);
    const bool isTout = a.top_out && t == a.top_out_lane;
    const bool isRow = a.lastrow_vals && lastLane;
    // top boundary of lane 0: the free boundary (plain constants + this lane's drift) or the saved row of the pass that
    // owns the rows above (drifted in that pass's frame: row a.top_in_row; here it is row -1)
    const uint32_t tinAdj = e32 * (uint32_t)(-1 - a.top_in_row);

    for (int base = a.p0 + warp_global * GPW; base < a.p1; base += nwarps * GPW) {
        const int p = base + gl;
        const bool valid = p < a.p1;
        const int Lb = valid ? a.plen[p] : 0;
        const int nst = (Lb + 1) >> 1;                                    // column pairs of this lane's read
        const int steps = __reduce_max_sync(0xffffffffu, nst) + G - 1;
        const int64_t pco = valid ? a.pc_off[p] : 0;
        const int64_t pco_rel = valid ? pco - a.pc_off[a.p0] : 0;
        const uint8_t *pcp = a.pc + pco;
        const int firstRealSlot = (G * K - a.La) - t * K;                // slots below it are padding rows

        Strip<K> st;
#pragma unroll
        for (int k = 0; k < K; ++k) {                                    // boundary column x = -1: max3 = 0, ix = -open
            const uint32_t d = e32 * (uint32_t)(t * K + k - 1);
            st.H3[k] = Z + d; st.IX[k] = NOPEN_ST + d;
        }
        st.mlast = Z + e32 * (uint32_t)(t * K + K - 2);
        uint32_t bot0[3] = {Z, NOPEN_ST, Z}, bot1[3] = {Z, NOPEN_ST, Z};
        uint32_t hd0 = Z + e32 * (uint32_t)(t * K - 2);                  // max3[row above, -1] = 0
        int cpn0 = 0, cpn1 = 0;
        if (t == 0 && Lb > 0) { cpn0 = pcp[0]; cpn1 = pcp[Lb > 1 ? 1 : 0]; }
        const int64_t tcol = top_base_col(pco_rel, p - a.p0);
        const uint4 *tin = (a.top_in && valid) ? reinterpret_cast<const uint4 *>(a.top_in) + tcol : nullptr;
        uint4 *tout = reinterpret_cast<uint4 *>(a.top_out) + tcol;                         // used by lane top_out_lane only
        uint32_t *rowv = a.lastrow_vals + tcol;                                            // used by the last lane only
        const bool useTin = tin && t == 0;
        uint4 tnA = make_uint4(Z, NOPEN_ST, Z, 0u), tnB = tnA;
        if (useTin && Lb > 0) tnA = load_top3(tin);
        if (useTin && Lb > 1) tnB = load_top3(tin + 1);
        // what this lane receives at its band columns (bandw[x]: column x of the upper sub-strip; the lower one's slot
        // for the same column is MIDOFF further), and its registers at the band's left edge
        const int64_t sub_id = ((int64_t)(p - a.p0) * G + t) * NSUB;     // this lane's upper sub-strip
        const int TOPW = band_topw(a.band_W);
        uint4 *bandw = reinterpret_cast<uint4 *>(a.band_tops) + sub_id * TOPW - (xlo1 - 1);
        const int MIDOFF = TOPW - Kb;
        uint32_t *leftp = a.band_left + sub_id * band_leftw(Kb);
        const int sB = valid ? sBand : (1 << 29);

        // One systolic step: lane t works on columns x0 = 2 (s - t) and x0 + 1.
        // MODE 0 = steady (every lane on interior columns of its read), 1 = tail (some lane on or past its last column pair:
        // per-lane flags), 2 = head (lanes t > s have not started yet, nobody is near the end: the steady arithmetic behind
        // an activity test)
        auto step = [&](auto mode_tag, const int s) {
            constexpr int MODE = decltype(mode_tag)::value;
            constexpr bool TAIL = MODE == 1;
            const int x0 = 2 * (s - t), x1 = x0 + 1;
            uint32_t r0H = __shfl_up_sync(0xffffffffu, bot0[0], 1, G);
            uint32_t r0Y = __shfl_up_sync(0xffffffffu, bot0[1], 1, G);
            uint32_t r0M = __shfl_up_sync(0xffffffffu, bot0[2], 1, G);
            uint32_t r1H = __shfl_up_sync(0xffffffffu, bot1[0], 1, G);
            uint32_t r1Y = __shfl_up_sync(0xffffffffu, bot1[1], 1, G);
            uint32_t r1M = __shfl_up_sync(0xffffffffu, bot1[2], 1, G);
            if (t == 0) {
                const uint32_t d0 = e32 * (uint32_t)(x0 - 1);             // drift of (row -1, column x0)
                const uint32_t adj0 = useTin ? tinAdj : d0, adj1 = useTin ? tinAdj : d0 + e32;
                r0H = tnA.x + adj0; r0Y = tnA.y + adj0; r0M = tnA.z + adj0;
                r1H = tnB.x + adj1; r1Y = tnB.y + adj1; r1M = tnB.z + adj1;
            }
            const bool active = MODE == 0 || (x0 >= 0 && (MODE == 2 || x0 < Lb));
            const bool do1 = !TAIL || x1 < Lb;
            const bool last0 = TAIL && active && x0 == Lb - 1;
            const bool last1 = TAIL && active && x1 == Lb - 1;
            // (a steady step prefetches one code past an odd-length read: never index the profile with it)
            const int cp0 = cpn0, cp1 = do1 ? cpn1 : cpn0;
            if (MODE != 0) {
                if (x0 + 2 >= 0 && x0 + 2 < Lb) {
                    cpn0 = pcp[x0 + 2];
                    cpn1 = pcp[x0 + 3 < Lb ? x0 + 3 : x0 + 2];
                    if (useTin) { tnA = load_top3(tin + x0 + 2); if (x0 + 3 < Lb) tnB = load_top3(tin + x0 + 3); }
                }
            } else {
                // (x0 + 3 may be Lb for an odd length: one byte / one column past the read, inside the allocations; never used)
                cpn0 = pcp[x0 + 2];
                cpn1 = pcp[x0 + 3];
                if (useTin) { tnA = load_top3(tin + x0 + 2); tnB = load_top3(tin + x0 + 3); }
            }
            if (active) {
                const bool inBand = (unsigned)(s - sB) <= nBand;
                if (inBand && t > 0) {
                    bandw[x0] = make_uint4(r0H, r0Y, r0M, r0H);
                    bandw[x1] = make_uint4(r1H, r1Y, r1M, r1H);
                }
                uint32_t mid0[3] = {Z, Z, Z}, mid1[3] = {Z, Z, Z};
                score_columns2<K, TAIL, NSUB>(st, sprof_t + (uint32_t)cp0 * (PS * 4), sprof_t + (uint32_t)cp1 * (PS * 4), r0H, r0Y, r0M,
                                              r1H, r1Y, r1M, hd0, cOpen, cA_last, cB_last, e32, lastLane, last0, last1, do1,
                                              bot0, bot1, mid0, mid1);
                hd0 = r1H;                                                // max3'[row above, x1] for the next step's x0
                if (NSUB == 2 && (unsigned)(s - sB - Kb / 2) <= nBand) {
                    bandw[x0 + MIDOFF] = make_uint4(mid0[0], mid0[1], mid0[2], mid0[0]);
                    bandw[x1 + MIDOFF] = make_uint4(mid1[0], mid1[1], mid1[2], mid1[0]);
                }
                if (isTout) {
                    tout[x0] = make_uint4(bot0[0], bot0[1], bot0[2], bot0[0]);
                    if (do1) tout[x1] = make_uint4(bot1[0], bot1[1], bot1[2], bot1[0]);
                }
                if (isRow) {
                    // (the start-cell scan along the last amplicon row is k_lastrow_scan's)
                    if (do1) *reinterpret_cast<uint2 *>(rowv + x0) = make_uint2(bot0[0], bot1[0]);
                    else rowv[x0] = bot0[0];
                }
                // registers after column xlo-1 of a sub-strip (always a second column): the band pass starts from them
                // (scalar stores on purpose: vector stores would make ptxas shuffle 2K registers into aligned quads)
                if (s == sB) {
                    uint32_t d = e32 * (uint32_t)(t * K - 1 + x1);
#pragma unroll
                    for (int k = 0; k < Kb; ++k) { d += e32; leftp[k] = st.H3[k] - d; leftp[Kb + k] = st.IX[k] - d; }
                    if (NSUB == 1) leftp[2 * Kb] = st.mlast - d;
                }
                if (NSUB == 2 && s == sB + Kb / 2) {
                    uint32_t *lp = leftp + band_leftw(Kb);
                    uint32_t d = e32 * (uint32_t)(t * K + Kb - 1 + x1);
#pragma unroll
                    for (int k = 0; k < Kb; ++k) { d += e32; lp[k] = st.H3[Kb + k] - d; lp[Kb + k] = st.IX[Kb + k] - d; }
                    lp[2 * Kb] = st.mlast - d;
                }
            }
        };
        // Steps G-1 .. nst_min-2 are steady for the whole warp: every lane on interior columns of its read.
        const int nst_min = __reduce_min_sync(0xffffffffu, nst);
        const int steady_end = min(nst_min - 1, steps);                       // first non-steady step after the steady run
        int s = 0;
        if (steady_end >= G - 1)                                              // (no lane ends during the head)
            for (; s < G - 1; ++s) step(std::integral_constant<int, 2>{}, s);
        for (; s < min(G - 1, steps); ++s) step(std::integral_constant<int, 1>{}, s);
        for (; s < steady_end; ++s) step(std::integral_constant<int, 0>{}, s);
        for (; s < steps; ++s) step(std::integral_constant<int, 1>{}, s);

        // start-cell scan down the last read column (App. A.4) on plain values: first row whose max(m,ix,iy) is
        // strictly greater than everything above it; padded rows are not part of the matrix.  A lane is idle after
        // its last column, so its state arrays still hold it: all lanes of the warp scan together.
        if (valid) {
            uint32_t colBest = 0, d = e32 * (uint32_t)(t * K - 1 + Lb - 1);
            int colPosLo = 0, colPosHi = 0;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                d += e32;
                if (k >= firstRealSlot) {
                    const uint32_t nb = vmax2(colBest, st.H3[k] - d);
                    const uint32_t df = nb ^ colBest;
                    if (df & 0xffffu) colPosLo = k;
                    if (df >> 16) colPosHi = k;
                    colBest = nb;
                }
            }
            uint32_t *lcp = a.lastcol + ((int64_t)(p - a.p0) * G + t) * 3;   // (best, slot_lo, slot_hi) of this lane's rows, column Lb-1
            lcp[0] = colBest; lcp[1] = (uint32_t)colPosLo; lcp[2] = (uint32_t)colPosHi;
        }
    }
}

// Start-cell scan along the last amplicon row (App. A.4): first column whose max(m,ix,iy) is strictly greater than all
// columns before it, per read of a pair.  One warp per pair over the words k_gotoh_score2's last lane stored per
// column (drifted: + ext * (last_row + x)).  -> lastrow[pair] = (best, x_lo, x_hi), as the fill kernels write it.
__global__ void __launch_bounds__(256) k_lastrow_scan(const FillArgs a, const int last_row)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= a.p1 - a.p0) return;
    const int p = a.p0 + w;
    const int Lb = a.plen[p];
    const uint32_t *v = a.lastrow_vals + top_base_col(a.pc_off[p] - a.pc_off[a.p0], w);
    // per half: key = value << 16 | (0xffff - x): the maximum is the largest value at its smallest column
    uint32_t klo = 0, khi = 0;
    for (int x = lane; x < Lb; x += 32) {
        const uint32_t pv = v[x] - a.d_e * (uint32_t)(last_row + x);
        const uint32_t inv = 0xffffu - (uint32_t)x;
        klo = max(klo, ((pv & 0xffffu) << 16) | inv);
        khi = max(khi, (pv & 0xffff0000u) | inv);
    }
    klo = __reduce_max_sync(0xffffffffu, klo);
    khi = __reduce_max_sync(0xffffffffu, khi);
    if (lane == 0) {
        uint32_t *lrp = a.lastrow + (int64_t)w * 3;
        lrp[0] = (klo >> 16) | (khi & 0xffff0000u);
        lrp[1] = 0xffffu - (klo & 0xffffu);
        lrp[2] = 0xffffu - (khi & 0xffffu);
    }
}

template <int G, int K, int NSUB>
static cudaError_t launch_score2_tile(const FillArgs &a, int num_sms, cudaStream_t stream)
{
    const size_t smem = (size_t)NPAIR * prof_stride(G, K) * 4;
    // (per call: the attribute is per device, and a process may hold contexts on several devices)
    cudaError_t e = cudaFuncSetAttribute(k_gotoh_score2<G, K, NSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int blocks_per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_gotoh_score2<G, K, NSUB>, 128, smem);
    if (e != cudaSuccess) return e;
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    const int npairs = a.p1 - a.p0;
    const int groups_per_block = 4 * (32 / G);
    int grid = (npairs + groups_per_block - 1) / groups_per_block;
    const int cap = num_sms * blocks_per_sm;            // persistent: a multiple of the SM count
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    k_gotoh_score2<G, K, NSUB><<<grid, 128, smem, stream>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_lastrow_scan<<<(npairs + 7) / 8, 256, 0, stream>>>(a, G * K - 1);
    return cudaGetLastError();
}

// the band pass works on sub-strips of a.band_K rows: K (one per lane) or K/2 (two per lane, K % 16 == 0)
template <int G, int K>
static cudaError_t launch_score2_sub(const FillArgs &a, int num_sms, cudaStream_t stream)
{
    if (a.band_K == K) return launch_score2_tile<G, K, 1>(a, num_sms, stream);
    if constexpr (K % 16 == 0) { if (2 * a.band_K == K) return launch_score2_tile<G, K, 2>(a, num_sms, stream); }
    return cudaErrorInvalidValue;
}

template <int G, int K, int NSUB>
static int64_t wave_pairs2_tile(int num_sms)
{
    const size_t smem = (size_t)NPAIR * prof_stride(G, K) * 4;
    int bps = 0;
    if (cudaFuncSetAttribute(k_gotoh_score2<G, K, NSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_gotoh_score2<G, K, NSUB>, 128, smem) != cudaSuccess) return 0;
    return (int64_t)num_sms * std::max(bps, 1) * 4 * (32 / G);
}

int64_t score2_wave_pairs(int G, int K, int nsub, int num_sms)
{
#define CASE(g, k) if (G == g && K == k) { if (nsub == 1) return wave_pairs2_tile<g, k, 1>(num_sms); if constexpr (k % 16 == 0) { if (nsub == 2) return wave_pairs2_tile<g, k, 2>(num_sms); } return 0; }
    CASE(4, 16) CASE(4, 24) CASE(4, 32) CASE(4, 40) CASE(8, 16) CASE(8, 24) CASE(8, 32) CASE(8, 40)
    CASE(16, 16) CASE(16, 24) CASE(16, 32) CASE(16, 40) CASE(32, 24) CASE(32, 32) CASE(4, 48) CASE(8, 48) CASE(16, 48)
#undef CASE
    return 0;
}

cudaError_t launch_score2(int G, int K, const FillArgs &a, int num_sms, cudaStream_t stream)
{
#define CASE(g, k) if (G == g && K == k) return launch_score2_sub<g, k>(a, num_sms, stream);
    CASE(4, 16) CASE(4, 24) CASE(4, 32) CASE(4, 40) CASE(8, 16) CASE(8, 24) CASE(8, 32) CASE(8, 40)
    CASE(16, 16) CASE(16, 24) CASE(16, 32) CASE(16, 40) CASE(32, 24) CASE(32, 32) CASE(4, 48) CASE(8, 48) CASE(16, 48)
#undef CASE
    return cudaErrorInvalidValue;
}

}  // namespace crgpu
