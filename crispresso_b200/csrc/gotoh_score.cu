// gotoh_score.cu -- k_gotoh_score<G,K>: score pass of the banded two-pass fill (DESIGN.md "Band").
//
// Same DP as k_gotoh_fill (needle's embAlignPathCalcWithEndGapPenalties as CRISPResso runs it,
// CRISPResso/CRISPRessoCORE.py:1791-1806; SURVEY.md App. A.1-A.3, exact integer form A.6), same work
// decomposition (a group of G lanes sweeps the read columns of one packed read pair as a systolic
// pipeline, lane t owns K amplicon rows), but it writes NO traceback flags: it produces the start-cell
// summaries (last amplicon row / last read column) and saves what the band pass needs -- per lane, the
// values received from the lane above at the band's columns and the register state at the band's left
// edge.  With no flags to derive, the recurrences are evaluated in DRIFT coordinates:
//
//     v'[r,x] = v[r,x] + ext * (r + x)          (r = row of the padded tile, x = read column)
//
// A gap extension then leaves a value unchanged, so the two `- ext` subtractions per cell disappear:
//     m'  = (S + 2 ext) + max3'[r-1,x-1]                       (the profile table carries the + 2 ext)
//     ix' = max(max3'[r,x-1] + (ext - open), ix'[r,x-1])       one VIADDMNMX.S16x2
//     iy' = max(max3'[r-1,x] + (ext - open), iy'[r-1,x])       one VIADDMNMX.S16x2
//     max3' = VIMNMX3.S16x2(m', ix', iy')
// i.e. 4 instructions per packed cell pair instead of 6 (3 of them on the integer-ALU pipe instead of 5).
// Needle's zero end-gap penalties (amplicon row La-1, read column Lb-1) become `+ ext` terms on those
// cells.  Everything that leaves the kernel is converted back (one subtraction of the cell's drift), so
// memory holds plain values and the band pass / the walker are unaffected.  Exact integer arithmetic: the
// host checks that value + drift stays inside the 15-bit range (run_plan_band).
#include "gotoh_tile.cuh"
#include <algorithm>
#include <type_traits>

namespace crgpu {

// One read column for the K rows of this lane, drift coordinates.
//  EDGE = true: some lane of the warp is on its LAST read column (iy opens from m only, zero penalties:
//  SURVEY App. A.2/A.3) -- the whole warp runs this body with per-lane parameters, no divergence.  The
//  first column needs no special case here (only its FY flag differs, and there are no flags).
template <int K, bool EDGE, int NSUB>
__device__ __forceinline__ void score_column(Strip<K> &st, const int32_t *__restrict__ prow,
                                             uint32_t upH3, uint32_t upIY, uint32_t upM, uint32_t hd,
                                             const uint32_t cOpen, const uint32_t cA_last, const uint32_t cB_last,
                                             const uint32_t e32, const bool lastLane, const bool isLastCol,
                                             const int firstRealSlot, uint32_t drift0,
                                             uint32_t &colBest, int &colPosLo, int &colPosHi,
                                             uint32_t &botH3, uint32_t &botIY, uint32_t &botM,
                                             uint32_t &midH3, uint32_t &midIY, uint32_t &midM)
{
    // every m of the column first: m'[k] needs max3'[k-1, x-1], the value row k-1 is about to overwrite; with the
    // adds hoisted, each row's new max3 can be written over the old one in place
    uint32_t M[K];
#pragma unroll
    for (int j = 0; j < K / 4; ++j) {
        const int4 S = *reinterpret_cast<const int4 *>(prow + 4 * j);
        M[4 * j] = (j == 0 ? hd : st.H3[4 * j - 1]) + (uint32_t)S.x;
        M[4 * j + 1] = st.H3[4 * j] + (uint32_t)S.y;
        M[4 * j + 2] = st.H3[4 * j + 1] + (uint32_t)S.z;
        M[4 * j + 3] = st.H3[4 * j + 2] + (uint32_t)S.w;
    }
    const uint32_t cV = (EDGE && isLastCol) ? e32 : cOpen;  // iy: what is added to the opening source
    const uint32_t bV = (EDGE && isLastCol) ? e32 : 0u;     // iy: what is added to the extended gap
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t h0 = st.H3[k], ix0 = st.IX[k];
        const uint32_t m = M[k];
        uint32_t ix;
        if (k == K - 1) {                                    // the only slot that can be amplicon row La-1
            const uint32_t src = lastLane ? st.mlast : h0;
            ix = vaddmax2(src, cA_last, ix0 + cB_last);
        } else {
            ix = vaddmax2(h0, cOpen, ix0);
        }
        uint32_t iy;
        if (EDGE) {
            const uint32_t src = isLastCol ? upM : upH3;
            iy = vaddmax2(src, cV, upIY + bV);
        } else {
            iy = vaddmax2(upH3, cOpen, upIY);
        }
        const uint32_t h3 = __vimax3_s16x2(m, ix, iy);
        if (EDGE) {
            // start-cell scan down the last read column (App. A.4) on plain values: first row whose
            // max(m,ix,iy) is strictly greater than everything above it; padded rows are not part of the matrix
            if (isLastCol && k >= firstRealSlot) {
                const uint32_t nb = vmax2(colBest, h3 - drift0);
                const uint32_t d = nb ^ colBest;
                if (d & 0xffffu) colPosLo = k;
                if (d >> 16) colPosHi = k;
                colBest = nb;
            }
            drift0 += e32;
        }
        st.H3[k] = h3;
        st.IX[k] = ix;
        if (k == K - 1) st.mlast = m;
        upH3 = h3; upIY = iy; upM = m;
        if (NSUB == 2 && k == K / 2 - 1) { midH3 = h3; midIY = iy; midM = m; }   // top boundary of the lower sub-strip
    }
    botH3 = upH3; botIY = upIY; botM = upM;
}

constexpr int SCORE_UNROLL = 2;          // steady steps per loop iteration (x4 measured the same, x1 8 % slower)
// 152 registers x 128 threads x 3 CTAs leave 7168 registers per SM: exactly one k_traceback_walk CTA, so the walks of the
// previous batch run beside the score pass instead of waiting for its tail (153+ registers: whole step 0.5 % slower)
#ifndef SCORE_MAXNREG
#define SCORE_MAXNREG 152
#endif

template <int K> constexpr int score_maxnreg() { return K <= 32 ? SCORE_MAXNREG : fill_maxnreg<K>(); }

template <int G, int K, int NSUB>
__global__ void __maxnreg__(score_maxnreg<K>()) k_gotoh_score(const FillArgs a)
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
    const uint32_t eK = a.d_eK;                                           // ext * K
    const uint32_t cOpen = a.d_copen;                                     // per-half two's complement of ext - open
    const uint32_t cA_last = lastLane ? e32 : cOpen;                      // amplicon row La-1: zero end-gap penalties
    const uint32_t cB_last = lastLane ? e32 : 0u;
    // band columns of this lane's upper sub-strip: xlo1+1 .. xlo1+W; of its lower one (NSUB == 2): shifted by Kb
    const int xlo1 = a.band_row0 + t * K - a.band_B - 1;
    const uint32_t eKb = a.d_eKb;                                         // ext * Kb

    for (int base = a.p0 + warp_global * GPW; base < a.p1; base += nwarps * GPW) {
        const int p = base + gl;
        const bool valid = p < a.p1;
        const int Lb = valid ? a.plen[p] : 0;
        const int steps = __reduce_max_sync(0xffffffffu, Lb) + G - 1;
        const int64_t pco = valid ? a.pc_off[p] : 0;
        const int64_t pco_rel = valid ? pco - a.pc_off[a.p0] : 0;
        const uint8_t *pcp = a.pc + pco;
        uint32_t *lrp = a.lastrow + (int64_t)(p - a.p0) * 3;             // (best, x_lo, x_hi) of amplicon row La-1
        uint32_t *lcp = a.lastcol + ((int64_t)(p - a.p0) * G + t) * 3;   // (best, slot_lo, slot_hi) of this lane's rows, column Lb-1
        const int firstRealSlot = (G * K - a.La) - t * K;                // slots below it are padding rows

        // drift of the row above this lane's strip (tile row t*K - 1) at the lane's current column x = s - t;
        // at x = -1 it is the drift of the boundary column
        uint32_t dTop = e32 * (uint32_t)(t * K - 2 - t);                 // value for s = -1 ...
        Strip<K> st;
#pragma unroll
        for (int k = 0; k < K; ++k) {                                    // boundary column x = -1: max3 = 0, ix = -open
            const uint32_t d = e32 * (uint32_t)(t * K + k - 1);
            st.H3[k] = Z + d; st.IX[k] = NOPEN_ST + d;
        }
        st.mlast = Z + e32 * (uint32_t)(t * K + K - 2);
        uint32_t botH3 = Z, botIY = NOPEN_ST, botM = Z;
        uint32_t hd0 = Z + e32 * (uint32_t)(t * K - 2);                  // max3[row above, -1] = 0
        uint32_t rowBest = 0, colBest = 0;                               // stored scores are > 0: 0 is -infinity
        int rowPosLo = 0, rowPosHi = 0, colPosLo = 0, colPosHi = 0;
        int cp_next = (t == 0 && Lb > 0) ? pcp[0] : 0;
        // top boundary (plain values) from the pass that owns the rows above (shared DP prefix), one column ahead
        const int64_t tcol = top_base_col(pco_rel, p - a.p0);
        const uint4 *tin = (a.top_in && valid) ? reinterpret_cast<const uint4 *>(a.top_in) + tcol : nullptr;
        uint4 *tout = (a.top_out && valid && t == a.top_out_lane) ? reinterpret_cast<uint4 *>(a.top_out) + tcol : nullptr;
        // read TWO columns ahead (tnA = column x, tnB = column x + 1): one systolic step is shorter than the latency of
        // an HBM read, and a value that has to be copied or used one step after its load stalls the whole warp
        uint4 tnA = make_uint4(Z, NOPEN_ST, Z, 0u), tnB = tnA;
        if (tin && t == 0 && Lb > 0) tnA = tin[0];
        if (tin && t == 0 && Lb > 1) tnB = tin[1];
        // what this lane receives at its band columns, and its registers at the band's left edge
        uint4 *bandw = nullptr, *midw = nullptr;                         // indexed by column x
        uint32_t *leftp = nullptr;
        if (valid && a.band_tops) {
            const int64_t sub_id = ((int64_t)(p - a.p0) * G + t) * NSUB;   // this lane's upper sub-strip
            if (t > 0) bandw = reinterpret_cast<uint4 *>(a.band_tops) + sub_id * band_topw(a.band_W) - xlo1;
            if (NSUB == 2) midw = reinterpret_cast<uint4 *>(a.band_tops) + (sub_id + 1) * band_topw(a.band_W) - (xlo1 + Kb);
            leftp = a.band_left + sub_id * band_leftw(Kb);
        }
        uint32_t midH3 = Z, midIY = Z, midM = Z;

        // One systolic step.  STEADY = every lane of the warp is on an interior column of its read (no lane
        // idle, none on its last column): the votes, the activity branch and the edge body drop out.
        // WHICH: 0 / 1 = first / second step of an unrolled pair (the top-boundary registers alternate, no copies);
        // 2 = a single step (uses tnA, then rotates)
        auto step = [&](auto steady_tag, auto which_tag, const int s) {
            constexpr bool STEADY = decltype(steady_tag)::value;
            constexpr int WHICH = decltype(which_tag)::value;
            const int x = s - t;
            dTop += e32;                                              // drift of (row t*K - 1, column x)
            uint32_t rH3 = __shfl_up_sync(0xffffffffu, botH3, 1, G);
            uint32_t rIY = __shfl_up_sync(0xffffffffu, botIY, 1, G);
            uint32_t rM = __shfl_up_sync(0xffffffffu, botM, 1, G);
            {
                const uint4 &tn = WHICH == 1 ? tnB : tnA;
                if (t == 0) { rH3 = tn.x + dTop; rIY = tn.y + dTop; rM = tn.z + dTop; }   // free boundary above the padded top, or the saved row
            }
            if (WHICH == 2) tnA = tnB;
            if (tin && t == 0 && x + 2 >= 0 && x + 2 < Lb) {
                if (WHICH == 0) tnA = tin[x + 2]; else tnB = tin[x + 2];
            }
            const bool active = STEADY || ((x >= 0) && (x < Lb));
            const bool lastCol = !STEADY && active && x == Lb - 1;
            const bool edge = !STEADY && __any_sync(0xffffffffu, lastCol);               // warp-uniform
            const int cp = cp_next;
            if (STEADY || (x + 1 >= 0 && x + 1 < Lb)) cp_next = pcp[x + 1];
            if (active) {
                const int32_t *prow = sprof + cp * PS + t * strip_stride(K);
                if (bandw && (unsigned)(x - xlo1) <= (unsigned)a.band_W) bandw[x] = make_uint4(rH3 - dTop, rIY - dTop, rM - dTop, 0u);
                if (edge)
                    score_column<K, true, NSUB>(st, prow, rH3, rIY, rM, hd0, cOpen, cA_last, cB_last, e32, lastLane, lastCol,
                                                firstRealSlot, dTop + e32, colBest, colPosLo, colPosHi, botH3, botIY, botM,
                                                midH3, midIY, midM);
                else
                    score_column<K, false, NSUB>(st, prow, rH3, rIY, rM, hd0, cOpen, cA_last, cB_last, e32, lastLane, false,
                                                 firstRealSlot, 0u, colBest, colPosLo, colPosHi, botH3, botIY, botM,
                                                 midH3, midIY, midM);
                if (NSUB == 2) {
                    if (midw && (unsigned)(x - xlo1 - Kb) <= (unsigned)a.band_W) {
                        const uint32_t dMid = dTop + eKb;             // drift of (row t*K + Kb - 1, column x)
                        midw[x] = make_uint4(midH3 - dMid, midIY - dMid, midM - dMid, 0u);
                    }
                }
                hd0 = rH3;                                            // max3'[row above, x] for column x+1
                // start-cell scan along the last amplicon row (meaningful in the last lane only), on plain
                // values: first column whose max(m,ix,iy) is strictly greater than all columns before it
                const uint32_t dBot = dTop + eK;                      // drift of this lane's bottom row at column x
                {
                    const uint32_t nb = vmax2(rowBest, botH3 - dBot);
                    const uint32_t d = nb ^ rowBest;
                    if (d & 0xffffu) rowPosLo = x;
                    if (d >> 16) rowPosHi = x;
                    rowBest = nb;
                }
                if (tout) tout[x] = make_uint4(botH3 - dBot, botIY - dBot, botM - dBot, 0u);
                if (!STEADY && leftp) {
                    // registers after column xlo-1 of a sub-strip: the band pass starts from them
                    // (scalar stores on purpose: vector stores would make ptxas shuffle 2K registers into aligned
                    // quads on EVERY step, outside this once-per-pair branch)
                    if (x == xlo1) {
                        uint32_t d = dTop;
#pragma unroll
                        for (int k = 0; k < Kb; ++k) { d += e32; leftp[k] = st.H3[k] - d; leftp[Kb + k] = st.IX[k] - d; }
                        if (NSUB == 1) leftp[2 * Kb] = st.mlast - d;
                    }
                    if (NSUB == 2 && x == xlo1 + Kb) {
                        uint32_t *lp = leftp + band_leftw(Kb);
                        uint32_t d = dTop + eKb;
#pragma unroll
                        for (int k = 0; k < Kb; ++k) { d += e32; lp[k] = st.H3[Kb + k] - d; lp[Kb + k] = st.IX[Kb + k] - d; }
                        lp[2 * Kb] = st.mlast - d;
                    }
                }
                if (lastCol) {
                    lcp[0] = colBest; lcp[1] = (uint32_t)colPosLo; lcp[2] = (uint32_t)colPosHi;
                    if (lastLane) { lrp[0] = rowBest; lrp[1] = (uint32_t)rowPosLo; lrp[2] = (uint32_t)rowPosHi; }
                }
            }
        };
        // Steps G .. Lmin-2 are steady for the whole warp (lane t is on column s - t), except the G steps on which
        // some lane saves its registers for the band pass (x == xlo1).  Steady steps run SCORE_UNROLL at a time: ptxas
        // resolves the loop-carried state (2K + a dozen registers) with register copies on the back edge, and the
        // unrolled body pays them once per SCORE_UNROLL columns.
        const int Lmin = __reduce_min_sync(0xffffffffu, Lb);
        const int steady_end = min(Lmin - 1, steps);                          // first non-steady step after the steady run
        int s = 0;
        while (s < steps) {
            const int d = xlo1 - (s - t);                                     // steps until this lane's (first) save
            const bool saves = (d >= 0 && d < SCORE_UNROLL) || (NSUB == 2 && d + Kb >= 0 && d + Kb < SCORE_UNROLL);
            if (s >= G && s + SCORE_UNROLL <= steady_end && !__any_sync(0xffffffffu, saves)) {
                static_assert(SCORE_UNROLL == 2, "the top-boundary registers alternate between the two steps of a pair");
                step(std::true_type{}, std::integral_constant<int, 0>{}, s);
                step(std::true_type{}, std::integral_constant<int, 1>{}, s + 1);
                s += SCORE_UNROLL;
            } else {
                step(std::false_type{}, std::integral_constant<int, 2>{}, s);
                ++s;
            }
        }
    }
}

template <int G, int K, int NSUB>
static cudaError_t launch_score_tile(const FillArgs &a, int num_sms, cudaStream_t stream)
{
    const size_t smem = (size_t)NPAIR * prof_stride(G, K) * 4;
    static bool configured = false;
    static int blocks_per_sm = 1;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_gotoh_score<G, K, NSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_gotoh_score<G, K, NSUB>, 128, smem);
        if (e != cudaSuccess) return e;
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        configured = true;
    }
    const int npairs = a.p1 - a.p0;
    const int groups_per_block = 4 * (32 / G);
    int grid = (npairs + groups_per_block - 1) / groups_per_block;
    const int cap = num_sms * blocks_per_sm;            // persistent: a multiple of the SM count
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    k_gotoh_score<G, K, NSUB><<<grid, 128, smem, stream>>>(a);
    return cudaGetLastError();
}

// the band pass works on sub-strips of a.band_K rows: K (one per lane) or K/2 (two per lane, K % 16 == 0)
template <int G, int K>
static cudaError_t launch_score_sub(const FillArgs &a, int num_sms, cudaStream_t stream)
{
    if (a.band_K == K) return launch_score_tile<G, K, 1>(a, num_sms, stream);
    if constexpr (K % 16 == 0) { if (2 * a.band_K == K) return launch_score_tile<G, K, 2>(a, num_sms, stream); }
    return cudaErrorInvalidValue;
}

template <int G, int K, int NSUB>
static int64_t wave_pairs_tile(int num_sms)
{
    const size_t smem = (size_t)NPAIR * prof_stride(G, K) * 4;
    int bps = 0;
    if (cudaFuncSetAttribute(k_gotoh_score<G, K, NSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_gotoh_score<G, K, NSUB>, 128, smem) != cudaSuccess) return 0;
    return (int64_t)num_sms * std::max(bps, 1) * 4 * (32 / G);
}

int64_t score_wave_pairs(int G, int K, int nsub, int num_sms)
{
#define CASE(g, k) if (G == g && K == k) { if (nsub == 1) return wave_pairs_tile<g, k, 1>(num_sms); if constexpr (k % 16 == 0) { if (nsub == 2) return wave_pairs_tile<g, k, 2>(num_sms); } return 0; }
    CASE(4, 16) CASE(4, 24) CASE(4, 32) CASE(4, 40) CASE(8, 16) CASE(8, 24) CASE(8, 32) CASE(8, 40)
    CASE(16, 16) CASE(16, 24) CASE(16, 32) CASE(16, 40) CASE(32, 24) CASE(32, 32) CASE(4, 48) CASE(8, 48) CASE(16, 48)
#undef CASE
    return 0;
}

cudaError_t launch_score(int G, int K, const FillArgs &a, int num_sms, cudaStream_t stream)
{
#define CASE(g, k) if (G == g && K == k) return launch_score_sub<g, k>(a, num_sms, stream);
    CASE(4, 16) CASE(4, 24) CASE(4, 32) CASE(4, 40) CASE(8, 16) CASE(8, 24) CASE(8, 32) CASE(8, 40)
    CASE(16, 16) CASE(16, 24) CASE(16, 32) CASE(16, 40) CASE(32, 24) CASE(32, 32) CASE(4, 48) CASE(8, 48) CASE(16, 48)
#undef CASE
    return cudaErrorInvalidValue;
}

}  // namespace crgpu
