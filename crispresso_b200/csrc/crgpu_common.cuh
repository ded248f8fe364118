// crgpu_common.cuh -- shared device/host definitions for libcrgpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace crgpu {

// ---- score representation -------------------------------------------------------------
// All DP values are exact integers (needle's float32 scores times `scale`),
// stored BIASED so that every value is a positive 15-bit number:  stored = value + BIAS.
// Two alignments (two reads of equal length against the same amplicon) share each 32-bit
// register: read "lo" in bits 0..15, read "hi" in bits 16..31.  Because both halves are always
// in [0, 0x7fff], (a) signed and unsigned 16x2 min/max agree, and (b) plain 32-bit adds of
// (c_hi*65536 + c_lo) and 32-bit subtracts of ordered operands never carry between halves, so
// they can be issued on either the integer-ALU or the FMA (IMAD) pipe.
constexpr int BIAS = 0x4000;
constexpr uint32_t BIAS2 = 0x40004000u;
constexpr uint32_t ONE2 = 0x00010001u;
constexpr int MAX_ABS_SCORE = 0x3c00;   // |value| must stay below this (checked on the host)

// read codes: A C G T N ; amplicon codes likewise plus PAD (virtual rows above row 0)
constexpr int NCODE = 5;
constexpr int NPAIR = NCODE * NCODE;    // pair code = lo + 5*hi

// traceback flag byte per DP cell ("not equal" bits; the walker inverts them):
//   bit0 nM : m  != max(m,ix,iy)      bit1 nX : ix != max3      bit2 nY : iy != max3
//   bit3 nFX: ix != ix[y,x-1]-gex(y)  bit4 nFY: iy != iy[y-1,x]-gey(x)
// Together they are sufficient for needle's traceback state machine.
constexpr int F_NM = 1, F_NX = 2, F_NY = 4, F_NFX = 8, F_NFY = 16;

// profile layout (int32 words) for a (G,K) kernel: lane t's strip of K rows starts at
// t*strip_stride(K); the stride is padded so that stride/4 is odd, i.e. the 8 lanes of a quarter
// warp hit 8 distinct 16-byte bank groups with their LDS.128.
__host__ __device__ constexpr int strip_stride(int K) { return ((K / 4) % 2 == 0) ? K + 4 : K; }
__host__ __device__ constexpr int prof_stride(int G, int K) { return G * strip_stride(K); }

struct FillArgs {
    const int32_t *prof;      // [NPAIR][prof_stride]  c32 = s_hi*65536 + s_lo (scaled scores)
    const uint8_t *pc;        // pair codes, pc[pc_off[p] + x]
    const int64_t *pc_off;    // [npairs+1] prefix sum of pair lengths
    const int32_t *plen;      // [npairs]
    uint32_t *tb;             // traceback flags of this batch; pair p starts at word (pc_off[p]-pc_off[p0]) * G*K/2
    uint32_t *lastrow;        // [p-p0][3]        (best max3, first column lo, hi) of amplicon row La-1
    uint32_t *lastcol;        // [(p-p0)*G + t][3] (best max3, first slot lo, hi) of lane t's rows in read column Lb-1
    int La;                   // amplicon length (padding rows = G*K - La)
    int p0, p1;               // pair range of this batch
    int open, ext;            // gap open / extend, scaled (positive)
    int one;                  // always 1 (an opaque multiplier, see fma_add in gotoh_fill.cu)
    // Shared DP prefix between the amplicon pass and the HDR-amplicon pass (crgpu_api.cu run_plan_dual):
    // the amplicon pass saves what lane `top_out_lane` hands down ((max3, iy, m, 0) of its bottom row) for
    // every column; the HDR pass starts below that row and takes those values as its top boundary.
    uint32_t *top_out;        // [(pc_off[p]-pc_off[p0]) + x][4] or null
    int top_out_lane;
    const uint32_t *top_in;   // same layout, or null (free end-gap boundary)
    // banded fill: top_out / top_in hold the score pass's drift coordinates, + ext * (top_in_row + x), top_in_row = the
    // row of the saved values in the frame of the pass that wrote them
    int top_in_row = 0;
    uint32_t *lastrow_vals = nullptr;   // score pass: max3 of the last tile row per column (drifted), indexed like top_out by
                                        // top_base_col(); k_lastrow_scan turns it into `lastrow`
    // Banded two-pass fill (DESIGN.md "Band"): k_gotoh_score2 evaluates every cell without flags and saves the state
    // k_gotoh_band needs to re-evaluate -- with flags -- only a diagonal band.  The band pass works on SUB-STRIPS of
    // band_K rows (band_K = K, or K/2 when that is a multiple of 8): sub-strip u = t*(K/band_K) + hh of a pair covers
    // the band_W read columns x in [xlo, xlo + band_W), xlo = band_row0 + u*band_K - band_B.
    int band_B, band_W;       // band_W = band_K + 2*band_B + 1; band_B = 0: not banded
    int band_K;               // rows per sub-strip of the band pass
    // drift constants of the score pass, packed per half (host-computed so that the kernel takes them straight from the
    // constant bank instead of re-deriving them from `ext` under register pressure): ext, ext*K, ext*band_K, ext - open
    uint32_t d_e, d_eK, d_eKb, d_copen;
    int band_row0;            // amplicon row of lane 0's first slot (-P for a full tile, split - P for the HDR sub-tile)
    uint32_t *band_tops;      // [(p-p0)*G2 + u][band_topw()][4]: (max3, iy, m, -) of the row above sub-strip u at columns xlo-2 .. xlo+W,
                              // in the score pass's drift coordinates: + ext * (u*band_K - 1 + x)
    uint32_t *band_left;      // [(p-p0)*G2 + u][band_leftw(band_K)]: H3[band_K], IX[band_K], mlast of column xlo-1
    uint32_t *band_tb;        // [(p-p0)*G2 + u][band_W][band_K/2]: flag words of the band columns      (G2 = G*K/band_K)
    // Diagonal shortcut (DESIGN.md "Diagonal shortcut"): the band pass only visits the pairs of the batch that still need a
    // traceback -- pair_list[0 .. *pair_list_n) holds their indices relative to p0 (null: every pair of the batch)
    const int32_t *pair_list = nullptr;
    const int *pair_list_n = nullptr;
    // ... and, with sub_n > 0, only sub-strips sub_lo .. sub_lo + sub_n - 1 of those pairs (the rows right above the split
    // of a shared-prefix run, for pairs whose amplicon alignments both took the shortcut: the HDR walk needs a few flags
    // there before it meets the amplicon walk's path)
    int sub_lo = 0, sub_n = 0;
};

// columns xlo-2 .. xlo+W of a sub-strip (W odd): the band pass reads xlo-1 .. xlo+W-1; k_gotoh_score2 stores both
// columns of every step that touches them
__host__ __device__ constexpr int band_topw(int W) { return (W + 4) & ~1; }
__host__ __device__ constexpr int band_leftw(int K) { return 2 * K + 4; }           // words, 16-byte multiple

constexpr int JOIN_NCK = 8, JOIN_CK = 8, JOIN_STRIDE = 4 + 4 * JOIN_NCK;    // checkpoints of the walk join (below)

struct WalkArgs {
    const uint32_t *tb;
    // optional upper part: padded rows below `split_row` were computed by another pass with tile
    // (G_upper, K) -- the shared DP prefix of the HDR pass -- and are read from there
    const uint32_t *tb_upper;
    const uint32_t *lastcol_upper;
    int G_upper, split_row;
    // banded fill: tb / tb_upper are band_tb arrays ([pair][sub-strip][band_W][band_K/2 words]); a cell outside the
    // band of its sub-strip raises escaped[read] (the caller re-aligns those reads with the full fill)
    int band_B, band_W, band_K;
    uint32_t kdiv_magic;      // ceil(2^32 / band_K): padded row / band_K by multiply-high
    uint8_t *escaped;         // [n reads] |= escape_bit
    int escape_bit;
    const uint32_t *lastrow;
    const uint32_t *lastcol;
    const int64_t *pc_off;
    const int32_t *plen;
    const int32_t *pair_lo;   // read index of the lo half
    const int32_t *pair_hi;   // read index of the hi half (== pair_lo for an unpaired read)
    const uint8_t *reads;     // original read bytes (device)
    const int64_t *offsets;   // [n+1]
    const uint8_t *amplicon;  // La bytes (upper-cased by the host)
    int La, GK, P, G, K;      // rows, padded rows, pad rows on top (P = GK - La), tile
    int p0, p1;
    int open, ext, scale;
    // outputs, indexed by read
    void *recs;               // crgpu_aln_rec*
    uint8_t *ref_out, *mark_out, *qry_out;   // may be null
    int64_t slot;
    uint32_t *ops_out;        // optional: 2-bit op per column in walk order, row stride ops_stride words
    int64_t ops_stride;
    const int32_t *out_index; // optional: output row of read r (null: row r)
    // Shared DP prefix: above padded row `join_row` the HDR walk reads the amplicon pass's own flags, so once both walks
    // of a read are at the same cell in the same state there, the rest of the two walks is identical.  The amplicon walk
    // records (x, state, n, ident) at JOIN_NCK checkpoint rows and its totals (join_out, JOIN_STRIDE ints per thread of the
    // batch, zeroed by the host before the batch); the HDR walk (identity only) compares at the checkpoints and, on a match,
    // adds the amplicon walk's remainder instead of walking it.
    int join_row;             // 0: off
    int32_t *join_out;        // amplicon walk: [2*(p-p0)+h][JOIN_STRIDE] = n_total, ident_total, -, -, then NCK x (x, state, n, ident)
    const int32_t *join_in;   // HDR walk: the same buffer
    int rc_out;               // 1: the amplicon is a reverse complement; emit rows flipped back to the
                              // forward strand (CORE:1982-1990), left-aligned in the slot
    // Diagonal shortcut (DESIGN.md "Diagonal shortcut").  If the score of the start cell equals the sum of the
    // substitution scores along the diagonal through it, needle's traceback IS that diagonal (every cell on it has
    // m == max(m,ix,iy)), so the alignment is emitted without any flag by k_diag_emit (one warp per read): it sets
    // fast[read] |= fast_bit for those and need[pair - p0] = 1 (set to 2 by the caller beforehand) for pairs with a
    // read it could not finish.  k_traceback_walk skips reads with fast[read] & fast_bit and -- with read_list -- only visits the listed
    // alignments.
    uint8_t *fast = nullptr;  // [n reads] or null
    int fast_bit = 0;
    uint8_t *need = nullptr;  // [p1 - p0] (k_diag_emit only)
    uint8_t *need_read = nullptr;         // [2 (p1 - p0)] (k_diag_emit only, zeroed by the caller): 1 = this alignment needs the walk
    const int32_t *read_list = nullptr;   // k_traceback_walk: the alignments to walk, 2 (pair - p0) + half each, or null (all)
    const int *read_list_n = nullptr;
    // HDR walk over a shared prefix: !(upper_need[pair - p0] & 1) means both amplicon alignments of the pair took the
    // shortcut and the amplicon pass wrote flags only for padded rows >= upper_lo; null: flags exist for every pair and row
    const uint8_t *upper_need = nullptr;
    int upper_lo = 0;
};

// The 4 bytes at byte offset i of the byte string b[0, len) -- any alignment -- as one little-endian word: aligned 32-bit
// loads + a funnel shift; bytes past len read as 0 and are not touched in memory (beyond the aligned word b[len-1] lies in).
__device__ __forceinline__ uint32_t load4(const uint8_t *__restrict__ b, int i, int len)
{
    const uintptr_t p = reinterpret_cast<uintptr_t>(b + i);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(p & ~(uintptr_t)3);
    const int mis = (int)(p & 3);
    const int rem = len - i;                                   // > 0
    const uint32_t lo = w[0];
    uint32_t hi = 0;
    if (mis && rem > 4 - mis) hi = w[1];
    uint32_t v = __funnelshift_r(lo, hi, mis * 8);
    if (rem < 4) v &= (1u << (8 * rem)) - 1u;
    return v;
}

}  // namespace crgpu
