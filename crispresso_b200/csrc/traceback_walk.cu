// traceback_walk.cu -- k_encode_pairs, k_traceback_walk, k_qualfilter.
//
// k_traceback_walk reproduces needle's start-cell choice, its value-comparing traceback state
// machine and the alignment walk (embAlignPathCalcWithEndGapPenalties' tail and
// embAlignWalkNWMatrixUsingCompressedTraceback) from the 5 flag bits
// per cell written by k_gotoh_fill, and emits what parse_needle_output extracts from the
// srspair text (CRISPResso/CRISPRessoCORE.py:1707-1786): the three aligned rows, identity
// (ident/alnlen and the printed 1-decimal value, App. B.3), score and read length.
#include "crgpu_common.cuh"
#include "../../include/crgpu.h"

namespace crgpu {

__device__ __forceinline__ int base_code(uint8_t c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': case 'U': case 'u': return 3;
    case 'N': case 'n': return 4;
    default: return -1;
    }
}

__device__ __forceinline__ int ednafull(int ca, int cb)
{
    if (ca == 4 && cb == 4) return -1;
    if (ca == 4 || cb == 4) return -2;
    return ca == cb ? 5 : -4;
}

// One warp per pair: pc[pc_off[p]+x] = code(lo[x]) + 5*code(hi[x]); a base outside ACGTN(U)
// raises *err (the call then fails with CRGPU_E_ALIGN -- no silent substitution).
__global__ void k_encode_pairs(const uint8_t *__restrict__ reads, const int64_t *__restrict__ offsets,
                               const int32_t *__restrict__ pair_lo, const int32_t *__restrict__ pair_hi,
                               const int64_t *__restrict__ pc_off, int npairs, uint8_t *__restrict__ pc, int *err)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int p = warp; p < npairs; p += nwarps) {
        const uint8_t *lo = reads + offsets[pair_lo[p]];
        const uint8_t *hi = reads + offsets[pair_hi[p]];
        const int64_t o = pc_off[p];
        const int len = (int)(pc_off[p + 1] - o);
        for (int x = lane; x < len; x += 32) {
            const int cl = base_code(lo[x]), ch = base_code(hi[x]);
            if (cl < 0 || ch < 0) { atomicOr(err, 1); pc[o + x] = 0; }
            else pc[o + x] = (uint8_t)(cl + NCODE * ch);
        }
    }
}

// reverse_complement() of the reference upper-cases and complements (CORE:141-144)
__device__ __forceinline__ uint8_t comp_upper(uint8_t c)
{
    switch (c) {
    case 'A': case 'a': return 'T';
    case 'C': case 'c': return 'G';
    case 'G': case 'g': return 'C';
    case 'T': case 't': case 'U': case 'u': return 'A';
    case 'N': case 'n': return 'N';
    default: return c;
    }
}

__device__ __forceinline__ int half16(uint32_t w, int h) { return (int)((w >> (16 * h)) & 0xffffu); }

// ix[0, j] of needle's first row (App. A.2), true scaled value.
__device__ int row0_ix(const uint8_t *b, int ca0, int j, int open, int ext, int scale)
{
    int v = -open;
    for (int i = 1; i <= j; ++i) {
        const int m = scale * ednafull(ca0, base_code(b[i - 1]));
        const int o = m - open, e = v - ext;
        v = o >= e ? o : e;
    }
    return v;
}

__global__ void __launch_bounds__(128) k_traceback_walk(const WalkArgs a)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int npairs = a.p1 - a.p0;
    if (idx >= 2 * npairs) return;
    const int p = a.p0 + (idx >> 1);
    const int h = idx & 1;
    const int rlo = a.pair_lo[p], rhi = a.pair_hi[p];
    if (h && rhi == rlo) return;
    const int r = h ? rhi : rlo;
    const int La = a.La, Lb = a.plen[p];
    const uint8_t *b = a.reads + a.offsets[r];
    const uint8_t *amp = a.amplicon;
    const uint32_t *lr = a.lastrow + (a.pc_off[p] - a.pc_off[a.p0]) * 3;
    const uint32_t *lc = a.lastcol + ((int64_t)(p - a.p0) * a.GK + a.P) * 3;

    // ---- start cell (App. A.4): last row left to right, then last column top to bottom,
    //      (m, ix, iy) in that order, strict '>' so the first maximum wins.
    int best = -1, s1 = La - 1, s2 = Lb - 1;
    for (int x = 0; x < Lb; ++x) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int v = half16(lr[3 * x + k], h);
            if (v > best) { best = v; s2 = x; }
        }
    }
    for (int y = 0; y < La; ++y) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int v = half16(lc[3 * y + k], h);
            if (v > best) { best = v; s1 = y; s2 = Lb - 1; }
        }
    }

    const bool want = a.ref_out != nullptr;
    const int64_t slot = a.slot;
    const int64_t orow = a.out_index ? a.out_index[r] : r;
    // Columns are produced from the alignment's right end to its left end.  Forward strand: store
    // them right to left from the end of the slot.  rc_out: storing them left to right from the
    // start of the slot, complemented, IS the reverse complement the reference applies afterwards.
    const int dirn = a.rc_out ? 1 : -1;
    const int64_t first = a.rc_out ? 0 : slot - 1;
    uint8_t *ro = want ? a.ref_out + orow * slot + first : nullptr;
    uint8_t *mo = want ? a.mark_out + orow * slot + first : nullptr;
    uint8_t *qo = want ? a.qry_out + orow * slot + first : nullptr;
    const bool rc = a.rc_out != 0;
    int n = 0, ident = 0;
#define OUTC(c) (rc ? comp_upper(c) : (c))
#define EMIT_GAP_A(cb) do { if (want) { ro[dirn * n] = '-'; mo[dirn * n] = ' '; qo[dirn * n] = OUTC(cb); } ++n; } while (0)
#define EMIT_GAP_B(ca) do { if (want) { ro[dirn * n] = OUTC(ca); mo[dirn * n] = ' '; qo[dirn * n] = '-'; } ++n; } while (0)

    // ---- walk (App. A.5): trailing end gaps first (the strings are built right to left)
    for (int x = Lb - 1; x > s2; --x) EMIT_GAP_A(b[x]);
    for (int y = La - 1; y > s1; --y) EMIT_GAP_B(amp[y]);

    const uint8_t *tb = reinterpret_cast<const uint8_t *>(a.tb + a.tb_off[p]);
    const int64_t colbytes = (int64_t)a.GK * 2;
    int y = s1, x = s2, prev = 0;
    bool contL = false, contD = false;
    const int ca0 = base_code(amp[0]);
    while (x >= 0 && y >= 0) {
        const int rr = y + a.P;
        const int f = tb[x * colbytes + ((rr >> 1) << 2) + (h << 1) + (rr & 1)];
        int dir;
        if (prev == 1 && contL) dir = 1;
        else if (prev == 2 && contD) dir = 2;
        else if (!(f & F_NM)) {
            if (prev == 1 && !(f & F_NX)) dir = 1;
            else if (prev == 2 && !(f & F_NY)) dir = 2;
            else dir = 0;
        }
        else if (!(f & F_NX)) dir = 1;
        else dir = 2;
        if (dir == 0) {
            const uint8_t ca = amp[y], cb = b[x];
            const bool same = base_code(ca) == base_code(cb);
            ident += same;
            if (want) { ro[dirn * n] = OUTC(ca); mo[dirn * n] = same ? '|' : '.'; qo[dirn * n] = OUTC(cb); }
            ++n; --x; --y;
        } else if (dir == 1) {
            // the next cell (y, x-1) continues LEFT iff ix[y,x-1] - gex(y) == ix[y,x].  The fill
            // kernel evaluates that with gex = gapextend except on amplicon row La-1; needle uses
            // gex = 0 on row 0 too (App. A.4), so row 0 is re-evaluated here from its closed form.
            if (y == 0 && x >= 1)
                contL = row0_ix(b, ca0, x - 1, a.open, a.ext, a.scale) == row0_ix(b, ca0, x, a.open, a.ext, a.scale);
            else
                contL = !(f & F_NFX);
            EMIT_GAP_A(b[x]);
            --x;
        } else {
            contD = !(f & F_NFY);
            EMIT_GAP_B(amp[y]);
            --y;
        }
        prev = dir;
    }
    for (; x >= 0; --x) EMIT_GAP_A(b[x]);
    for (; y >= 0; --y) EMIT_GAP_B(amp[y]);
#undef EMIT_GAP_A
#undef EMIT_GAP_B
#undef OUTC

    crgpu_aln_rec rec;
    rec.score = (float)(best - BIAS) / (float)a.scale;
    rec.alnlen = n;
    rec.ident = ident;
    // App. B.3: "%4.1f" of (float)100 * ident / len, as tenths (round-half-even on the exact product)
    const float fpct = __fdiv_rn(100.0f * (float)ident, (float)n);
    rec.tenths = __double2int_rn((double)fpct * 10.0);
    rec.aln_off = rc ? 0 : (int32_t)(slot - n);
    rec.start1 = s1;
    rec.start2 = s2;
    rec.read_len = Lb;
    reinterpret_cast<crgpu_aln_rec *>(a.recs)[orow] = rec;
}

// S1: one warp per read; keep iff sum(phred) >= q*len and min(phred) >= s (CORE:186-190, 300-305).
__global__ void k_qualfilter(const uint8_t *__restrict__ qual, const int64_t *__restrict__ offsets, int64_t n,
                             int min_mean_q, int min_single_q, uint8_t *__restrict__ keep)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += nwarps) {
        const int64_t o = offsets[i];
        const int len = (int)(offsets[i + 1] - o);
        long long sum = 0;
        int mn = 1 << 30;
        for (int x = lane; x < len; x += 32) {
            const int q = (int)qual[o + x] - 33;
            sum += q;
            mn = q < mn ? q : mn;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, d);
            const int o2 = __shfl_xor_sync(0xffffffffu, mn, d);
            mn = o2 < mn ? o2 : mn;
        }
        if (lane == 0) keep[i] = (len > 0 && sum >= (long long)min_mean_q * len && mn >= min_single_q) ? 1 : 0;
    }
}

// ---- launch wrappers (called from crgpu_api.cu) -------------------------------------------
cudaError_t launch_encode(const uint8_t *reads, const int64_t *offsets, const int32_t *pair_lo, const int32_t *pair_hi,
                          const int64_t *pc_off, int npairs, uint8_t *pc, int *err, int num_sms, cudaStream_t s)
{
    int grid = (npairs + 3) / 4;
    if (grid > num_sms * 16) grid = num_sms * 16;
    if (grid < 1) grid = 1;
    k_encode_pairs<<<grid, 128, 0, s>>>(reads, offsets, pair_lo, pair_hi, pc_off, npairs, pc, err);
    return cudaGetLastError();
}

cudaError_t launch_walk(const WalkArgs &a, cudaStream_t s)
{
    const int nthreads = 2 * (a.p1 - a.p0);
    if (nthreads <= 0) return cudaSuccess;
    k_traceback_walk<<<(nthreads + 127) / 128, 128, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_qualfilter(const uint8_t *qual, const int64_t *offsets, int64_t n, int q, int sq, uint8_t *keep,
                              int num_sms, cudaStream_t s)
{
    int64_t grid = (n + 3) / 4;
    if (grid > (int64_t)num_sms * 16) grid = (int64_t)num_sms * 16;
    if (grid < 1) grid = 1;
    k_qualfilter<<<(int)grid, 128, 0, s>>>(qual, offsets, n, q, sq, keep);
    return cudaGetLastError();
}

}  // namespace crgpu
