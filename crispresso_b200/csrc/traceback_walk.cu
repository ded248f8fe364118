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

// One warp per pair: pc[pc_off[p]+x] = code(lo[x]) + 5*code(hi[x]).  A base outside ACGTN(U) -- an IUPAC ambiguity code,
// which needle would score with EDNAFULL's ambiguity rows -- is not aligned: bad[read] = 1 and the read is reported
// (kept bit 3 / an empty record, include/crgpu.h); it is encoded as A so that the pass over its pair stays defined.
// Without `bad` the call fails with CRGPU_E_ALIGN (*err) -- no silent substitution either way.
__global__ void k_encode_pairs(const uint8_t *__restrict__ reads, const int64_t *__restrict__ offsets,
                               const int32_t *__restrict__ pair_lo, const int32_t *__restrict__ pair_hi,
                               const int64_t *__restrict__ pc_off, int npairs, uint8_t *__restrict__ pc, int *err,
                               uint8_t *__restrict__ bad)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int p = warp; p < npairs; p += nwarps) {
        const int rl = pair_lo[p], rh = pair_hi[p];
        const uint8_t *lo = reads + offsets[rl];
        const uint8_t *hi = reads + offsets[rh];
        const int64_t o = pc_off[p];
        const int len = (int)(pc_off[p + 1] - o);
        for (int x = lane; x < len; x += 32) {
            int cl = base_code(lo[x]), ch = base_code(hi[x]);
            if (cl < 0 || ch < 0) {
                if (bad) { if (cl < 0) bad[rl] = 1; if (ch < 0) bad[rh] = 1; }
                else atomicOr(err, 1);
                cl = cl < 0 ? 0 : cl; ch = ch < 0 ? 0 : ch;
            }
            pc[o + x] = (uint8_t)(cl + NCODE * ch);
        }
    }
}

// crgpu_align: a read that was not aligned (bad[read]) comes back as an empty record (alnlen 0, identity 0, aln_off = slot)
__global__ void k_clear_bad_recs(crgpu_aln_rec *recs, const uint8_t *__restrict__ bad, int64_t n, int64_t slot)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !bad[i]) return;
    crgpu_aln_rec r = recs[i];
    r.score = 0.f; r.alnlen = 0; r.ident = 0; r.tenths = 0; r.aln_off = (int32_t)slot; r.start1 = -1; r.start2 = -1;
    recs[i] = r;
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

// base_code(x) == base_code(y) for bases of the accepted alphabet (A C G T U N, either case; reads with anything else are
// flagged by k_encode_pairs and their records cleared): upper case, U read as T
__device__ __forceinline__ uint32_t fold_base(uint32_t c)
{
    c &= 0xdfu;
    return c == 'U' ? (uint32_t)'T' : c;
}

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

// JOIN: 0 = plain walk; 1 = amplicon walk that records where it crosses WalkArgs.join_row; 2 = HDR walk (identity only:
// no text rows, no ops) that joins the amplicon walk's remainder there.  Separate instantiations keep each at 56 registers.
//
// Reads the diagonal shortcut already emitted (WalkArgs.fast) are skipped; with WalkArgs.read_list only the listed
// alignments of the batch are visited.
template <int JOIN>
__global__ void __launch_bounds__(128, 9) k_traceback_walk(const WalkArgs a)     // 9 CTAs/SM = at most 56 registers
{
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (a.read_list) {
        if (idx >= *a.read_list_n) return;
        idx = a.read_list[idx];
    }
    const int pj = idx >> 1;                             // pair of the batch (relative to p0)
    if (pj >= a.p1 - a.p0) return;
    const int p = a.p0 + pj;
    const int h = idx & 1;
    const int rlo = a.pair_lo[p], rhi = a.pair_hi[p];
    if (h && rhi == rlo) return;
    const int r = h ? rhi : rlo;
    if (a.fast && (a.fast[r] & a.fast_bit)) return;                // emitted by k_diag_emit
    const int La = a.La, Lb = a.plen[p];
    const uint8_t *b = a.reads + a.offsets[r];
    const uint8_t *amp = a.amplicon;
    const uint32_t *lr = a.lastrow + (int64_t)(p - a.p0) * 3;
    const uint32_t *lc = a.lastcol + (int64_t)(p - a.p0) * a.G * 3;

    // ---- start cell (App. A.4): the fill kernel already scanned the last amplicon row left to right
    //      and, per lane, the last read column top to bottom ((m, ix, iy) in that order, strict '>':
    //      the first maximum wins).  Combine: the column only wins with a strictly greater value.
    // With a shared prefix (tb_upper) the padded rows below split_row belong to the other pass: their
    // column summaries come from that pass's G_upper lanes; rows are numbered in the full padded frame.
    int best = half16(lr[0], h), s1 = La - 1, s2 = (int)lr[1 + h];
    if (a.tb_upper) {
        const uint32_t *lcu = a.lastcol_upper + (int64_t)(p - a.p0) * a.G_upper * 3;
        for (int t = 0; t < a.split_row / a.K; ++t) {
            const int v = half16(lcu[3 * t], h);
            if (v > best) { best = v; s1 = t * a.K + (int)lcu[3 * t + 1 + h] - a.P; s2 = Lb - 1; }
        }
    }
    for (int t = 0; t < a.G; ++t) {
        const int v = half16(lc[3 * t], h);
        if (v > best) { best = v; s1 = a.split_row + t * a.K + (int)lc[3 * t + 1 + h] - a.P; s2 = Lb - 1; }
    }

    const bool want = JOIN != 2 && a.ref_out != nullptr;
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
    // 2-bit op per column for k_quantify, in walk order (entry 0 = last column of the forward strand)
    uint32_t *opo = (JOIN != 2 && a.ops_out) ? a.ops_out + orow * a.ops_stride : nullptr;
    uint32_t opw = 0;
#define EMIT_OP(op) do { opw |= (uint32_t)(op) << ((n & 15) * 2); if ((n & 15) == 15) { if (opo) opo[n >> 4] = opw; opw = 0; } } while (0)
#define OUTC(c) (rc ? comp_upper(c) : (c))
#define EMIT_GAP_A(cb) do { if (want) { ro[dirn * n] = '-'; mo[dirn * n] = ' '; qo[dirn * n] = OUTC(cb); } EMIT_OP(2); ++n; } while (0)
#define EMIT_GAP_B(ca) do { if (want) { ro[dirn * n] = OUTC(ca); mo[dirn * n] = ' '; qo[dirn * n] = '-'; } EMIT_OP(3); ++n; } while (0)

    // ---- walk (App. A.5): trailing end gaps first (the strings are built right to left)
    for (int x = Lb - 1; x > s2; --x) EMIT_GAP_A(b[x]);
    for (int y = La - 1; y > s1; --y) EMIT_GAP_B(amp[y]);

    const bool band = a.band_W > 0;
    const int64_t pair_rel = p - a.p0;
    const uint8_t *tb = band ? reinterpret_cast<const uint8_t *>(a.tb) + pair_rel * a.G * a.band_W * a.K * 2
                             : reinterpret_cast<const uint8_t *>(a.tb + (a.pc_off[p] - a.pc_off[a.p0]) * (a.GK / 2));
    const int64_t colbytes = (int64_t)a.GK * 2;
    const uint8_t *tbu = !a.tb_upper ? nullptr
                         : band ? reinterpret_cast<const uint8_t *>(a.tb_upper) + pair_rel * a.G_upper * a.band_W * a.K * 2
                                : reinterpret_cast<const uint8_t *>(a.tb_upper + (a.pc_off[p] - a.pc_off[a.p0]) * (a.G_upper * a.K / 2));
    const int64_t colbytes_u = (int64_t)a.G_upper * a.K * 2;
    const int split = a.tb_upper ? a.split_row : 0;
    const int split_sub = band ? split / a.band_K : 0;
    const int P = a.P;
    // flag byte of cell (yy, xx).  Banded fill: sub-strip u = padded row / band_K holds columns [u*band_K - P - B, +W) only;
    // anything else reads as 0xff (never a flag byte), which ends the walk with an escape when it is consumed.
    // (HDR walk of a pair whose amplicon alignments were both emitted by the diagonal shortcut: the amplicon pass made flags
    //  only for the rows right above the split.  Such a walk meets the amplicon alignment's diagonal -- and ends at a join
    //  checkpoint -- within a few rows; otherwise the escape marker sends the read to the full fill.)
    const int upper_lo = (a.upper_need && !(a.upper_need[pair_rel] & 1)) ? a.upper_lo : 0;
    auto tb_at = [&](int yy, int xx) -> uint8_t {
        const int v = yy + P;
        if (v < upper_lo) return (uint8_t)0xff;
        if (band) {
            const int tf = (int)__umulhi((unsigned)v, a.kdiv_magic);           // sub-strip
            const int rr = v - tf * a.band_K;
            const int xr = xx - (tf * a.band_K - P - a.band_B);
            if ((unsigned)xr >= (unsigned)a.band_W) return (uint8_t)0xff;
            const uint8_t *q = v < split ? tbu + ((int64_t)tf * a.band_W + xr) * (a.band_K * 2)
                                         : tb + ((int64_t)(tf - split_sub) * a.band_W + xr) * (a.band_K * 2);
            return q[((rr >> 1) << 2) + (h << 1) + (rr & 1)];
        }
        return v < split ? tbu[xx * colbytes_u + ((v >> 1) << 2) + (h << 1) + (v & 1)]
                         : tb[xx * colbytes + (((v - split) >> 1) << 2) + (h << 1) + ((v - split) & 1)];
    };
    int y = s1, x = s2, prev = 0;
    bool contL = false, contD = false;
    const int ca0 = base_code(amp[0]);
    // The walk is a chain of dependent one-byte loads.  Paths are diagonal almost everywhere, so the
    // flags of the next PF cells down the diagonal are loaded ahead of time (pf[k] = cell (y-k, x-k));
    // a LEFT / DOWN step invalidates the window and refills it with PF independent loads.
    constexpr int PF = 8;
    uint8_t pf[PF];
#define PF_LOAD(k) pf[k] = (y - (k) >= 0 && x - (k) >= 0) ? tb_at(y - (k), x - (k)) : (uint8_t)0
#pragma unroll
    for (int k = 0; k < PF; ++k) PF_LOAD(k);
    // Join with the amplicon walk of the same read above the shared-prefix row (WalkArgs.join_row).  The amplicon walk
    // (JOIN 1) leaves (x, state, n, ident) at JOIN_NCK checkpoint rows, JOIN_CK rows apart, and its totals at the end; the
    // HDR walk (JOIN 2) compares at each checkpoint until it is at the same cell in the same state, then adds the
    // amplicon walk's remainder.  (Several checkpoints: a read edited at the cut site joins a few rows above the edit, and a
    // warp is only as fast as its last thread.)  Records are indexed by the thread's slot in the batch and zeroed per batch.
    bool shortcut = false;
    int ck_y = a.join_row - 1 - P, ck_k = 0;
    int32_t *jrec = JOIN != 0 ? a.join_out + ((int64_t)(p - a.p0) * 2 + h) * JOIN_STRIDE : nullptr;
    while (x >= 0 && y >= 0) {
        if (JOIN != 0 && ck_k < JOIN_NCK && y <= ck_y) {
            if (y == ck_y) {
                const int st = 0x100 | prev | ((prev == 1 && contL) ? 4 : 0) | ((prev == 2 && contD) ? 8 : 0);
                int4 *e = reinterpret_cast<int4 *>(jrec + 4 + 4 * ck_k);
                if (JOIN == 1) *e = make_int4(x, st, n, ident);
                else {
                    const int4 v = *e;
                    if (v.x == x && v.y == st) {
                        const int2 tot = *reinterpret_cast<const int2 *>(jrec);
                        if (tot.x > 0) {                               // (0: the amplicon walk left the band)
                            n += tot.x - v.z; ident += tot.y - v.w;
                            shortcut = true;
                            break;
                        }
                    }
                }
            }
            ck_y -= JOIN_CK; ++ck_k;
        }
        const int f = pf[0];
        if (f == 0xff) {                               // the path left the band: this read is re-aligned with the full fill
            a.escaped[r] |= (uint8_t)a.escape_bit;
            return;
        }
        // (selects, no branches: as an if-chain the three-way decision of App. A.4 was a fifth of the step's instructions)
        const bool nm = (f & F_NM) != 0, nx = (f & F_NX) != 0, ny = (f & F_NY) != 0;
        const bool wasL = prev == 1, wasD = prev == 2;
        const int dir_m = (wasL & !nx) ? 1 : (wasD & !ny) ? 2 : 0;       // m is a maximum: stay in the gap state that also is
        const int dir_g = nx ? 2 : 1;
        int dir = nm ? dir_g : dir_m;
        dir = (wasD & contD) ? 2 : dir;
        dir = (wasL & contL) ? 1 : dir;
        if (dir == 0) {
            const uint8_t ca = amp[y], cb = b[x];
            const bool same = fold_base(ca) == fold_base(cb);
            ident += same;
            if (want) { ro[dirn * n] = OUTC(ca); mo[dirn * n] = same ? '|' : '.'; qo[dirn * n] = OUTC(cb); }
            EMIT_OP(same ? 0 : 1);
            ++n; --x; --y;
#pragma unroll
            for (int k = 0; k < PF - 1; ++k) pf[k] = pf[k + 1];
            PF_LOAD(PF - 1);
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
#pragma unroll
            for (int k = 0; k < PF; ++k) PF_LOAD(k);
        } else {
            contD = !(f & F_NFY);
            EMIT_GAP_B(amp[y]);
            --y;
#pragma unroll
            for (int k = 0; k < PF; ++k) PF_LOAD(k);
        }
        prev = dir;
    }
#undef PF_LOAD
    if (!shortcut) {
        for (; x >= 0; --x) EMIT_GAP_A(b[x]);
        for (; y >= 0; --y) EMIT_GAP_B(amp[y]);
    }
    if (JOIN == 1) *reinterpret_cast<int2 *>(jrec) = make_int2(n, ident);
#undef EMIT_GAP_A
#undef EMIT_GAP_B
#undef OUTC
#undef EMIT_OP
    if (opo && (n & 15)) opo[n >> 4] = opw;

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


// ---- pairing plan, built on the device (crgpu_api.cu: build_plan) ---------------------------
// Reads are bucketed by length; consecutive reads of a bucket form the lo/hi halves of a pair.
// Only the 2049-bin length histogram visits the host.
__global__ void k_len_hist(const int64_t *__restrict__ offsets, const int32_t *__restrict__ subset, int64_t n,
                           int min_len, int max_len, int *__restrict__ hist, int *err)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < n;
    int len = -1;
    if (in) {
        const int64_t r = subset ? subset[i] : i;
        const int64_t l = offsets[r + 1] - offsets[r];
        if (l < min_len || l > max_len) atomicOr(err, 2);
        else len = (int)l;
    }
    // warp-aggregated increment: one atomic per distinct length per warp
    const unsigned peers = __match_any_sync(0xffffffffu, len);
    if (len >= 0 && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(hist + len, __popc(peers));
}

__global__ void k_scatter_order(const int64_t *__restrict__ offsets, const int32_t *__restrict__ subset, int64_t n,
                                const int64_t *__restrict__ read_start /* [max_len+1] */, int *__restrict__ cursor,
                                int32_t *__restrict__ order)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < n;
    int len = -1;
    int64_t r = 0;
    if (in) { r = subset ? subset[i] : i; len = (int)(offsets[r + 1] - offsets[r]); }
    const unsigned peers = __match_any_sync(0xffffffffu, len);
    const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    int base = 0;
    if (in && lane == leader) base = atomicAdd(cursor + len, __popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (in) order[read_start[len] + base + __popc(peers & ((1u << lane) - 1))] = (int32_t)r;
}

struct LenSeg { int len; int cnt; int64_t read_start; int64_t pair_start; int64_t pc_start; };

__global__ void k_build_pairs(const LenSeg *__restrict__ segs, int nseg, int np, const int32_t *__restrict__ order,
                              int32_t *__restrict__ pair_lo, int32_t *__restrict__ pair_hi, int32_t *__restrict__ plen,
                              int64_t *__restrict__ pc_off, int64_t total_pc)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p == np) pc_off[np] = total_pc;
    if (p >= np) return;
    int lo = 0, hi = nseg - 1;
    while (lo < hi) {                       // last segment with pair_start <= p
        const int mid = (lo + hi + 1) >> 1;
        if (segs[mid].pair_start <= p) lo = mid; else hi = mid - 1;
    }
    const LenSeg sg = segs[lo];
    const int64_t j = p - sg.pair_start;
    const int32_t a = order[sg.read_start + 2 * j];
    const int32_t b = (2 * j + 1 < sg.cnt) ? order[sg.read_start + 2 * j + 1] : a;
    pair_lo[p] = a; pair_hi[p] = b; plen[p] = sg.len;
    pc_off[p] = sg.pc_start + j * sg.len;
}

cudaError_t launch_len_hist(const int64_t *offsets, const int32_t *subset, int64_t n, int min_len, int max_len, int *hist,
                            int *err, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    k_len_hist<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(offsets, subset, n, min_len, max_len, hist, err);
    return cudaGetLastError();
}

cudaError_t launch_scatter_order(const int64_t *offsets, const int32_t *subset, int64_t n, const int64_t *read_start,
                                 int *cursor, int32_t *order, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    k_scatter_order<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(offsets, subset, n, read_start, cursor, order);
    return cudaGetLastError();
}

cudaError_t launch_build_pairs(const void *segs, int nseg, int np, const int32_t *order, int32_t *pair_lo, int32_t *pair_hi,
                               int32_t *plen, int64_t *pc_off, int64_t total_pc, cudaStream_t s)
{
    k_build_pairs<<<(unsigned)((np + 1 + 255) / 256), 256, 0, s>>>(reinterpret_cast<const LenSeg *>(segs), nseg, np, order,
                                                                   pair_lo, pair_hi, plen, pc_off, total_pc);
    return cudaGetLastError();
}

// ---- launch wrappers (called from crgpu_api.cu) -------------------------------------------
cudaError_t launch_encode(const uint8_t *reads, const int64_t *offsets, const int32_t *pair_lo, const int32_t *pair_hi,
                          const int64_t *pc_off, int npairs, uint8_t *pc, int *err, uint8_t *bad, int num_sms, cudaStream_t s)
{
    int grid = (npairs + 3) / 4;
    if (grid > num_sms * 16) grid = num_sms * 16;
    if (grid < 1) grid = 1;
    k_encode_pairs<<<grid, 128, 0, s>>>(reads, offsets, pair_lo, pair_hi, pc_off, npairs, pc, err, bad);
    return cudaGetLastError();
}

cudaError_t launch_clear_bad_recs(crgpu_aln_rec *recs, const uint8_t *bad, int64_t n, int64_t slot, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    k_clear_bad_recs<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(recs, bad, n, slot);
    return cudaGetLastError();
}

// Diagonal shortcut (DESIGN.md "Diagonal shortcut"), one WARP per read.  Claim: if the score of needle's start cell
// c_n equals D_n, the sum of the substitution scores along the diagonal c_0 (on row 0 or column 0) .. c_n, the
// traceback is that diagonal.  Proof: m[c_0] = S_0 (end gaps are free, App. A.2) and m[c_i] = S_i + max3[c_{i-1}] >=
// S_i + m[c_{i-1}] (A.3), so m[c_i] >= D_i; max3[c_n] = D_n then forces max3[c_i] = m[c_i] = D_i all the way down (a
// larger max3 anywhere would propagate into m[c_n] > D_n).  The walk starts with prev = 0 and takes the diagonal
// whenever m >= ix and m >= iy (A.4), which keeps prev = 0.
// One pass over the candidate alignment, 32 columns at a time, in the order k_traceback_walk emits them (App. A.5:
// trailing end gaps, the diagonal, leading end gaps): it sums the diagonal's scores and writes ops / text rows as it goes;
// when the sum confirms the claim the record, fast[read] |= fast_bit and (JOIN = 1) the walk's join records follow --
// state 0x100 = previous move diagonal at every checkpoint row on the path.  Otherwise need[pair - p0] = 1 and
// need_read[2 (pair - p0) + half] = 1: k_traceback_walk rewrites everything this kernel wrote for the read.
template <int JOIN>
__global__ void __launch_bounds__(256) k_diag_emit(const WalkArgs a)
{
    __shared__ int8_t lut[256];                                    // byte -> base code (one entry per thread)
    lut[threadIdx.x] = (int8_t)base_code((uint8_t)threadIdx.x);
    __syncthreads();
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int pj = w >> 1, h = w & 1;
    if (pj >= a.p1 - a.p0) return;
    const int p = a.p0 + pj;
    const int rlo = a.pair_lo[p], rhi = a.pair_hi[p];
    if (h && rhi == rlo) return;
    const int r = h ? rhi : rlo;
    const int La = a.La, Lb = a.plen[p];
    const uint8_t *b = a.reads + a.offsets[r];
    const uint8_t *amp = a.amplicon;
    // start cell, as k_traceback_walk combines it: lane t looks at lane t's column summary, the warp keeps the first maximum
    int best = half16(a.lastrow[(int64_t)pj * 3], h), s1 = La - 1, s2 = (int)a.lastrow[(int64_t)pj * 3 + 1 + h];
    {
        const int nu = a.tb_upper ? a.split_row / a.K : 0;         // lanes of the pass that owns the rows above the split
        int v = -1, row = 0;
        if (lane < nu) {
            const uint32_t *q = a.lastcol_upper + ((int64_t)pj * a.G_upper + lane) * 3;
            v = half16(q[0], h); row = lane * a.K + (int)q[1 + h] - a.P;
        } else if (lane < nu + a.G) {
            const uint32_t *q = a.lastcol + ((int64_t)pj * a.G + (lane - nu)) * 3;
            v = half16(q[0], h); row = a.split_row + (lane - nu) * a.K + (int)q[1 + h] - a.P;
        }
        // strictly greater than the last row's best and than every earlier lane's: max value, then the lowest lane
        const unsigned key = v > best ? ((unsigned)v << 8) | (unsigned)(255 - lane) : 0u;
        const unsigned top = __reduce_max_sync(0xffffffffu, key);
        if (top) {
            const int src = 255 - (int)(top & 0xffu);
            best = (int)(top >> 8); s1 = __shfl_sync(0xffffffffu, row, src); s2 = Lb - 1;
        }
    }
    const int len = (s1 < s2 ? s1 : s2) + 1;                        // cells of the diagonal through the start cell

    const bool want = a.ref_out != nullptr;
    const int64_t slot = a.slot;
    const int64_t orow = a.out_index ? a.out_index[r] : r;
    const bool rc = a.rc_out != 0;
    const int dirn = rc ? 1 : -1;
    const int64_t first = rc ? 0 : slot - 1;
    uint8_t *ro = want ? a.ref_out + orow * slot + first : nullptr;
    uint8_t *mo = want ? a.mark_out + orow * slot + first : nullptr;
    uint8_t *qo = want ? a.qry_out + orow * slot + first : nullptr;
    uint32_t *opo = a.ops_out ? a.ops_out + orow * a.ops_stride : nullptr;
    // columns in walk order (0 = the alignment's right end): T1 read bases opposite gaps, T2 amplicon bases opposite
    // gaps (one of the two is empty: the start cell is on the last row or the last column), the diagonal, then the
    // leftover read bases, then the leftover amplicon bases (again one of them is empty)
    const int T1 = Lb - 1 - s2, T2 = La - 1 - s1, L1 = s2 + 1 - len;
    const int d0 = T1 + T2, d1 = d0 + len, e1 = d1 + L1, n = e1 + (s1 + 1 - len);
    // join checkpoints: rows ck_y0, ck_y0 - CK, ...; checkpoint k sits on diagonal cell j = jk0 + k CK (column d0 + j)
    const int jk0 = s1 - (a.join_row - 1 - a.P);
    int ck_ident = 0;                                               // lane k: identities before checkpoint k's cell
    int ident = 0, sum = 0;
    for (int c0 = 0; c0 < n; c0 += 32) {
        const int c = c0 + lane;
        int op = 0;
        uint8_t ca = '-', cb = '-';
        bool same = false;
        if (c < n) {
            if (c < T1) { op = 2; cb = b[Lb - 1 - c]; }
            else if (c < d0) { op = 3; ca = amp[La - 1 - (c - T1)]; }
            else if (c < d1) {
                ca = amp[s1 - (c - d0)]; cb = b[s2 - (c - d0)];
                const int xa = lut[ca], xb = lut[cb];
                same = xa == xb;
                op = same ? 0 : 1;
                sum += (xa == 4 || xb == 4) ? (same ? -1 : -2) : (same ? 5 : -4);
            }
            else if (c < e1) { op = 2; cb = b[s2 - len - (c - d1)]; }
            else { op = 3; ca = amp[s1 - len - (c - e1)]; }
            if (want) {
                ro[dirn * c] = (ca == '-' || !rc) ? ca : comp_upper(ca);
                mo[dirn * c] = op >= 2 ? ' ' : (same ? '|' : '.');
                qo[dirn * c] = (cb == '-' || !rc) ? cb : comp_upper(cb);
            }
        }
        const unsigned eq = __ballot_sync(0xffffffffu, same);
        if (JOIN == 1 && lane < JOIN_NCK) {
            const int cc = d0 + jk0 + lane * JOIN_CK;               // column of checkpoint `lane`
            if (cc >= c0 && cc < c0 + 32) ck_ident = ident + __popc(eq & ((1u << (cc - c0)) - 1u));
        }
        ident += __popc(eq);
        const unsigned word = __reduce_or_sync(lane < 16 ? 0x0000ffffu : 0xffff0000u, (unsigned)op << ((lane & 15) * 2));
        if (opo && (lane & 15) == 0 && c < n) opo[c >> 4] = word;
    }
    sum = __reduce_add_sync(0xffffffffu, sum);
    if (sum * a.scale + BIAS != best) {
        if (lane == 0) { a.need[pj] = 1; if (a.need_read) a.need_read[2 * pj + h] = 1; }      // (need: set to 2 by the caller beforehand)
        return;
    }
    if (JOIN == 1) {
        int32_t *jrec = a.join_out + ((int64_t)pj * 2 + h) * JOIN_STRIDE;
        if (lane < JOIN_NCK) {
            // the walk is about to consume the diagonal cell of the checkpoint row: (x, state, columns so far, identities so far)
            const int j = jk0 + lane * JOIN_CK, y = s1 - j, x = s2 - j;
            if (j >= 0 && y >= 0 && x >= 0) *reinterpret_cast<int4 *>(jrec + 4 + 4 * lane) = make_int4(x, 0x100, d0 + j, ck_ident);
        }
        if (lane == 0) *reinterpret_cast<int2 *>(jrec) = make_int2(n, ident);
    }
    if (lane != 0) return;
    a.fast[r] |= (uint8_t)a.fast_bit;
    if (a.need_read) a.need_read[2 * pj + h] = 2;                    // emitted: this read's HDR walk need not wait for an amplicon walk
    crgpu_aln_rec rec;
    rec.score = (float)(best - BIAS) / (float)a.scale;
    rec.alnlen = n;
    rec.ident = ident;
    const float fpct = __fdiv_rn(100.0f * (float)ident, (float)n);     // App. B.3, as in k_traceback_walk
    rec.tenths = __double2int_rn((double)fpct * 10.0);
    rec.aln_off = rc ? 0 : (int32_t)(slot - n);
    rec.start1 = s1;
    rec.start2 = s2;
    rec.read_len = Lb;
    reinterpret_cast<crgpu_aln_rec *>(a.recs)[orow] = rec;
}

cudaError_t launch_walk(const WalkArgs &a, cudaStream_t s)
{
    const int nthreads = 2 * (a.p1 - a.p0);
    if (nthreads <= 0) return cudaSuccess;
    const int grid = (nthreads + 127) / 128;
    if (a.join_row > 0 && a.join_out) k_traceback_walk<1><<<grid, 128, 0, s>>>(a);
    else if (a.join_row > 0 && a.join_in && !a.ref_out && !a.ops_out) {
        WalkArgs b = a;
        b.join_out = const_cast<int32_t *>(a.join_in);          // one pointer in the kernel: JOIN 1 writes it, JOIN 2 reads it
        k_traceback_walk<2><<<grid, 128, 0, s>>>(b);
    }
    else k_traceback_walk<0><<<grid, 128, 0, s>>>(a);
    return cudaGetLastError();
}

// probing kernel of the diagonal shortcut (a.need zeroed by the caller); an amplicon alignment it emits still leaves
// its join records
cudaError_t launch_diag_emit(const WalkArgs &a, cudaStream_t s)
{
    const int64_t nwarps = 2 * (int64_t)(a.p1 - a.p0);
    if (nwarps <= 0) return cudaSuccess;
    if (!a.fast || !a.need) return cudaErrorInvalidValue;
    const int grid = (int)((nwarps + 7) / 8);
    if (a.join_row > 0 && a.join_out) k_diag_emit<1><<<grid, 256, 0, s>>>(a);
    else k_diag_emit<0><<<grid, 256, 0, s>>>(a);
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
