// quantify.cu -- k_quantify: fused indel/substitution extraction, read classification and
// histogram reduction.  Replaces CRISPResso's per-read Python loop process_df_chunk
// (CRISPResso/CRISPRessoCORE.py:428-753) together with the per-row preparation that feeds it
// (ignore_n_in_alignment CORE:2040-2052, compute_ref_positions CORE:2055-2067).
//
// One thread per aligned read, fed with 2 bits per alignment column (match / mismatch / insertion /
// deletion) -- written by k_traceback_walk, or derived from the three text rows by k_rows_to_ops
// when the caller comes through crgpu_quantify.  The read's substitution / deletion / insertion-flank position
// sets are kept as amplicon-length bitmaps (<= 32 words each) in local memory, which makes the
// reference's set semantics -- INCLUDE_IDXS.intersection(...), numpy's buffered fancy-index
// `vec[list] += 1` that increments a duplicated index once, negative flank indices wrapping
// round the amplicon -- plain AND / OR / popcount.
// Per-position vectors, the frameshift histograms and the counters are reduced with 64-bit
// global atomics (REDG); per-read results go to crgpu_read_rec.
#include "../../include/crgpu.h"
#include "crgpu_common.cuh"

#include "quant_args.cuh"

namespace crgpu {

constexpr int MAXW = CRGPU_MAX_AMPLICON / 32;


// CORE:2059: only upper-case A,T,C,G,N advance the amplicon index (everything else in the amplicon
// row is a gap column); CORE:504/518: '-' runs are the indels; CORE:491: '.' is a substitution
__device__ __forceinline__ bool is_ref_base(uint8_t c) { return c == 'A' || c == 'T' || c == 'C' || c == 'G' || c == 'N'; }

// rows (as parse_needle_output delivers them) -> 2-bit ops, forward order
__global__ void k_rows_to_ops(const uint8_t *__restrict__ ref, const uint8_t *__restrict__ mark, const uint8_t *__restrict__ qry,
                              int64_t slot, const int32_t *__restrict__ aln_off, const int32_t *__restrict__ alnlen, int64_t n,
                              uint32_t *__restrict__ ops, int64_t ops_stride)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t base = i * slot + (aln_off ? aln_off[i] : 0);
    const int len = alnlen[i];
    uint32_t w = 0;
    for (int c = 0; c < len; ++c) {
        const uint32_t op = !is_ref_base(ref[base + c]) ? 2u : (qry[base + c] == '-' ? 3u : (mark[base + c] == '.' ? 1u : 0u));
        w |= op << ((c & 15) * 2);
        if ((c & 15) == 15) { ops[i * ops_stride + (c >> 4)] = w; w = 0; }
    }
    if (len & 15) ops[i * ops_stride + (len >> 4)] = w;
}

cudaError_t launch_rows_to_ops(const uint8_t *ref, const uint8_t *mark, const uint8_t *qry, int64_t slot, const int32_t *aln_off,
                               const int32_t *alnlen, int64_t n, uint32_t *ops, int64_t ops_stride, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    k_rows_to_ops<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(ref, mark, qry, slot, aln_off, alnlen, n, ops, ops_stride);
    return cudaGetLastError();
}

__device__ __forceinline__ void set_bit(uint32_t *bits, int v) { bits[v >> 5] |= 1u << (v & 31); }
__device__ __forceinline__ bool get_bit(const uint32_t *bits, int v) { return (bits[v >> 5] >> (v & 31)) & 1u; }

// The per-position vectors: rows pile their events onto the few dozen positions around the cut site, so global atomics
// would queue on those addresses (~30 M adds on ~60 addresses per 2^20 reads).  Every CTA keeps the 15 vectors in shared
// memory (32-bit, `sv`) and adds what it gathered to the global 64-bit vectors once, at its end.
struct VecAcc {
    unsigned *sv;
    int L;
    __device__ __forceinline__ void add(int vec, int pos, unsigned val) const { atomicAdd(sv + vec * L + pos, val); }
    __device__ __forceinline__ void add_bits(int vec, const uint32_t *bits, int W) const
    {
        unsigned *v = sv + vec * L;
        for (int w = 0; w < W; ++w) {
            uint32_t x = bits[w];
            while (x) {
                const int b = __ffs(x) - 1;
                x &= x - 1;
                atomicAdd(v + w * 32 + b, 1u);
            }
        }
    }
};

struct OpReader {
    const uint32_t *w;
    int n;
    bool rev;
    __device__ __forceinline__ int at(int c) const
    {
        const int j = rev ? n - 1 - c : c;
        return (int)((__ldg(w + (j >> 4)) >> ((j & 15) * 2)) & 3u);
    }
};
constexpr int OP_MATCH = 0, OP_MISMATCH = 1, OP_INS = 2, OP_DEL = 3;

// Columns from c on that are matches, up to the end of the op word c lies in (0: column c is not a match).  Most of a row
// is matches: the passes below step over them a word (16 columns) at a time.
__device__ __forceinline__ int match_run(const OpReader &ops, int c)
{
    if (ops.rev) {
        const int j = ops.n - 1 - c, k = j & 15;
        const uint32_t word = __ldg(ops.w + (j >> 4));
        const uint32_t m = k == 15 ? word : (word & ((1u << (2 * k + 2)) - 1u));
        if (m == 0) return k + 1;
        return (31 - __clz(m)) / 2 < k ? k - (31 - __clz(m)) / 2 : 0;          // matches above the highest non-match op
    }
    const int k = c & 15;
    const int avail = min(16 - k, ops.n - c);
    uint32_t m = __ldg(ops.w + (c >> 4)) >> (2 * k);
    if (avail < 16) m &= (1u << (2 * avail)) - 1u;
    if (m == 0) return avail;
    return (__ffs(m) - 1) / 2;                                                  // matches below the lowest non-match op
}

// One kept row.  `sc`: the CTA's shared-memory copy of the nine scalar counters (4 frameshift counters, 4 class counts,
// rows seen) -- every row bumps two or three of them, so they are reduced per warp (ptxas aggregates the uniform-address
// shared atomics) and per CTA before one global atomic each; the per-position vectors and the histograms get few,
// scattered hits per row and go straight to global memory.
constexpr int NSCALAR = CRGPU_NUM_COUNTERS + 5;
// The light part: a row that came in UNMODIFIED (identity 100, CORE:2014) only bumps two counters.  Returns true when the
// row needs quantify_row.  (With an N in the amplicon every row takes the long way: its markup is re-read first.)
__device__ __forceinline__ bool quantify_row_light(const QuantArgs &a, const int64_t i, unsigned *sc)
{
    if (!a.unmod_in[i]) return true;
    crgpu_read_rec rec;
    rec.cls = CRGPU_C_UNMODIFIED; rec.pad[0] = rec.pad[1] = rec.pad[2] = 0;
    rec.n_mutated = rec.n_inserted = rec.n_deleted = 0;
    a.recs[i] = rec;
    atomicAdd(sc + CRGPU_NUM_COUNTERS + 4, 1u);            // rows seen (n_total)
    atomicAdd(sc + CRGPU_NUM_COUNTERS + 0, 1u);            // CORE:480-481
    return false;
}

__device__ __forceinline__ void quantify_row(const QuantArgs &a, const int64_t i, unsigned *sc, const VecAcc V)
{
    const int L = a.L, W = a.W, flags = a.flags;
    const int n = a.alnlen[i];
    const OpReader ops{a.ops + i * a.ops_stride, n, a.ops_reversed != 0};
    const uint8_t *amp = a.amp;
    const bool maskN = flags & CRGPU_Q_MASK_N;
    crgpu_read_rec rec;
    rec.cls = 0; rec.pad[0] = rec.pad[1] = rec.pad[2] = 0;
    rec.n_mutated = rec.n_inserted = rec.n_deleted = 0;

    bool unmod = a.unmod_in[i] != 0;
    if (maskN && !unmod) {
        // ignore_n_in_alignment (CORE:2040-2048): markup -> '|' where the amplicon row has N; a
        // markup made of ONE distinct character (whatever it is) marks the read UNMODIFIED.
        // markup classes: 0 '|', 1 '.', 2 ' '
        int idx = 0, first = -1;
        bool uniform = n > 0;
        for (int c = 0; c < n && uniform; ++c) {
            const int op = ops.at(c);
            int cl;
            if (op == OP_INS) cl = 2;
            else {
                cl = amp[idx] == 'N' ? 0 : (op == OP_MATCH ? 0 : op == OP_MISMATCH ? 1 : 2);
                ++idx;
            }
            if (first < 0) first = cl;
            else uniform = cl == first;
        }
        if (uniform) unmod = true;
    }
    atomicAdd(sc + CRGPU_NUM_COUNTERS + 4, 1u);            // rows seen (n_total)
    if (unmod) {                                                     // CORE:480-481
        rec.cls = CRGPU_C_UNMODIFIED;
        a.recs[i] = rec;
        atomicAdd(sc + CRGPU_NUM_COUNTERS + 0, 1u);
        return;
    }

    uint32_t S[MAXW], D[MAXW], I[MAXW], Dk[MAXW];
    for (int w = 0; w < W; ++w) { S[w] = D[w] = I[w] = Dk[w] = 0; }
    const bool doS = !(flags & CRGPU_Q_IGNORE_SUBS), doD = !(flags & CRGPU_Q_IGNORE_DEL), doI = !(flags & CRGPU_Q_IGNORE_INS);

    // ---- pass 1: position sets (CORE:488-533) and the INCLUDE_IDXS test (CORE:553-571) ----
    bool inc_hit = false, ins_splice = false;
    {
        int idx = 0;              // amplicon bases consumed so far == ref_positions of the next base column
        int c = 0;
        while (c < n) {
            { const int run = match_run(ops, c); if (run) { idx += run; c += run; continue; } }
            const int op = ops.at(c);
            if (op == OP_INS) {
                // maximal '-' run in the amplicon row: an insertion [st, en)
                const int st = c;
                while (c < n && ops.at(c) == OP_INS) ++c;
                if (doI) {
                    const int fl = st == 0 ? -1 : idx - 1;                       // ref_positions[max(0, st-1)]
                    const int fr = c < n ? idx : (idx == 0 ? -1 : -idx);         // ref_positions[min(n-1, en)]
                    if (fl >= 0) { if (get_bit(a.inc, fl)) inc_hit = true; if (a.splice && get_bit(a.splice, fl)) ins_splice = true; }
                    if (fr >= 0) { if (get_bit(a.inc, fr)) inc_hit = true; if (a.splice && get_bit(a.splice, fr)) ins_splice = true; }
                    set_bit(I, fl >= 0 ? fl : L + fl);                           // numpy wraps negative indices (Q1)
                    set_bit(I, fr >= 0 ? fr : L + fr);
                }
                continue;
            }
            if (op == OP_DEL) { if (doD) set_bit(D, idx); }
            else if (doS && op == OP_MISMATCH && !(maskN && amp[idx] == 'N')) set_bit(S, idx);
            ++idx; ++c;
        }
        for (int w = 0; w < W; ++w) if ((S[w] | D[w]) & a.inc[w]) inc_hit = true;
    }

    // ---- classification (CORE:535-576) ----
    const bool has_hdr = flags & CRGPU_Q_HAS_HDR;
    int cls;
    {
        const int tr = a.tenths_ref[i];
        const int tp = (has_hdr && a.tenths_rep) ? a.tenths_rep[i] : -1;
        const bool diff_neg = tp >= 0 && tr < tp;                         // score_ref - score_repaired < 0 (NaN -> false)
        if (has_hdr && diff_neg) cls = ((double)tp / 10.0 >= a.hdr_thr) ? CRGPU_C_HDR : CRGPU_C_MIXED;
        else cls = inc_hit ? CRGPU_C_NHEJ : CRGPU_C_UNMODIFIED;
    }
    rec.cls = (uint8_t)cls;
    const bool hide = flags & CRGPU_Q_HIDE_OUTSIDE;
    const bool windowed = cls == CRGPU_C_NHEJ && (flags & CRGPU_Q_WINDOW);

    // ---- vectors that show every event (CORE:578-606) ----
    if (cls == CRGPU_C_MIXED) {
        V.add_bits(CRGPU_V_MUT_MIXED, S, W); V.add_bits(CRGPU_V_DEL_MIXED, D, W);
        V.add_bits(CRGPU_V_INS_MIXED, I, W);
    } else if (cls == CRGPU_C_HDR) {
        V.add_bits(CRGPU_V_MUT_HDR, S, W); V.add_bits(CRGPU_V_DEL_HDR, D, W);
        V.add_bits(CRGPU_V_INS_HDR, I, W);
    } else if (cls == CRGPU_C_NHEJ && !hide) {
        V.add_bits(CRGPU_V_MUT, S, W); V.add_bits(CRGPU_V_DEL, D, W);
        V.add_bits(CRGPU_V_INS, I, W);
    }
    {
        uint32_t A[MAXW];
        for (int w = 0; w < W; ++w) A[w] = S[w] | D[w] | I[w];
        V.add_bits(CRGPU_V_ANY, A, W);                  // also for rows re-classified UNMODIFIED (Q8)
    }
    const int cls_slot = cls == CRGPU_C_UNMODIFIED ? 0 : cls == CRGPU_C_NHEJ ? 1 : cls == CRGPU_C_HDR ? 2 : 3;
    atomicAdd(sc + CRGPU_NUM_COUNTERS + cls_slot, 1u);
    if (cls == CRGPU_C_UNMODIFIED) { a.recs[i] = rec; return; }

    // ---- window filter for NHEJ (CORE:611-641) ----
    if (windowed) for (int w = 0; w < W; ++w) S[w] &= a.inc[w];

    // ---- pass 2: per-run sizes, average-size vectors, exon lengths (CORE:652-673) ----
    const bool fs = flags & CRGPU_Q_FRAMESHIFT;
    int n_ins = 0, n_del = 0, exon_len = 0;
    bool have_len = false, exons_modified = false, any_del_kept = false;
    {
        int idx = 0, c = 0;
        while (c < n) {
            { const int run = match_run(ops, c); if (run) { idx += run; c += run; continue; } }
            const int op = ops.at(c);
            if (op == OP_INS) {
                const int st = c;
                while (c < n && ops.at(c) == OP_INS) ++c;
                if (doI) {
                    const int size = c - st;
                    const int fl = st == 0 ? -1 : idx - 1;
                    const int fr = c < n ? idx : (idx == 0 ? -1 : -idx);
                    bool keep = true;
                    if (windowed) keep = (fl >= 0 && get_bit(a.inc, fl)) || (fr >= 0 && get_bit(a.inc, fr));
                    if (keep) {
                        n_ins += size;
                        const int wl = fl >= 0 ? fl : L + fl, wr = fr >= 0 ? fr : L + fr;
                        V.add(CRGPU_V_AVG_INS, wl, (unsigned)size);
                        if (wr != wl) V.add(CRGPU_V_AVG_INS, wr, (unsigned)size);
                        if (fs && ((fl >= 0 && get_bit(a.exon, fl)) || (fr >= 0 && get_bit(a.exon, fr)))) {
                            exon_len += size; have_len = true; exons_modified = true;        // CORE:665-670
                        }
                    }
                }
                continue;
            }
            if (op == OP_DEL && doD) {
                const int st = c, p0 = idx;
                while (c < n && ops.at(c) == OP_DEL) { ++c; ++idx; }   // amplicon has bases under a deletion run
                const int size = c - st;
                bool keep = true;
                if (windowed) {
                    keep = false;
                    for (int p = p0; p < p0 + size; ++p) if (get_bit(a.inc, p)) { keep = true; break; }
                }
                if (keep) {
                    n_del += size;
                    any_del_kept = true;
                    for (int p = p0; p < p0 + size; ++p) {
                        set_bit(Dk, p);
                        V.add(CRGPU_V_AVG_DEL, p, (unsigned)size);
                    }
                }
                continue;
            }
            ++idx; ++c;
        }
    }
    // deletion_positions_flat is rebuilt only when some deletion survived the window (Q3)
    const uint32_t *Dflat = (windowed && any_del_kept) ? Dk : D;

    if (cls == CRGPU_C_NHEJ && hide) {                                 // CORE:643-649 (Q5)
        V.add_bits(CRGPU_V_MUT, S, W); V.add_bits(CRGPU_V_DEL, Dflat, W);
        V.add_bits(CRGPU_V_INS, I, W);
    }
    int n_mut = 0;
    for (int w = 0; w < W; ++w) n_mut += __popc(S[w]);
    rec.n_mutated = n_mut; rec.n_inserted = n_ins; rec.n_deleted = n_del;
    a.recs[i] = rec;

    // ---- frameshift / splice analysis (CORE:675-725) ----
    if (fs) {
        int del_exon = 0;
        bool sub_exon = false, spliced = ins_splice;
        for (int w = 0; w < W; ++w) {
            del_exon += __popc(Dflat[w] & a.exon[w]);
            if (S[w] & a.exon[w]) sub_exon = true;
            if ((S[w] | Dflat[w]) & a.splice[w]) spliced = true;
        }
        if (del_exon > 0) { exons_modified = true; exon_len -= del_exon; have_len = true; }
        if (sub_exon) exons_modified = true;
        if (spliced) atomicAdd(sc + CRGPU_K_SPLICING_MODIFIED, 1u);
        if (exons_modified) {
            const int key = have_len ? exon_len : 0;
            const int bin = key + a.hist_zero;
            const bool inframe = !have_len || (key % 3) == 0;
            if (inframe) {
                atomicAdd(sc + CRGPU_K_MOD_NON_FRAMESHIFT, 1u);
                if (bin >= 0 && bin < a.hist_len) atomicAdd(a.hist_in + bin, 1ull);
            } else {
                atomicAdd(sc + CRGPU_K_MOD_FRAMESHIFT, 1u);
                if (bin >= 0 && bin < a.hist_len) atomicAdd(a.hist_fs + bin, 1ull);
            }
        } else {
            atomicAdd(sc + CRGPU_K_NON_MOD_NON_FRAMESHIFT, 1u);
            V.add_bits(CRGPU_V_INS_NONCODING, I, W);
            V.add_bits(CRGPU_V_DEL_NONCODING, Dflat, W);
            V.add_bits(CRGPU_V_MUT_NONCODING, S, W);
        }
    }
}

// Rows that need the long way are a fraction of a CTA's rows, spread over its warps: they are gathered first, so that the
// warps that walk op rows are full (thread-per-row over all rows ran with 4.6 of 32 lanes active on average).
// A few CTAs per SM walk over the rows in chunks of QUANT_THREADS (fewer flushes of the shared vectors).
constexpr int QUANT_THREADS = 256;
__global__ void __launch_bounds__(QUANT_THREADS) k_quantify(const QuantArgs a)
{
    extern __shared__ unsigned sv[];                       // [CRGPU_NUM_VECTORS][L]
    __shared__ unsigned sc[NSCALAR];
    __shared__ int heavy[QUANT_THREADS];
    __shared__ int nheavy;
    const int nv = CRGPU_NUM_VECTORS * a.L;
    for (int k = threadIdx.x; k < nv; k += QUANT_THREADS) sv[k] = 0;
    if (threadIdx.x < NSCALAR) sc[threadIdx.x] = 0;
    const VecAcc V{sv, a.L};
    for (int64_t base = (int64_t)blockIdx.x * QUANT_THREADS; base < a.n; base += (int64_t)gridDim.x * QUANT_THREADS) {
        if (threadIdx.x == 0) nheavy = 0;
        __syncthreads();
        const int64_t i = base + threadIdx.x;
        if (i < a.n && (!a.active || (a.active[i] & a.active_bit)) && quantify_row_light(a, i, sc))
            heavy[atomicAdd(&nheavy, 1)] = threadIdx.x;
        __syncthreads();
        if ((int)threadIdx.x < nheavy) quantify_row(a, base + heavy[threadIdx.x], sc, V);
        __syncthreads();
    }
    if (threadIdx.x < NSCALAR && sc[threadIdx.x]) atomicAdd(a.counters + threadIdx.x, (unsigned long long)sc[threadIdx.x]);
    for (int k = threadIdx.x; k < nv; k += QUANT_THREADS)
        if (sv[k]) atomicAdd(a.vectors + k, (unsigned long long)sv[k]);
}

cudaError_t launch_quantify(const QuantArgs &a, cudaStream_t s)
{
    if (a.n <= 0) return cudaSuccess;
    int dev = 0, sms = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    const size_t smem = (size_t)CRGPU_NUM_VECTORS * a.L * sizeof(unsigned);
    if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(k_quantify, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    // every row adds at most L to a 32-bit shared counter: a CTA must see fewer than 2^32 / L rows (2^22 with L = 1024)
    const int64_t chunks = (a.n + QUANT_THREADS - 1) / QUANT_THREADS;
    const int64_t min_grid = (a.n + ((int64_t)1 << 21) - 1) >> 21;
    const unsigned grid = (unsigned)std::max<int64_t>(min_grid, std::min<int64_t>(chunks, (int64_t)sms * 6));
    k_quantify<<<grid, QUANT_THREADS, smem, s>>>(a);
    return cudaGetLastError();
}

// From the alignment records of the forward pass(es): SoA views the quantifier wants plus the
// keep / rescue decision of CORE:1843-1871.  flags_out[i]: bit0 = forward row kept, bit2 = goes to
// the reverse-complement rescue (score_ref < min_identity), bit3 = not aligned (a base outside ACGTN(U)).
__global__ void k_prepare_rows(const crgpu_aln_rec *__restrict__ ref, const crgpu_aln_rec *__restrict__ rep, int64_t n,
                               double min_identity, int32_t *tenths_ref, int32_t *tenths_rep, int32_t *aln_off,
                               int32_t *alnlen, uint8_t *unmod, uint8_t *flags_out, const uint8_t *__restrict__ bad)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (bad && bad[i]) {
        // a base outside ACGTN(U): the read was not aligned (k_encode_pairs); bit3 reports it, nothing else looks at it
        tenths_ref[i] = 0; if (tenths_rep) tenths_rep[i] = -1;
        aln_off[i] = 0; alnlen[i] = 0; unmod[i] = 0;
        if (flags_out) flags_out[i] = 8;
        return;
    }
    const int tr = ref[i].tenths;
    const int tp = rep ? rep[i].tenths : -1;
    tenths_ref[i] = tr;
    if (tenths_rep) tenths_rep[i] = tp;
    aln_off[i] = ref[i].aln_off;
    alnlen[i] = ref[i].alnlen;
    unmod[i] = tr == 1000;                                    // score_ref == 100 (CORE:2014)
    const double sr = (double)tr / 10.0;
    uint8_t f = 0;
    if (sr > min_identity || (tp >= 0 && (double)tp / 10.0 > min_identity)) f |= 1;
    if (sr < min_identity) f |= 4;
    if (flags_out) flags_out[i] = f;
}

// RC rows (compact): keep iff score_ref(rc) > min_identity (CORE:1956-1959 with NaN score_repaired,
// 1976-1978); sets bit1 of kept[read].
__global__ void k_prepare_rc_rows(const crgpu_aln_rec *__restrict__ rc, const int32_t *__restrict__ rc_read, int64_t n,
                                  double min_identity, int32_t *tenths_ref, int32_t *aln_off, int32_t *alnlen,
                                  uint8_t *unmod, uint8_t *active, uint8_t *kept)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int tr = rc[j].tenths;
    tenths_ref[j] = tr;
    aln_off[j] = rc[j].aln_off;
    alnlen[j] = rc[j].alnlen;
    unmod[j] = tr == 1000;
    const bool k = (double)tr / 10.0 > min_identity;
    active[j] = k ? 1 : 0;
    if (k) kept[rc_read[j]] |= 2;
}

cudaError_t launch_prepare_rows(const crgpu_aln_rec *ref, const crgpu_aln_rec *rep, int64_t n, double min_identity,
                                int32_t *tenths_ref, int32_t *tenths_rep, int32_t *aln_off, int32_t *alnlen, uint8_t *unmod,
                                uint8_t *flags_out, const uint8_t *bad, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    k_prepare_rows<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ref, rep, n, min_identity, tenths_ref, tenths_rep, aln_off,
                                                              alnlen, unmod, flags_out, bad);
    return cudaGetLastError();
}

cudaError_t launch_prepare_rc_rows(const crgpu_aln_rec *rc, const int32_t *rc_read, int64_t n, double min_identity,
                                   int32_t *tenths_ref, int32_t *aln_off, int32_t *alnlen, uint8_t *unmod, uint8_t *active,
                                   uint8_t *kept, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    k_prepare_rc_rows<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(rc, rc_read, n, min_identity, tenths_ref, aln_off, alnlen,
                                                                 unmod, active, kept);
    return cudaGetLastError();
}

}  // namespace crgpu
