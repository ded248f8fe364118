// hotpath.cu -- host side of seam S3 (crgpu_quantify) and of the fused path crgpu_align_quantify
// (CRISPResso/CRISPRessoCORE.py:1791-2072 + 2773-2864 in one call, no needle text in between).
#include "crgpu_internal.h"

#include "quant_args.cuh"

namespace crgpu {

cudaError_t launch_quantify(const QuantArgs &a, cudaStream_t s);
cudaError_t launch_rows_to_ops(const uint8_t *ref, const uint8_t *mark, const uint8_t *qry, int64_t slot, const int32_t *aln_off,
                               const int32_t *alnlen, int64_t n, uint32_t *ops, int64_t ops_stride, cudaStream_t s);
cudaError_t launch_prepare_rows(const crgpu_aln_rec *ref, const crgpu_aln_rec *rep, int64_t n, double min_identity,
                                int32_t *tenths_ref, int32_t *tenths_rep, int32_t *aln_off, int32_t *alnlen, uint8_t *unmod,
                                uint8_t *flags_out, const uint8_t *bad, cudaStream_t s);
cudaError_t launch_prepare_rc_rows(const crgpu_aln_rec *rc, const int32_t *rc_read, int64_t n, double min_identity,
                                   int32_t *tenths_ref, int32_t *aln_off, int32_t *alnlen, uint8_t *unmod, uint8_t *active,
                                   uint8_t *kept, cudaStream_t s);

constexpr int N_ACC_TAIL = CRGPU_NUM_COUNTERS + 5;   // counters + 4 class counts + rows seen

// Device accumulators for one amplicon: [vectors 15*L][hist_in][hist_fs][counters tail]
struct Accum {
    int L = 0, hist_len = 0;
    size_t words() const { return (size_t)CRGPU_NUM_VECTORS * L + 2 * (size_t)hist_len + N_ACC_TAIL; }
    unsigned long long *base = nullptr;
    unsigned long long *vectors() const { return base; }
    unsigned long long *hist_in() const { return base + (size_t)CRGPU_NUM_VECTORS * L; }
    unsigned long long *hist_fs() const { return hist_in() + hist_len; }
    unsigned long long *tail() const { return hist_fs() + hist_len; }
};

static void pack_mask(const uint8_t *mask, int L, std::vector<uint32_t> &bits, int W, int which)
{
    for (int w = 0; w < W; ++w) bits[(size_t)which * W + w] = 0;
    if (!mask) return;
    for (int p = 0; p < L; ++p) if (mask[p]) bits[(size_t)which * W + (p >> 5)] |= 1u << (p & 31);
}

static int quant_setup(crgpu_ctx *ctx, const crgpu_quant_params *q, int hist_len, Accum *acc, const uint32_t **d_bits, int *W_out)
{
    const int L = q->amplicon_len;
    if (L < CRGPU_MIN_LEN || L > CRGPU_MAX_AMPLICON) return fail(ctx, CRGPU_E_ARG, "quantify: amplicon_len %d out of range", L);
    if (!q->include_mask) return fail(ctx, CRGPU_E_ARG, "quantify: include_mask is required");
    if ((q->flags & CRGPU_Q_FRAMESHIFT) && (!q->exon_mask || !q->splice_mask || hist_len <= 0))
        return fail(ctx, CRGPU_E_ARG, "quantify: frameshift analysis needs exon_mask, splice_mask and histograms");
    const int W = (L + 31) / 32;
    std::vector<uint32_t> bits((size_t)3 * W);
    pack_mask(q->include_mask, L, bits, W, 0);
    pack_mask(q->exon_mask, L, bits, W, 1);
    pack_mask(q->splice_mask, L, bits, W, 2);
    CK(ctx->q_in[0].reserve(bits.size() * 4));
    CK(push_small(ctx, ctx->q_in[0].p, bits.data(), bits.size() * 4, ctx->stream));
    acc->L = L; acc->hist_len = hist_len > 0 ? hist_len : 0;
    CK(ctx->q_out[0].reserve(acc->words() * 8));
    acc->base = ctx->q_out[0].as<unsigned long long>();
    CK(cudaMemsetAsync(acc->base, 0, acc->words() * 8, ctx->stream));
    *d_bits = ctx->q_in[0].as<uint32_t>();
    *W_out = W;
    return CRGPU_OK;
}

// Add the device accumulators into the caller's host arrays.
static int quant_collect(crgpu_ctx *ctx, const Accum &acc, int64_t *vectors, int64_t *hist_in, int64_t *hist_fs,
                         int64_t *counters, int64_t *class_counts, int64_t *n_total)
{
    std::vector<unsigned long long> h(acc.words());
    CK(fetch_small(ctx, h.data(), acc.base, acc.words() * 8, ctx->stream));
    CK(fetch_wait(ctx, ctx->stream));
    const size_t nv = (size_t)CRGPU_NUM_VECTORS * acc.L;
    if (vectors) for (size_t i = 0; i < nv; ++i) vectors[i] += (int64_t)h[i];
    if (hist_in) for (int i = 0; i < acc.hist_len; ++i) hist_in[i] += (int64_t)h[nv + i];
    if (hist_fs) for (int i = 0; i < acc.hist_len; ++i) hist_fs[i] += (int64_t)h[nv + acc.hist_len + i];
    const unsigned long long *tail = h.data() + nv + 2 * (size_t)acc.hist_len;
    if (counters) for (int i = 0; i < CRGPU_NUM_COUNTERS; ++i) counters[i] += (int64_t)tail[i];
    if (class_counts) for (int i = 0; i < 4; ++i) class_counts[i] += (int64_t)tail[CRGPU_NUM_COUNTERS + i];
    if (n_total) *n_total += (int64_t)tail[CRGPU_NUM_COUNTERS + 4];
    return CRGPU_OK;
}

static void fill_quant_args(QuantArgs *qa, const crgpu_quant_params *q, const Accum &acc, const uint32_t *d_bits, int W,
                            int hist_zero)
{
    qa->L = q->amplicon_len; qa->W = W; qa->flags = q->flags; qa->hdr_thr = q->hdr_perfect_alignment_threshold;
    qa->inc = d_bits; qa->exon = d_bits + W; qa->splice = d_bits + 2 * W;
    qa->vectors = acc.vectors(); qa->hist_in = acc.hist_in(); qa->hist_fs = acc.hist_fs();
    qa->hist_len = acc.hist_len; qa->hist_zero = hist_zero; qa->counters = acc.tail();
}

}  // namespace crgpu

using namespace crgpu;

extern "C" {

static int quantify_impl(crgpu_ctx *ctx, int mem, const crgpu_quant_params *params,
                   const uint8_t *ref_rows, const uint8_t *mark_rows, const uint8_t *qry_rows, int64_t slot,
                   const int32_t *aln_off, const int32_t *alnlen,
                   const int32_t *tenths_ref, const int32_t *tenths_rep, const uint8_t *unmodified_in,
                   int64_t n, crgpu_read_rec *out_recs,
                   int64_t *vectors, int64_t *hist_inframe, int64_t *hist_frameshift, int32_t hist_len,
                   int32_t hist_zero, int64_t *counters)
{
    if (!ctx) return CRGPU_E_ARG;
    if (!params || n < 0 || slot <= 0) return fail(ctx, CRGPU_E_ARG, "crgpu_quantify: bad argument");
    if (n > 0 && (!ref_rows || !mark_rows || !qry_rows || !alnlen || !tenths_ref || !unmodified_in || !out_recs))
        return fail(ctx, CRGPU_E_ARG, "crgpu_quantify: null per-read array");
    if (mem != CRGPU_MEM_HOST && mem != CRGPU_MEM_DEVICE) return fail(ctx, CRGPU_E_ARG, "bad mem");
    timing_reset(ctx);
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    Accum acc; const uint32_t *d_bits; int W;
    int rc = quant_setup(ctx, params, hist_len, &acc, &d_bits, &W);
    if (rc) return rc;
    QuantArgs qa{};
    fill_quant_args(&qa, params, acc, d_bits, W, hist_zero);
    qa.n = n; qa.active = nullptr; qa.active_bit = 0;
    const uint8_t *d_r = ref_rows, *d_m = mark_rows, *d_q = qry_rows;
    const int32_t *d_aoff = aln_off;
    if (mem == CRGPU_MEM_DEVICE) {
        qa.alnlen = alnlen;
        qa.tenths_ref = tenths_ref; qa.tenths_rep = tenths_rep; qa.unmod_in = unmodified_in; qa.recs = out_recs;
    } else if (n > 0) {
        const size_t rb = (size_t)n * slot;
        CK(ctx->sref.reserve(rb)); CK(ctx->smark.reserve(rb)); CK(ctx->sqry.reserve(rb));
        CK(ctx->q_in[1].reserve((size_t)n * 4)); CK(ctx->q_in[2].reserve((size_t)n * 4));
        CK(ctx->q_in[3].reserve((size_t)n * 4)); CK(ctx->q_in[4].reserve((size_t)n * 4));
        CK(ctx->q_in[5].reserve((size_t)n)); CK(ctx->q_out[1].reserve((size_t)n * sizeof(crgpu_read_rec)));
        CK(cudaMemcpyAsync(ctx->sref.p, ref_rows, rb, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->smark.p, mark_rows, rb, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->sqry.p, qry_rows, rb, cudaMemcpyHostToDevice, s));
        if (aln_off) CK(cudaMemcpyAsync(ctx->q_in[1].p, aln_off, (size_t)n * 4, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->q_in[2].p, alnlen, (size_t)n * 4, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->q_in[3].p, tenths_ref, (size_t)n * 4, cudaMemcpyHostToDevice, s));
        if (tenths_rep) CK(cudaMemcpyAsync(ctx->q_in[4].p, tenths_rep, (size_t)n * 4, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->q_in[5].p, unmodified_in, (size_t)n, cudaMemcpyHostToDevice, s));
        d_r = ctx->sref.as<uint8_t>(); d_m = ctx->smark.as<uint8_t>(); d_q = ctx->sqry.as<uint8_t>();
        d_aoff = aln_off ? ctx->q_in[1].as<int32_t>() : nullptr; qa.alnlen = ctx->q_in[2].as<int32_t>();
        qa.tenths_ref = ctx->q_in[3].as<int32_t>(); qa.tenths_rep = tenths_rep ? ctx->q_in[4].as<int32_t>() : nullptr;
        qa.unmod_in = ctx->q_in[5].as<uint8_t>(); qa.recs = ctx->q_out[1].as<crgpu_read_rec>();
    }
    // the three text rows -> 2-bit ops (the caller's markup is already N-masked, CORE:2052 runs before
    // process_df_chunk; the amplicon is therefore not consulted: all-'A' placeholder)
    const int64_t ops_stride = (slot + 15) / 16;
    if (n > 0) {
        CK(ctx->ops.reserve((size_t)n * ops_stride * 4));
        CK(ctx->amp.reserve((size_t)params->amplicon_len));
        CK(cudaMemsetAsync(ctx->amp.p, 'A', (size_t)params->amplicon_len, s));
        span_begin(ctx, T_OTHER);
        CK(launch_rows_to_ops(d_r, d_m, d_q, slot, d_aoff, qa.alnlen, n, ctx->ops.as<uint32_t>(), ops_stride, s));
        span_end(ctx);
    }
    qa.ops = ctx->ops.as<uint32_t>(); qa.ops_stride = ops_stride; qa.ops_reversed = 0; qa.amp = ctx->amp.as<uint8_t>();
    qa.flags &= ~CRGPU_Q_MASK_N;
    span_begin(ctx, T_QUANT);
    CK(launch_quantify(qa, s));
    span_end(ctx);
    if (mem == CRGPU_MEM_HOST && n > 0)
        CK(cudaMemcpyAsync(out_recs, qa.recs, (size_t)n * sizeof(crgpu_read_rec), cudaMemcpyDeviceToHost, s));
    rc = quant_collect(ctx, acc, vectors, hist_inframe, hist_frameshift, counters, nullptr, nullptr);
    if (rc) return rc;
    timing_collect(ctx);
    return CRGPU_OK;
}

static std::string revcomp_upper(const char *s, int n)
{
    std::string out((size_t)n, 'N');
    for (int i = 0; i < n; ++i) {
        char c = s[n - 1 - i];
        switch (c) {
        case 'A': case 'a': c = 'T'; break;
        case 'C': case 'c': c = 'G'; break;
        case 'G': case 'g': c = 'C'; break;
        case 'T': case 't': c = 'A'; break;
        default: c = 'N'; break;
        }
        out[(size_t)i] = c;
    }
    return out;
}

// ---- exact-read shortcut ------------------------------------------------------------------------------------------------
// A read that IS the amplicon (same length, same bases, case aside; amplicon of A C G T only) aligns to it along the diagonal with
// the maximum possible score 5 L -- every column scores at most 5 and a gap costs -- so needle's start cell is (L-1, L-1) and its
// traceback the diagonal (the argument of k_diag_emit with D_n = 5 L); identity 100.0, UNMODIFIED (CORE:2014).  Its alignment to
// the HDR amplicon is the same for every such read.  So these reads -- the largest group of an amplicon-sequencing run -- skip
// the DP altogether: k_mark_exact finds them, ONE representative (the first) stays in the plan, and
// k_emit_exact writes the others' records, ops and text rows and copies the representative's HDR record.
// (eight lanes per read, four reads per warp: the kernel is a chain of dependent loads per read, so what counts is how many
//  reads are in flight; every lane compares 4-byte granules of its read)
__global__ void __launch_bounds__(256) k_mark_exact(const uint8_t *__restrict__ reads, const int64_t *__restrict__ offsets, int64_t n,
                                                    const uint8_t *__restrict__ amp, int La, uint8_t *__restrict__ go, int *rep)
{
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int sub = threadIdx.x & 7;
    const bool in = r < n;
    const int64_t o = in ? offsets[r] : 0;
    bool same = in && (offsets[r + 1] - o) == La;
    if (same) {
        // four bases per step: bit 5 cleared folds the case (the amplicon is plain upper-case A C G T here, and only a / A fold to A ...)
        const uint8_t *b = reads + o;
        const uint32_t *aw = reinterpret_cast<const uint32_t *>(amp);
        for (int x = sub * 4; x < La; x += 32) {
            uint32_t want = aw[x >> 2];
            if (La - x < 4) want &= (1u << (8 * (La - x))) - 1u;
            same &= (load4(b, x, La) & 0xdfdfdfdfu) == want;
        }
    }
    unsigned v = same ? 1u : 0u;
    v &= __shfl_xor_sync(0xffffffffu, v, 1);
    v &= __shfl_xor_sync(0xffffffffu, v, 2);
    v &= __shfl_xor_sync(0xffffffffu, v, 4);
    if (in && sub == 0) {
        go[r] = v ? 0 : 1;                                           // 1: goes through the DP
        // the first exact read is the representative (most candidates see a smaller index already there and skip the atomic)
        if (v && (int)r < *reinterpret_cast<volatile int *>(rep)) atomicMin(rep, (int)r);
    }
}

// reads whose traceback left the band in the amplicon pass (bit 0) / the HDR pass (bit 1): out[0], out[1]
__global__ void k_count_escapes(const uint8_t *__restrict__ esc, int64_t n, int *out)
{
    int c1 = 0, c2 = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint8_t e = esc[i];
        c1 += e & 1; c2 += (e >> 1) & 1;
    }
    c1 = __reduce_add_sync(0xffffffffu, c1);
    c2 = __reduce_add_sync(0xffffffffu, c2);
    if ((threadIdx.x & 31) == 0) { if (c1) atomicAdd(out, c1); if (c2) atomicAdd(out + 1, c2); }
}

__global__ void k_keep_representative(uint8_t *go, const int *rep, int64_t n)
{
    if (*rep >= 0 && *rep < n) go[*rep] = 3;                         // through the DP, and the source of the others' HDR record
}

struct ExactArgs {
    const uint8_t *reads; const int64_t *offsets; int64_t n; const uint8_t *amp; int La; int scale5;
    const uint8_t *go; const int *rep;
    crgpu_aln_rec *aln, *aln_hdr;
    uint32_t *ops; int64_t ops_stride;
    uint8_t *ref, *mark, *qry; int64_t slot;
};

__global__ void __launch_bounds__(256) k_emit_exact(const ExactArgs a)
{
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= a.n || a.go[r]) return;
    const int L = a.La;
    for (int w = lane; w < (L + 15) / 16; w += 32) a.ops[r * a.ops_stride + w] = 0;        // every column a match
    if (a.ref) {
        const uint8_t *b = a.reads + a.offsets[r];
        const int64_t base = r * a.slot + (a.slot - L);                                    // right-aligned, as the walk leaves it
        for (int x = lane; x < L; x += 32) { a.ref[base + x] = a.amp[x]; a.mark[base + x] = '|'; a.qry[base + x] = b[x]; }
    }
    if (lane) return;
    crgpu_aln_rec rec;
    rec.score = 5.0f * (float)L;
    rec.alnlen = L; rec.ident = L; rec.tenths = 1000;
    rec.aln_off = (int32_t)(a.slot - L);
    rec.start1 = L - 1; rec.start2 = L - 1; rec.read_len = L;
    a.aln[r] = rec;
    if (a.aln_hdr) a.aln_hdr[r] = a.aln_hdr[*a.rep];
}

// staged_slot >= 0: the reads are the batch crgpu_stage_reads put into that slot (device memory, filled on the copy stream);
// every output is host memory, as with CRGPU_MEM_HOST
static int align_quantify_impl(crgpu_ctx *ctx, int mem, const char *amplicon, int amplicon_len,
                         const crgpu_path_params *path, const crgpu_quant_params *quant,
                         const uint8_t *reads, const int64_t *offsets, int64_t n, crgpu_path_out *out, int staged_slot = -1)
{
    if (!ctx) return CRGPU_E_ARG;
    if (!amplicon || !path || !quant || !out || n < 0 || (n > 0 && (!reads || !offsets)))
        return fail(ctx, CRGPU_E_ARG, "crgpu_align_quantify: null pointer or negative count");
    if (mem != CRGPU_MEM_HOST && mem != CRGPU_MEM_DEVICE) return fail(ctx, CRGPU_E_ARG, "bad mem");
    if (quant->amplicon_len != amplicon_len) return fail(ctx, CRGPU_E_ARG, "quant->amplicon_len != amplicon_len");
    if (n >= (int64_t)1 << 31) return fail(ctx, CRGPU_E_ARG, "too many reads in one call");
    const bool want_rows = out->ref_rows || out->mark_rows || out->qry_rows;
    if (want_rows && !(out->ref_rows && out->mark_rows && out->qry_rows)) return fail(ctx, CRGPU_E_ARG, "pass all three row buffers or none");
    // (the caller sized its row buffers as n x slot: the library must not pick the slot for it, in either memory mode)
    if (want_rows && out->slot <= 0) return fail(ctx, CRGPU_E_ARG, "out->slot must be set when rows are requested");
    const bool want_rc_rows = out->rc_ref_rows || out->rc_mark_rows || out->rc_qry_rows;
    if (want_rc_rows && !(out->rc_ref_rows && out->rc_mark_rows && out->rc_qry_rows)) return fail(ctx, CRGPU_E_ARG, "pass all three rc row buffers or none");
    const bool has_hdr = path->hdr_amplicon != nullptr && path->hdr_amplicon_len > 0;
    if (has_hdr != ((quant->flags & CRGPU_Q_HAS_HDR) != 0)) return fail(ctx, CRGPU_E_ARG, "CRGPU_Q_HAS_HDR must match path->hdr_amplicon");
    timing_reset(ctx);
    ctx->n_escaped[0] = ctx->n_escaped[1] = 0;
    ctx->n_diag_pairs[0] = ctx->n_diag_pairs[1] = 0;
    out->rc_n = 0;
    if (n == 0) return CRGPU_OK;
    if (!out->kept || !out->aln || !out->recs) return fail(ctx, CRGPU_E_ARG, "kept/aln/recs are required");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const bool host = mem == CRGPU_MEM_HOST;
    ctx->trace.clear();
    trace_mark(ctx, "enter");

    // ---- inputs on the device ----
    const uint8_t *d_reads = reads; const int64_t *d_off = offsets;
    if (staged_slot >= 0) {
        // (pointers: the staging slot's.)  A batch that is run before any call had the chance to start its copy: start it now;
        // the other slot's copy waits for this call's first score-pass launches (flush_stages, run_plan_band)
        if (ctx->stage_pend[staged_slot].on) {
            crgpu_ctx::StagePending other = ctx->stage_pend[staged_slot ^ 1];
            ctx->stage_pend[staged_slot ^ 1].on = false;
            const int frc = flush_stages(ctx);
            ctx->stage_pend[staged_slot ^ 1] = other;
            if (frc) return frc;
        }
        CK(cudaStreamWaitEvent(s, ctx->staged_ev[staged_slot], 0));
    }
    else if (host) {
        const int64_t total = offsets[n];
        CK(ctx->reads.reserve((size_t)std::max<int64_t>(total, 1)));
        CK(ctx->offsets.reserve((size_t)(n + 1) * 8));
        CK(cudaMemcpyAsync(ctx->reads.p, reads, (size_t)total, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->offsets.p, offsets, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, s));
        d_reads = ctx->reads.as<uint8_t>(); d_off = ctx->offsets.as<int64_t>();
    }
    // reads with a base outside ACGTN(U) are flagged here by every build_plan of this call and reported through kept bit 3
    CK(ctx->badbase.reserve((size_t)n));
    CK(cudaMemsetAsync(ctx->badbase.p, 0, (size_t)n, s));
    struct BadScope { crgpu_ctx *c; ~BadScope() { c->d_bad = nullptr; } } bad_scope{ctx};
    ctx->d_bad = ctx->badbase.as<uint8_t>();
    // forward amplicon, upper case: the exact-read test here, the N mask of the quantifier below
    bool plain_acgt = true;
    CK(ctx->aux[6].reserve((size_t)amplicon_len + 16));
    {
        std::string up(amplicon, amplicon + amplicon_len);
        for (auto &ch : up) {
            if (ch >= 'a' && ch <= 'z') ch = (char)(ch - 32);
            plain_acgt &= ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T';
        }
        CK(push_small(ctx, ctx->aux[6].p, up.data(), (size_t)amplicon_len, s));
    }
    trace_mark(ctx, "amplicon-up");
    // ---- 0. reads that ARE the amplicon need no DP (k_mark_exact above): all but one representative leave the plan ----
    const int32_t *d_subset = nullptr;
    int64_t nsub = n;
    uint8_t *d_go = nullptr;
    int *d_rep = nullptr;
    ctx->n_exact = 0;
    if (ctx->exact_shortcut && plain_acgt && !getenv("CRGPU_NO_EXACT")) {
        const size_t sb = select_scratch_bytes(n);
        CK(ctx->exact_go.reserve((size_t)n + 16)); CK(ctx->exact_sel.reserve((size_t)n * 4 + 16)); CK(ctx->alleles.reserve(sb));
        d_go = ctx->exact_go.as<uint8_t>();
        int32_t *d_sel = ctx->exact_sel.as<int32_t>();
        int *d_cnt = reinterpret_cast<int *>(d_sel + n);
        d_rep = d_cnt + 1;
        const int big = 0x7fffffff;
        CK(push_small(ctx, d_rep, &big, 4, s));
        span_begin(ctx, T_ENCODE);
        k_mark_exact<<<(unsigned)((n * 8 + 255) / 256), 256, 0, s>>>(d_reads, d_off, n, ctx->aux[6].as<uint8_t>(), amplicon_len, d_go, d_rep);
        k_keep_representative<<<1, 1, 0, s>>>(d_go, d_rep, n);
        CK(cudaGetLastError());
        span_end(ctx, 2);
        CK(select_flagged(d_go, n, 1, d_sel, d_cnt, ctx->alleles.p, sb, s));
        int h_cnt = 0;
        CK(fetch_small(ctx, &h_cnt, d_cnt, 4, s));
        CK(fetch_wait(ctx, s));
        if (h_cnt < n) { d_subset = d_sel; nsub = h_cnt; ctx->n_exact = n - h_cnt; }
        else d_go = nullptr;                                          // no such read: the plan covers everything
    }
    // pairing plan of the read set (minus the exact reads): shared by the amplicon and the HDR-amplicon pass
    trace_mark(ctx, "exact");
    int rc = build_plan(ctx, d_reads, d_off, d_subset, nsub);
    if (rc) { cudaStreamSynchronize(s); return rc; }
    trace_mark(ctx, "plan");
    const int maxlen = ctx->plan.maxlen;
    const int max_amp = std::max(amplicon_len, has_hdr ? path->hdr_amplicon_len : 0);
    // rows are always produced on the device (the quantifier reads them); slot = caller's or minimal
    int64_t slot = out->slot > 0 ? out->slot : (int64_t)max_amp + maxlen;
    if (slot < (int64_t)amplicon_len + maxlen) return fail(ctx, CRGPU_E_ARG, "slot %lld too small", (long long)slot);

    // ---- device outputs ----
    uint8_t *d_kept; crgpu_aln_rec *d_aln; crgpu_read_rec *d_recs; int32_t *d_trep;
    uint8_t *d_ref, *d_mark, *d_qry;
    // deferred outputs (crgpu_set_deferred_outputs, staged calls only): the per-read arrays leave on the copy stream after
    // the call has returned, from a per-slot set of device buffers that the next call (other slot) does not touch
    const bool deferred = staged_slot >= 0 && ctx->deferred_out;
    if (deferred) {
        DBuf *o = ctx->stage_out[staged_slot];
        if (ctx->out_pend[staged_slot].on) { const int frc = flush_stages(ctx); if (frc) return frc; }
        CK(cudaStreamWaitEvent(s, ctx->out_ev[staged_slot], 0));          // the slot's previous outputs have left
        CK(o[0].reserve((size_t)n)); CK(o[1].reserve((size_t)n * sizeof(crgpu_aln_rec)));
        CK(o[2].reserve((size_t)n * sizeof(crgpu_read_rec))); CK(o[3].reserve((size_t)n * 4));
        d_kept = o[0].as<uint8_t>(); d_aln = o[1].as<crgpu_aln_rec>(); d_recs = o[2].as<crgpu_read_rec>(); d_trep = o[3].as<int32_t>();
    } else if (host) {
        CK(ctx->aux[0].reserve((size_t)n)); CK(ctx->recs.reserve((size_t)n * sizeof(crgpu_aln_rec)));
        CK(ctx->q_out[1].reserve((size_t)n * sizeof(crgpu_read_rec))); CK(ctx->q_in[4].reserve((size_t)n * 4));
        d_kept = ctx->aux[0].as<uint8_t>(); d_aln = ctx->recs.as<crgpu_aln_rec>();
        d_recs = ctx->q_out[1].as<crgpu_read_rec>(); d_trep = ctx->q_in[4].as<int32_t>();
    } else {
        d_kept = out->kept; d_aln = out->aln; d_recs = out->recs;
        if (out->tenths_rep) d_trep = out->tenths_rep;
        else { CK(ctx->q_in[4].reserve((size_t)n * 4)); d_trep = ctx->q_in[4].as<int32_t>(); }
    }
    // text rows only when the caller wants them; the quantifier reads the walker's 2-bit ops
    d_ref = d_mark = d_qry = nullptr;
    if (!host && want_rows) { d_ref = out->ref_rows; d_mark = out->mark_rows; d_qry = out->qry_rows; }
    else if (want_rows) {
        const size_t rb = (size_t)n * slot;
        CK(ctx->sref.reserve(rb)); CK(ctx->smark.reserve(rb)); CK(ctx->sqry.reserve(rb));
        d_ref = ctx->sref.as<uint8_t>(); d_mark = ctx->smark.as<uint8_t>(); d_qry = ctx->sqry.as<uint8_t>();
    }
    const int64_t ops_stride = (slot + 15) / 16;
    CK(ctx->ops.reserve((size_t)n * ops_stride * 4));
    uint32_t *d_ops = ctx->ops.as<uint32_t>();

    // ---- 1. forward alignments: amplicon (+ HDR amplicon, CORE:1810-1828) ----
    int64_t cells = 0, cells_computed = 0;
    crgpu_aln_rec *d_aln_hdr = nullptr;
    bool amp_done = false, hdr_done = !has_hdr, banded = false;
    uint8_t *d_esc = nullptr;
    if (has_hdr) {
        CK(ctx->aux[1].reserve((size_t)n * sizeof(crgpu_aln_rec)));
        d_aln_hdr = ctx->aux[1].as<crgpu_aln_rec>();
    }
    if (ctx->band_B > 0 && ctx->band_holdoff > 0) --ctx->band_holdoff;          // recent calls escaped too often: single-pass fill
    else if (ctx->band_B > 0) {
        // banded two-pass fill; reads whose traceback leaves the band are flagged in d_esc and re-aligned below
        CK(ctx->escaped.reserve((size_t)n));
        d_esc = ctx->escaped.as<uint8_t>();
        CK(cudaMemsetAsync(d_esc, 0, (size_t)n, s));
        // diagonal shortcut: per-read bits 1 / 2 = the alignment vs the amplicon / the HDR amplicon needed no traceback
        CK(ctx->fastflags.reserve((size_t)n));
        uint8_t *d_fast = ctx->fastflags.as<uint8_t>();
        CK(cudaMemsetAsync(d_fast, 0, (size_t)n, s));
        bool done = false;
        if (has_hdr && path->hdr_amplicon_len == amplicon_len) {
            rc = run_plan_band(ctx, amplicon, path->hdr_amplicon, amplicon_len, d_reads, d_off, path->gapopen, path->gapextend,
                               d_aln, d_aln_hdr, d_ref, d_mark, d_qry, slot, &cells, &cells_computed, d_ops, ops_stride, d_esc, 1, d_fast, &done);
            if (rc) { cudaStreamSynchronize(s); return rc; }
            if (done) amp_done = hdr_done = banded = true;
        }
        if (!amp_done) {
            rc = run_plan_band(ctx, amplicon, nullptr, amplicon_len, d_reads, d_off, path->gapopen, path->gapextend,
                               d_aln, nullptr, d_ref, d_mark, d_qry, slot, &cells, &cells_computed, d_ops, ops_stride, d_esc, 1, d_fast, &done);
            if (rc) { cudaStreamSynchronize(s); return rc; }
            if (done) {
                amp_done = banded = true;
                if (has_hdr) {
                    rc = run_plan_band(ctx, path->hdr_amplicon, nullptr, path->hdr_amplicon_len, d_reads, d_off, path->gapopen,
                                       path->gapextend, d_aln_hdr, nullptr, nullptr, nullptr, nullptr, slot, &cells, &cells_computed,
                                       nullptr, 0, d_esc, 2, d_fast, &done);
                    if (rc) { cudaStreamSynchronize(s); return rc; }
                    hdr_done = done;
                }
            }
        }
    }
    if (!amp_done && has_hdr && path->hdr_amplicon_len == amplicon_len) {
        bool dual_done = false;
        rc = run_plan_dual(ctx, amplicon, path->hdr_amplicon, amplicon_len, d_reads, d_off, path->gapopen, path->gapextend,
                           d_aln, d_aln_hdr, d_ref, d_mark, d_qry, slot, &cells, &cells_computed, d_ops, ops_stride, &dual_done);
        if (rc) { cudaStreamSynchronize(s); return rc; }
        if (dual_done) amp_done = hdr_done = true;
    }
    if (!amp_done) {
        int64_t c0 = 0;
        rc = run_plan(ctx, amplicon, amplicon_len, d_reads, d_off, nullptr, 0, path->gapopen, path->gapextend, d_aln, d_ref, d_mark,
                      d_qry, slot, &c0, d_ops, ops_stride);
        if (rc) { cudaStreamSynchronize(s); return rc; }
        cells += c0; cells_computed += c0;
    }
    if (!hdr_done) {
        int64_t c0 = 0;
        rc = run_plan(ctx, path->hdr_amplicon, path->hdr_amplicon_len, d_reads, d_off, nullptr, 0, path->gapopen,
                      path->gapextend, d_aln_hdr, nullptr, nullptr, nullptr, slot, &c0);
        if (rc) { cudaStreamSynchronize(s); return rc; }
        cells += c0; cells_computed += c0;
    }
    { const int frc = flush_stages(ctx); if (frc) return frc; }          // (paths without a banded pass)
    trace_mark(ctx, "main-issued");
    if (banded) {
        // reads that escaped the band: compacted on the device, re-aligned with the single-pass fill
        const size_t sb = select_scratch_bytes(n);
        CK(ctx->aux[3].reserve((size_t)n * 4 + 16));
        CK(ctx->alleles.reserve(sb));
        int32_t *d_sel = ctx->aux[3].as<int32_t>();
        int *d_cnt = reinterpret_cast<int *>(d_sel + n);
        // ONE plan over the reads that escaped in either pass: a read that escaped only one of them is re-aligned to both
        // amplicons (the single-pass fill gives the same records as the band did; cheaper than a second plan and its syncs)
        CK(cudaMemsetAsync(d_cnt + 1, 0, 8, s));
        k_count_escapes<<<(unsigned)std::min<int64_t>((n + 255) / 256, 1024), 256, 0, s>>>(d_esc, n, d_cnt + 1);
        CK(cudaGetLastError());
        ctx->launches[T_OTHER] += 1;
        CK(select_flagged(d_esc, n, has_hdr ? 3 : 1, d_sel, d_cnt, ctx->alleles.p, sb, s));
        int h_cnt[3] = {0, 0, 0};                           // union, amplicon pass, HDR pass
        CK(fetch_small(ctx, h_cnt, d_cnt, 12, s));
        CK(fetch_wait(ctx, s));
        trace_mark(ctx, "main-done+esc-count");
        ctx->n_escaped[0] = h_cnt[1];
        ctx->n_escaped[1] = has_hdr ? h_cnt[2] : 0;
        // a read set whose tracebacks mostly leave the band pays for both fills: skip the band for the next 8 calls
        if ((int64_t)std::max(h_cnt[1], h_cnt[2]) * 4 > n) ctx->band_holdoff = 8;
        if (h_cnt[0] > 0) {
            int64_t c0 = 0;
            rc = build_plan(ctx, d_reads, d_off, d_sel, h_cnt[0]);
            // both passes are a single wave of a few thousand pairs, bound by the latency of one pair's fill and walk: they
            // run side by side (the HDR pass as lane 1: third stream, second scratch set)
            const bool both = h_cnt[1] > 0 && has_hdr && h_cnt[2] > 0;
            if (rc == CRGPU_OK && has_hdr && h_cnt[2] > 0)
                rc = run_plan(ctx, path->hdr_amplicon, path->hdr_amplicon_len, d_reads, d_off, nullptr, 0, path->gapopen,
                              path->gapextend, d_aln_hdr, nullptr, nullptr, nullptr, slot, &c0, nullptr, 0, both ? 1 : 0);
            if (rc == CRGPU_OK && h_cnt[1] > 0)
                rc = run_plan(ctx, amplicon, amplicon_len, d_reads, d_off, nullptr, 0, path->gapopen, path->gapextend, d_aln, d_ref,
                              d_mark, d_qry, slot, &c0, d_ops, ops_stride);
            if (rc == CRGPU_OK && both) CK(cudaStreamWaitEvent(s, ctx->walk_done[1], 0));
            if (rc) { cudaStreamSynchronize(s); return rc; }
            cells_computed += c0;
        }
    }

    trace_mark(ctx, "esc-issued");
    if (d_go) {
        // the exact reads: alignment = the diagonal (no DP), HDR record = the representative's
        ExactArgs ea;
        ea.reads = d_reads; ea.offsets = d_off; ea.n = n; ea.amp = ctx->aux[6].as<uint8_t>(); ea.La = amplicon_len; ea.scale5 = 0;
        ea.go = d_go; ea.rep = d_rep; ea.aln = d_aln; ea.aln_hdr = d_aln_hdr; ea.ops = d_ops; ea.ops_stride = ops_stride;
        ea.ref = d_ref; ea.mark = d_mark; ea.qry = d_qry; ea.slot = slot;
        span_begin(ctx, T_WALK);
        k_emit_exact<<<(unsigned)((n * 32 + 255) / 256), 256, 0, s>>>(ea);
        CK(cudaGetLastError());
        span_end(ctx, 1);
        cells += (int64_t)(has_hdr ? 2 : 1) * amplicon_len * amplicon_len * ctx->n_exact;
    }

    // ---- 2. keep / rescue decision ----
    CK(ctx->q_in[1].reserve((size_t)n * 4)); CK(ctx->q_in[2].reserve((size_t)n * 4)); CK(ctx->q_in[3].reserve((size_t)n * 4));
    CK(ctx->q_in[5].reserve((size_t)n));
    int32_t *d_aln_off = ctx->q_in[1].as<int32_t>(), *d_alnlen = ctx->q_in[2].as<int32_t>(), *d_tref = ctx->q_in[3].as<int32_t>();
    uint8_t *d_unmod = ctx->q_in[5].as<uint8_t>();
    span_begin(ctx, T_OTHER);
    CK(launch_prepare_rows(d_aln, d_aln_hdr, n, path->min_identity_score, d_tref, d_trep, d_aln_off, d_alnlen, d_unmod, d_kept,
                           ctx->d_bad, s));
    span_end(ctx);

    // reads with score_ref < min_identity, in read order: compacted on the device (cub::DeviceSelect),
    // only their count and indices visit the host
    std::vector<int32_t> rc_read;
    if (path->rc_rescue) {
        const size_t sb = select_scratch_bytes(n);
        CK(ctx->aux[3].reserve((size_t)n * 4 + 16));
        CK(ctx->alleles.reserve(sb));
        int32_t *d_sel = ctx->aux[3].as<int32_t>();
        int *d_cnt = reinterpret_cast<int *>(d_sel + n);
        CK(select_flagged(d_kept, n, 4, d_sel, d_cnt, ctx->alleles.p, sb, s));
        int h_cnt = 0;
        CK(fetch_small(ctx, &h_cnt, d_cnt, 4, s));
        CK(fetch_wait(ctx, s));
        if (h_cnt > 0) {
            rc_read.resize((size_t)h_cnt);
            CK(fetch_small(ctx, rc_read.data(), d_sel, (size_t)h_cnt * 4, s));
            CK(fetch_wait(ctx, s));
        }
    }
    const int64_t nrc = (int64_t)rc_read.size();
    out->rc_n = nrc;
    trace_mark(ctx, "rc-select");
    if (nrc > out->rc_cap && (out->rc_read || out->rc_aln || out->rc_recs || want_rc_rows))
        return fail(ctx, CRGPU_E_ARG, "rc_cap %lld < %lld reads re-aligned to the reverse complement", (long long)out->rc_cap, (long long)nrc);

    // ---- 3. quantification of the forward rows ----
    Accum acc; const uint32_t *d_bits; int W;
    rc = quant_setup(ctx, quant, out->hist_len, &acc, &d_bits, &W);
    if (rc) return rc;
    QuantArgs qa{};
    fill_quant_args(&qa, quant, acc, d_bits, W, out->hist_zero);
    // (forward amplicon, upper case, for the N mask: uploaded to ctx->aux[6] at the top of the call)
    qa.ops = d_ops; qa.ops_stride = ops_stride; qa.ops_reversed = 1; qa.amp = ctx->aux[6].as<uint8_t>(); qa.alnlen = d_alnlen;
    qa.tenths_ref = d_tref; qa.tenths_rep = has_hdr ? d_trep : nullptr; qa.unmod_in = d_unmod;
    qa.active = d_kept; qa.active_bit = 1; qa.n = n; qa.recs = d_recs;
    if (amplicon_len > 0) {
        bool hasN = false;
        for (int i = 0; i < amplicon_len; ++i) hasN |= (amplicon[i] == 'N' || amplicon[i] == 'n');
        if (hasN) qa.flags |= CRGPU_Q_MASK_N;                      // CORE:2033
    }
    CK(cudaMemsetAsync(d_recs, 0, (size_t)n * sizeof(crgpu_read_rec), s));
    span_begin(ctx, T_QUANT);
    CK(launch_quantify(qa, s));
    span_end(ctx);

    // ---- 4. reverse-complement rescue (CORE:1873-2000) ----
    crgpu_aln_rec *d_rc_aln = nullptr; crgpu_read_rec *d_rc_recs = nullptr;
    uint8_t *d_rc_ref = nullptr, *d_rc_mark = nullptr, *d_rc_qry = nullptr;
    if (nrc > 0) {
        const std::string amp_rc = revcomp_upper(amplicon, amplicon_len);
        // out_index: read -> compact RC row
        std::vector<int32_t> out_index((size_t)n, -1);
        for (int64_t j = 0; j < nrc; ++j) out_index[(size_t)rc_read[(size_t)j]] = (int32_t)j;
        CK(ctx->aux[2].reserve((size_t)n * 4));          // aux[3] already holds rc_read (device compaction above)
        CK(cudaMemcpyAsync(ctx->aux[2].p, out_index.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
        const size_t rb = (size_t)nrc * slot;
        if (!host && out->rc_aln) d_rc_aln = out->rc_aln;
        else { CK(ctx->aux[4].reserve((size_t)nrc * sizeof(crgpu_aln_rec))); d_rc_aln = ctx->aux[4].as<crgpu_aln_rec>(); }
        if (!host && out->rc_recs) d_rc_recs = out->rc_recs;
        else { CK(ctx->aux[5].reserve((size_t)nrc * sizeof(crgpu_read_rec))); d_rc_recs = ctx->aux[5].as<crgpu_read_rec>(); }
        if (!host && want_rc_rows) { d_rc_ref = out->rc_ref_rows; d_rc_mark = out->rc_mark_rows; d_rc_qry = out->rc_qry_rows; }
        else if (want_rc_rows) {
            CK(ctx->q_out[2].reserve(rb * 3));
            d_rc_ref = ctx->q_out[2].as<uint8_t>(); d_rc_mark = d_rc_ref + rb; d_rc_qry = d_rc_mark + rb;
        }
        CK(ctx->ops_rc.reserve((size_t)nrc * ops_stride * 4));
        int64_t c_rc = 0;
        rc = build_plan(ctx, d_reads, d_off, ctx->aux[3].as<int32_t>(), nrc);
        if (rc == CRGPU_OK)
            rc = run_plan(ctx, amp_rc.data(), amplicon_len, d_reads, d_off, ctx->aux[2].as<int32_t>(), 1, path->gapopen,
                          path->gapextend, d_rc_aln, d_rc_ref, d_rc_mark, d_rc_qry, slot, &c_rc, ctx->ops_rc.as<uint32_t>(),
                          ops_stride);
        cells += c_rc; cells_computed += c_rc;
        if (rc) { cudaStreamSynchronize(s); return rc; }
        // per-row SoA for the quantifier (separate scratch: the forward views are still in use by the stream)
        CK(ctx->q_out[3].reserve((size_t)nrc * 14));
        int32_t *r_tref = ctx->q_out[3].as<int32_t>();
        int32_t *r_off = r_tref + nrc, *r_len = r_off + nrc;
        uint8_t *r_unmod = reinterpret_cast<uint8_t *>(r_len + nrc), *r_active = r_unmod + nrc;
        span_begin(ctx, T_OTHER);
        CK(launch_prepare_rc_rows(d_rc_aln, ctx->aux[3].as<int32_t>(), nrc, path->min_identity_score, r_tref, r_off, r_len,
                                  r_unmod, r_active, d_kept, s));
        span_end(ctx);
        QuantArgs qr = qa;
        qr.ops = ctx->ops_rc.as<uint32_t>(); qr.ops_reversed = 0; qr.alnlen = r_len;    // RC rows are walked in forward-strand order
        qr.tenths_ref = r_tref; qr.tenths_rep = nullptr; qr.unmod_in = r_unmod; qr.active = r_active; qr.active_bit = 1;
        qr.n = nrc; qr.recs = d_rc_recs;
        CK(cudaMemsetAsync(d_rc_recs, 0, (size_t)nrc * sizeof(crgpu_read_rec), s));
        span_begin(ctx, T_QUANT);
        CK(launch_quantify(qr, s));
        span_end(ctx);
    }

    trace_mark(ctx, "quant+rc-issued");
    // ---- 4b. allele table: group the kept rows on the device (CORE:2923-2946) ----
    if (out->allele_cap > 0) {
        if (!out->allele_row || !out->allele_count) return fail(ctx, CRGPU_E_ARG, "allele_row / allele_count are required with allele_cap > 0");
        const int64_t m = n + nrc;
        const size_t sb = allele_scratch_bytes(m);
        CK(ctx->alleles.reserve(sb));
        int32_t *d_rep, *d_cnt; int *d_nruns, *d_aerr; uint64_t *d_kp;
        const bool reads_upper = staged_slot >= 0 && ctx->stage_pend[staged_slot].format == CRGPU_READS_BAM4;
        span_begin(ctx, T_OTHER);
        CK(allele_groups(d_reads, d_off, n, d_kept, nrc ? ctx->aux[3].as<int32_t>() : nullptr, nrc, d_ops,
                         nrc ? ctx->ops_rc.as<uint32_t>() : nullptr, ops_stride, d_aln, d_rc_aln, d_recs, d_rc_recs,
                         // (reads that came as BAM 4-bit codes were unpacked to upper case)
                         reads_upper ? 1 : 0,
                         ctx->alleles.p, sb, s, &d_rep, &d_cnt, &d_nruns, &d_aerr, &d_kp));
        span_end(ctx, reads_upper ? 5 : 6);     // (k_any_lower,) k_hash_rows, k_check_groups, k_group_reps, k_gather_i32, k_allele_keys (+ cub's own kernels, not counted)
        int h3[3] = {0, 0, 0};                                  // runs, collision flag, "one run is the rows that were not kept"
        CK(fetch_small(ctx, h3, d_nruns, 12, s));
        CK(fetch_wait(ctx, s));
        trace_mark(ctx, "allele-kernels");
        if (h3[1]) return fail(ctx, CRGPU_E_CUDA, "allele grouping: 64-bit hash collision between different alleles");
        const int64_t take = std::min<int64_t>(out->allele_cap, h3[0]);
        std::vector<int32_t> hrep((size_t)take), hcnt((size_t)take);
        if (take > 0) {
            CK(fetch_small(ctx, hrep.data(), d_rep, (size_t)take * 4, s));
            CK(fetch_small(ctx, hcnt.data(), d_cnt, (size_t)take * 4, s));
            if (out->allele_key) CK(fetch_small(ctx, out->allele_key, d_kp, (size_t)take * 16, s));
            CK(fetch_wait(ctx, s));
        }
        int64_t na = 0;
        for (int64_t i = 0; i < take; ++i) {
            if (hcnt[(size_t)i] <= 0) break;
            out->allele_row[na] = hrep[(size_t)i];
            out->allele_count[na] = hcnt[(size_t)i];
            ++na;
        }
        // runs = alleles (+ 1 run of rows that were not kept, which sorts last with count 0)
        out->allele_n = (int64_t)h3[0] - (h3[2] ? 1 : 0);
    }

    trace_mark(ctx, "alleles");
    // ---- 5. results ----
    if (host) {
        if (deferred) {
            // ... valid after crgpu_sync: the copies start inside the NEXT call, once its first score-pass launches are queued
            // (flush_stages) -- right now they would share PCIe with that call's first launches and small transfers
            crgpu_ctx::OutPending &op = ctx->out_pend[staged_slot];
            op.n = 0;
            auto add = [&](void *h, const void *d, size_t bytes) { if (h && bytes) { op.h[op.n] = h; op.d[op.n] = d; op.bytes[op.n] = bytes; ++op.n; } };
            add(out->kept, d_kept, (size_t)n);
            add(out->aln, d_aln, (size_t)n * sizeof(crgpu_aln_rec));
            add(out->recs, d_recs, (size_t)n * sizeof(crgpu_read_rec));
            add(out->tenths_rep, d_trep, (size_t)n * 4);
            CK(cudaEventRecord(ctx->out_ready[staged_slot], s));
            op.on = true;
        } else {
            CK(cudaMemcpyAsync(out->kept, d_kept, (size_t)n, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(out->aln, d_aln, (size_t)n * sizeof(crgpu_aln_rec), cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(out->recs, d_recs, (size_t)n * sizeof(crgpu_read_rec), cudaMemcpyDeviceToHost, s));
            if (out->tenths_rep) CK(cudaMemcpyAsync(out->tenths_rep, d_trep, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
        }
        if (want_rows) {
            if (out->slot != slot) return fail(ctx, CRGPU_E_ARG, "out->slot must be set when rows are requested");
            const size_t rb = (size_t)n * slot;
            CK(cudaMemcpyAsync(out->ref_rows, d_ref, rb, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(out->mark_rows, d_mark, rb, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(out->qry_rows, d_qry, rb, cudaMemcpyDeviceToHost, s));
        }
        if (nrc > 0) {
            if (out->rc_read) memcpy(out->rc_read, rc_read.data(), (size_t)nrc * 4);
            if (out->rc_aln) CK(cudaMemcpyAsync(out->rc_aln, d_rc_aln, (size_t)nrc * sizeof(crgpu_aln_rec), cudaMemcpyDeviceToHost, s));
            if (out->rc_recs) CK(cudaMemcpyAsync(out->rc_recs, d_rc_recs, (size_t)nrc * sizeof(crgpu_read_rec), cudaMemcpyDeviceToHost, s));
            if (want_rc_rows) {
                const size_t rb = (size_t)nrc * slot;
                CK(cudaMemcpyAsync(out->rc_ref_rows, d_rc_ref, rb, cudaMemcpyDeviceToHost, s));
                CK(cudaMemcpyAsync(out->rc_mark_rows, d_rc_mark, rb, cudaMemcpyDeviceToHost, s));
                CK(cudaMemcpyAsync(out->rc_qry_rows, d_rc_qry, rb, cudaMemcpyDeviceToHost, s));
            }
        }
    } else if (nrc > 0 && out->rc_read) {
        CK(cudaMemcpyAsync(out->rc_read, rc_read.data(), (size_t)nrc * 4, cudaMemcpyHostToDevice, s));
    }
    rc = quant_collect(ctx, acc, out->vectors, out->hist_inframe, out->hist_frameshift, out->counters, out->class_counts,
                       &out->n_total);
    if (rc) return rc;
    out->n_cells += cells;
    out->n_cells_computed += cells_computed;
    timing_collect(ctx);
    trace_mark(ctx, "collect");
    trace_dump(ctx);
    return CRGPU_OK;
}

// ---- the exported entry points: device guard + "no work of a failed call is left running" (ApiGuard, crgpu_internal.h)
int crgpu_align_quantify_staged(crgpu_ctx *ctx, int slot, const char *amplicon, int amplicon_len,
                                const crgpu_path_params *path, const crgpu_quant_params *quant, crgpu_path_out *out)
{
    if (!ctx) return CRGPU_E_ARG;
    ApiGuard guard(ctx);
    if (slot < 0 || slot > 1 || ctx->stage_n[slot] < 0) return fail(ctx, CRGPU_E_ARG, "crgpu_align_quantify_staged: nothing staged in slot %d", slot);
    const int64_t n = ctx->stage_n[slot];
    ctx->stage_n[slot] = -1;                                        // consumed
    return guard.done(align_quantify_impl(ctx, CRGPU_MEM_HOST, amplicon, amplicon_len, path, quant, ctx->stage_reads[slot].as<uint8_t>(),
                                          ctx->stage_off[slot].as<int64_t>(), n, out, slot));
}

int crgpu_quantify(crgpu_ctx *ctx, int mem, const crgpu_quant_params *params,
                   const uint8_t *ref_rows, const uint8_t *mark_rows, const uint8_t *qry_rows, int64_t slot,
                   const int32_t *aln_off, const int32_t *alnlen,
                   const int32_t *tenths_ref, const int32_t *tenths_rep, const uint8_t *unmodified_in,
                   int64_t n, crgpu_read_rec *out_recs,
                   int64_t *vectors, int64_t *hist_inframe, int64_t *hist_frameshift, int32_t hist_len,
                   int32_t hist_zero, int64_t *counters)
{
    if (!ctx) return CRGPU_E_ARG;
    ApiGuard guard(ctx);
    return guard.done(quantify_impl(ctx, mem, params, ref_rows, mark_rows, qry_rows, slot, aln_off, alnlen, tenths_ref, tenths_rep, unmodified_in, n, out_recs, vectors, hist_inframe, hist_frameshift, hist_len, hist_zero, counters));
}

int crgpu_align_quantify(crgpu_ctx *ctx, int mem, const char *amplicon, int amplicon_len,
                         const crgpu_path_params *path, const crgpu_quant_params *quant,
                         const uint8_t *reads, const int64_t *offsets, int64_t n, crgpu_path_out *out)
{
    if (!ctx) return CRGPU_E_ARG;
    ApiGuard guard(ctx);
    return guard.done(align_quantify_impl(ctx, mem, amplicon, amplicon_len, path, quant, reads, offsets, n, out));
}

}  // extern "C"
