// crgpu_api.cu -- the C ABI declared in include/crgpu.h: context, S1, S2, batching, host-side pairing.
#include "crgpu_internal.h"

using namespace crgpu;

extern "C" {

int crgpu_abi_version(void) { return CRGPU_ABI_VERSION; }

int crgpu_create(crgpu_ctx **out, int device)
{
    if (!out) return CRGPU_E_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return CRGPU_E_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return CRGPU_E_CUDA;
    crgpu_ctx *c = new crgpu_ctx();
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return CRGPU_E_CUDA; }
    if (prop.major != 10) { delete c; return CRGPU_E_CUDA; }   // sm_100a SASS only: no other target, no fallback
    c->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return CRGPU_E_CUDA; }
    *out = c;
    return CRGPU_OK;
}

void crgpu_destroy(crgpu_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    DBuf *all[] = {&c->reads, &c->offsets, &c->amp, &c->prof, &c->pc, &c->pc_off, &c->plen, &c->pair_lo, &c->pair_hi,
                   &c->tb_off, &c->tb, &c->lastrow, &c->lastcol, &c->errflag, &c->recs, &c->sref, &c->smark, &c->sqry};
    for (DBuf *b : all) b->release();
    for (auto &b : c->q_in) b.release();
    for (auto &b : c->q_out) b.release();
    for (auto &b : c->aux) b.release();
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    cudaStreamDestroy(c->stream);
    delete c;
}

const char *crgpu_last_error(const crgpu_ctx *c) { return c ? c->err.c_str() : "null context"; }

int crgpu_set_traceback_budget(crgpu_ctx *c, size_t bytes)
{
    if (!c || bytes < ((size_t)16 << 20)) return CRGPU_E_ARG;
    c->tb_budget = bytes;
    return CRGPU_OK;
}

int crgpu_last_timing(const crgpu_ctx *c, float out_ms[6], int64_t out_launches[6])
{
    if (!c) return CRGPU_E_ARG;
    for (int i = 0; i < T_N; ++i) {
        if (out_ms) out_ms[i] = c->ms[i];
        if (out_launches) out_launches[i] = c->launches[i];
    }
    return CRGPU_OK;
}

int crgpu_sync(crgpu_ctx *ctx)
{
    if (!ctx) return CRGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return CRGPU_OK;
}

void *crgpu_stream(crgpu_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int crgpu_qualfilter(crgpu_ctx *ctx, int mem, const uint8_t *qual, const int64_t *offsets, int64_t n,
                     int min_mean_q, int min_single_q, uint8_t *keep)
{
    if (!ctx) return CRGPU_E_ARG;
    if (n < 0 || (n > 0 && (!qual || !offsets || !keep))) return fail(ctx, CRGPU_E_ARG, "crgpu_qualfilter: bad argument");
    timing_reset(ctx);
    if (n == 0) return CRGPU_OK;
    CK(cudaSetDevice(ctx->device));
    const uint8_t *d_q = qual; const int64_t *d_o = offsets; uint8_t *d_k = keep;
    if (mem == CRGPU_MEM_HOST) {
        const int64_t total = offsets[n];
        CK(ctx->reads.reserve((size_t)std::max<int64_t>(total, 1)));
        CK(ctx->offsets.reserve((size_t)(n + 1) * 8));
        CK(ctx->aux[0].reserve((size_t)n));
        CK(cudaMemcpyAsync(ctx->reads.p, qual, (size_t)total, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->offsets.p, offsets, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        d_q = ctx->reads.as<uint8_t>(); d_o = ctx->offsets.as<int64_t>(); d_k = ctx->aux[0].as<uint8_t>();
    } else if (mem != CRGPU_MEM_DEVICE) return fail(ctx, CRGPU_E_ARG, "bad mem");
    span_begin(ctx, T_QUAL);
    CK(launch_qualfilter(d_q, d_o, n, min_mean_q, min_single_q, d_k, ctx->num_sms, ctx->stream));
    span_end(ctx);
    if (mem == CRGPU_MEM_HOST) CK(cudaMemcpyAsync(keep, d_k, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    timing_collect(ctx);
    return CRGPU_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Alignment core shared by crgpu_align and crgpu_align_quantify.  All pointers are DEVICE
// pointers except h_offsets (host copy of the offsets, needed to pair reads by length).
// ---------------------------------------------------------------------------------------------
static int host_code(char c)
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
static int host_ednafull(int ca, int cb)
{
    if (ca == 4 && cb == 4) return -1;
    if (ca == 4 || cb == 4) return -2;
    return ca == cb ? 5 : -4;
}

// Exact integer scaling of the gap penalties: needle computes in float32, so the result is only
// reproducible with integers when both penalties are dyadic rationals.
static bool scale_penalties(double gapopen, double gapextend, int *scale, int *open_s, int *ext_s)
{
    for (int s = 1; s <= 8; s *= 2) {
        const double o = gapopen * s, e = gapextend * s;
        if (o == std::floor(o) && e == std::floor(e)) {
            *scale = s; *open_s = (int)o; *ext_s = (int)e;
            return true;
        }
    }
    return false;
}

namespace crgpu {

// subset: optional list of read indices to align (device reads/offsets cover ALL reads); when
// null all n_total reads are aligned.  recs/strings are indexed by ORIGINAL read index.
int align_core(crgpu_ctx *ctx, const char *amplicon, int La, const uint8_t *d_reads, const int64_t *d_offsets,
               const int64_t *h_offsets, const int32_t *subset, int64_t nsub, const int32_t *d_out_index, int rc_out,
               double gapopen, double gapextend, crgpu_aln_rec *d_recs, uint8_t *d_ref, uint8_t *d_mark, uint8_t *d_qry,
               int64_t slot, int64_t *n_cells)
{
    if (La < CRGPU_MIN_LEN || La > CRGPU_MAX_AMPLICON)
        return fail(ctx, CRGPU_E_ALIGN, "amplicon length %d outside [%d, %d]", La, CRGPU_MIN_LEN, CRGPU_MAX_AMPLICON);
    int scale, open_s, ext_s;
    if (!scale_penalties(gapopen, gapextend, &scale, &open_s, &ext_s))
        return fail(ctx, CRGPU_E_ALIGN, "gapopen=%g gapextend=%g are not multiples of 1/8: needle's float32 result is not "
                                        "reproducible exactly", gapopen, gapextend);
    if (gapopen < 4.0 || gapextend < 0.0 || gapextend > gapopen)
        return fail(ctx, CRGPU_E_ALIGN, "gapopen=%g gapextend=%g: need gapopen >= 4 (>= -min(EDNAFULL)) and "
                                        "0 <= gapextend <= gapopen", gapopen, gapextend);
    std::vector<int> acode(La);
    std::string amp_up(La, 'N');
    for (int i = 0; i < La; ++i) {
        acode[i] = host_code(amplicon[i]);
        if (acode[i] < 0) return fail(ctx, CRGPU_E_ALIGN, "amplicon has a base outside ACGTN at %d", i);
        amp_up[i] = amplicon[i];
    }
    int G, K;
    if (!choose_tile(La, &G, &K)) return fail(ctx, CRGPU_E_ALIGN, "no kernel tile for amplicon length %d", La);
    const int GK = G * K, P = GK - La;

    // ---- pair reads of equal length (host) ----
    int maxlen = 0;
    std::vector<int32_t> order((size_t)nsub);
    {
        std::vector<int64_t> cnt(CRGPU_MAX_READ + 2, 0);
        for (int64_t i = 0; i < nsub; ++i) {
            const int64_t r = subset ? subset[i] : i;
            const int64_t len = h_offsets[r + 1] - h_offsets[r];
            if (len < CRGPU_MIN_LEN || len > CRGPU_MAX_READ)
                return fail(ctx, CRGPU_E_ALIGN, "read %lld has length %lld outside [%d, %d]", (long long)r, (long long)len,
                            CRGPU_MIN_LEN, CRGPU_MAX_READ);
            cnt[len + 1]++;
            if (len > maxlen) maxlen = (int)len;
        }
        for (int l = 1; l <= CRGPU_MAX_READ + 1; ++l) cnt[l] += cnt[l - 1];
        for (int64_t i = 0; i < nsub; ++i) {
            const int64_t r = subset ? subset[i] : i;
            const int64_t len = h_offsets[r + 1] - h_offsets[r];
            order[(size_t)cnt[len]++] = (int32_t)r;
        }
    }
    if ((int64_t)scale * 5 * std::min(La, maxlen) + 64 >= MAX_ABS_SCORE ||
        (int64_t)2 * open_s + (int64_t)ext_s * (La + maxlen) + 8 * scale + 64 >= MAX_ABS_SCORE)
        return fail(ctx, CRGPU_E_ALIGN, "scores would leave the exact int16 range (scale %d, lengths %d/%d)", scale, La, maxlen);
    if (slot < (int64_t)La + maxlen && d_ref) return fail(ctx, CRGPU_E_ARG, "slot %lld < amplicon + longest read %d", (long long)slot, La + maxlen);

    std::vector<int32_t> pair_lo, pair_hi, plen;
    std::vector<int64_t> pc_off(1, 0);
    pair_lo.reserve((size_t)nsub / 2 + 1); pair_hi.reserve((size_t)nsub / 2 + 1); plen.reserve((size_t)nsub / 2 + 1);
    pc_off.reserve((size_t)nsub / 2 + 2);
    int64_t cells = 0;
    for (int64_t i = 0; i < nsub;) {
        const int32_t r0 = order[(size_t)i];
        const int len0 = (int)(h_offsets[r0 + 1] - h_offsets[r0]);
        int32_t r1 = r0;
        int64_t step = 1;
        if (i + 1 < nsub) {
            const int32_t c = order[(size_t)i + 1];
            if ((int)(h_offsets[c + 1] - h_offsets[c]) == len0) { r1 = c; step = 2; }
        }
        pair_lo.push_back(r0); pair_hi.push_back(r1); plen.push_back(len0);
        pc_off.push_back(pc_off.back() + len0);
        cells += step * (int64_t)La * len0;
        i += step;
    }
    if (n_cells) *n_cells += cells;
    const int np = (int)plen.size();
    if (np == 0) return CRGPU_OK;

    // ---- batches bounded by the traceback budget ----
    std::vector<int64_t> tb_off((size_t)np);
    std::vector<int> batch_start(1, 0);
    {
        const int64_t words_per_col = GK / 2;
        int64_t acc = 0;
        const int64_t budget_words = (int64_t)(ctx->tb_budget / 4);
        for (int p = 0; p < np; ++p) {
            const int64_t w = (int64_t)plen[p] * words_per_col;
            if (acc > 0 && acc + w > budget_words) { batch_start.push_back(p); acc = 0; }
            tb_off[(size_t)p] = acc;
            acc += w;
        }
        batch_start.push_back(np);
    }
    int64_t max_tb_words = 0, max_lr = 0; int max_bp = 0;
    for (size_t b = 0; b + 1 < batch_start.size(); ++b) {
        const int p0 = batch_start[b], p1 = batch_start[b + 1];
        const int64_t w = tb_off[(size_t)p1 - 1] + (int64_t)plen[(size_t)p1 - 1] * (GK / 2);
        max_tb_words = std::max(max_tb_words, w);
        max_lr = std::max(max_lr, pc_off[(size_t)p1] - pc_off[(size_t)p0]);
        max_bp = std::max(max_bp, p1 - p0);
    }

    // ---- profile table ----
    const int PS = prof_stride(G, K), SS = strip_stride(K);
    std::vector<int32_t> prof((size_t)NPAIR * PS, 0);
    for (int cp = 0; cp < NPAIR; ++cp) {
        const int lo = cp % NCODE, hi = cp / NCODE;
        for (int r = P; r < GK; ++r) {
            const int a = acode[r - P];
            const int32_t slo = scale * host_ednafull(a, lo), shi = scale * host_ednafull(a, hi);
            prof[(size_t)cp * PS + (r / K) * SS + (r % K)] = shi * 65536 + slo;
        }
    }

    // ---- uploads ----
    cudaStream_t s = ctx->stream;
    CK(ctx->amp.reserve((size_t)La));
    CK(ctx->prof.reserve(prof.size() * 4));
    CK(ctx->pc.reserve((size_t)pc_off.back()));
    CK(ctx->pc_off.reserve(pc_off.size() * 8));
    CK(ctx->plen.reserve((size_t)np * 4));
    CK(ctx->pair_lo.reserve((size_t)np * 4));
    CK(ctx->pair_hi.reserve((size_t)np * 4));
    CK(ctx->tb_off.reserve((size_t)np * 8));
    CK(ctx->tb.reserve((size_t)max_tb_words * 4));
    CK(ctx->lastrow.reserve((size_t)max_lr * 12));
    CK(ctx->lastcol.reserve((size_t)max_bp * GK * 12));
    CK(ctx->errflag.reserve(4));
    CK(cudaMemcpyAsync(ctx->amp.p, amp_up.data(), (size_t)La, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->prof.p, prof.data(), prof.size() * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->pc_off.p, pc_off.data(), pc_off.size() * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->plen.p, plen.data(), (size_t)np * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->pair_lo.p, pair_lo.data(), (size_t)np * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->pair_hi.p, pair_hi.data(), (size_t)np * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->tb_off.p, tb_off.data(), (size_t)np * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(ctx->errflag.p, 0, 4, s));

    span_begin(ctx, T_ENCODE);
    CK(launch_encode(d_reads, d_offsets, ctx->pair_lo.as<int32_t>(), ctx->pair_hi.as<int32_t>(), ctx->pc_off.as<int64_t>(), np,
                     ctx->pc.as<uint8_t>(), ctx->errflag.as<int>(), ctx->num_sms, s));
    span_end(ctx);
    int h_err = 0;
    CK(cudaMemcpyAsync(&h_err, ctx->errflag.p, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));   // also keeps the host vectors above alive until the copies are done
    if (h_err) return fail(ctx, CRGPU_E_ALIGN, "a read contains a base outside ACGTN(U)");

    for (size_t b = 0; b + 1 < batch_start.size(); ++b) {
        FillArgs fa;
        fa.prof = ctx->prof.as<int32_t>();
        fa.pc = ctx->pc.as<uint8_t>();
        fa.pc_off = ctx->pc_off.as<int64_t>();
        fa.plen = ctx->plen.as<int32_t>();
        fa.tb_off = ctx->tb_off.as<int64_t>();
        fa.tb = ctx->tb.as<uint32_t>();
        fa.lastrow = ctx->lastrow.as<uint32_t>();
        fa.lastcol = ctx->lastcol.as<uint32_t>();
        fa.p0 = batch_start[b]; fa.p1 = batch_start[b + 1];
        fa.open = open_s; fa.ext = ext_s;
        span_begin(ctx, T_FILL);
        CK(launch_fill(G, K, fa, ctx->num_sms, s));
        span_end(ctx);

        WalkArgs wa;
        wa.tb = fa.tb; wa.tb_off = fa.tb_off; wa.lastrow = fa.lastrow; wa.lastcol = fa.lastcol;
        wa.pc_off = fa.pc_off; wa.plen = fa.plen;
        wa.pair_lo = ctx->pair_lo.as<int32_t>(); wa.pair_hi = ctx->pair_hi.as<int32_t>();
        wa.reads = d_reads; wa.offsets = d_offsets; wa.amplicon = ctx->amp.as<uint8_t>();
        wa.La = La; wa.GK = GK; wa.P = P; wa.p0 = fa.p0; wa.p1 = fa.p1;
        wa.open = open_s; wa.ext = ext_s; wa.scale = scale;
        wa.recs = d_recs; wa.ref_out = d_ref; wa.mark_out = d_mark; wa.qry_out = d_qry; wa.slot = slot;
        wa.out_index = d_out_index; wa.rc_out = rc_out;
        span_begin(ctx, T_WALK);
        CK(launch_walk(wa, s));
        span_end(ctx);
    }
    return CRGPU_OK;
}

}  // namespace crgpu

extern "C" {

int crgpu_align(crgpu_ctx *ctx, int mem, const char *amplicon, int amplicon_len,
                const uint8_t *reads, const int64_t *offsets, int64_t n, double gapopen, double gapextend,
                crgpu_aln_rec *recs, uint8_t *ref_out, uint8_t *mark_out, uint8_t *qry_out, int64_t slot)
{
    if (!ctx) return CRGPU_E_ARG;
    if (!amplicon || n < 0 || (n > 0 && (!reads || !offsets || !recs)))
        return fail(ctx, CRGPU_E_ARG, "crgpu_align: null pointer or negative count");
    const bool want = ref_out || mark_out || qry_out;
    if (want && !(ref_out && mark_out && qry_out)) return fail(ctx, CRGPU_E_ARG, "crgpu_align: pass all three row buffers or none");
    if (mem != CRGPU_MEM_HOST && mem != CRGPU_MEM_DEVICE) return fail(ctx, CRGPU_E_ARG, "bad mem");
    if (n >= (int64_t)1 << 31) return fail(ctx, CRGPU_E_ARG, "too many reads in one call");
    timing_reset(ctx);
    if (n == 0) return CRGPU_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;

    std::vector<int64_t> h_off_copy;
    const int64_t *h_off = offsets;
    const uint8_t *d_reads = reads; const int64_t *d_off = offsets;
    crgpu_aln_rec *d_recs = recs; uint8_t *d_ref = ref_out, *d_mark = mark_out, *d_qry = qry_out;
    if (mem == CRGPU_MEM_DEVICE) {
        h_off_copy.resize((size_t)n + 1);
        CK(cudaMemcpyAsync(h_off_copy.data(), offsets, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        h_off = h_off_copy.data();
    } else {
        const int64_t total = offsets[n];
        CK(ctx->reads.reserve((size_t)std::max<int64_t>(total, 1)));
        CK(ctx->offsets.reserve((size_t)(n + 1) * 8));
        CK(ctx->recs.reserve((size_t)n * sizeof(crgpu_aln_rec)));
        CK(cudaMemcpyAsync(ctx->reads.p, reads, (size_t)total, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->offsets.p, offsets, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, s));
        d_reads = ctx->reads.as<uint8_t>(); d_off = ctx->offsets.as<int64_t>(); d_recs = ctx->recs.as<crgpu_aln_rec>();
        if (want) {
            CK(ctx->sref.reserve((size_t)n * slot));
            CK(ctx->smark.reserve((size_t)n * slot));
            CK(ctx->sqry.reserve((size_t)n * slot));
            d_ref = ctx->sref.as<uint8_t>(); d_mark = ctx->smark.as<uint8_t>(); d_qry = ctx->sqry.as<uint8_t>();
        }
    }
    int rc = align_core(ctx, amplicon, amplicon_len, d_reads, d_off, h_off, nullptr, n, nullptr, 0, gapopen, gapextend, d_recs,
                        d_ref, d_mark, d_qry, slot, nullptr);
    if (rc != CRGPU_OK) { cudaStreamSynchronize(s); return rc; }
    if (mem == CRGPU_MEM_HOST) {
        CK(cudaMemcpyAsync(recs, d_recs, (size_t)n * sizeof(crgpu_aln_rec), cudaMemcpyDeviceToHost, s));
        if (want) {
            CK(cudaMemcpyAsync(ref_out, d_ref, (size_t)n * slot, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(mark_out, d_mark, (size_t)n * slot, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(qry_out, d_qry, (size_t)n * slot, cudaMemcpyDeviceToHost, s));
        }
    }
    CK(cudaStreamSynchronize(s));
    timing_collect(ctx);
    return CRGPU_OK;
}

int crgpu_int_peak(crgpu_ctx *ctx, int which, double *lane_ops_per_s)
{
    if (!ctx || !lane_ops_per_s) return CRGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->aux[7].reserve(4096));
    const int iters = 4096;
    double ops = 0;
    // warm-up, then best of 5
    CK(launch_int_peak(which, ctx->num_sms, 64, ctx->aux[7].as<unsigned>(), ctx->stream, &ops));
    CK(cudaStreamSynchronize(ctx->stream));
    double best = 0;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(a, ctx->stream));
        CK(launch_int_peak(which, ctx->num_sms, iters, ctx->aux[7].as<unsigned>(), ctx->stream, &ops));
        CK(cudaEventRecord(b, ctx->stream));
        CK(cudaEventSynchronize(b));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (ms > 0) best = std::max(best, ops / (ms * 1e-3));
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    *lane_ops_per_s = best;
    return CRGPU_OK;
}

}  // extern "C"
