// crgpu_api.cu -- the C ABI declared in include/crgpu.h: context, S1, S2, batching, host-side pairing.
#include "crgpu_internal.h"
#include <cstdlib>

using namespace crgpu;

extern "C" {

int crgpu_abi_version(void) { return CRGPU_ABI_VERSION; }

int crgpu_create(crgpu_ctx **out, int device)
{
    if (!out) return CRGPU_E_ARG;
    *out = nullptr;
    int ndev = 0, prev = -1;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return CRGPU_E_CUDA;
    cudaGetDevice(&prev);
    if (cudaSetDevice(device) != cudaSuccess) return CRGPU_E_CUDA;
    crgpu_ctx *c = new crgpu_ctx();
    c->device = device;
    auto bail = [&]() {
        if (c->stream) cudaStreamDestroy(c->stream);
        if (c->stream2) cudaStreamDestroy(c->stream2);
        if (c->stream3) cudaStreamDestroy(c->stream3);
        if (c->stream4) cudaStreamDestroy(c->stream4);
        for (int i = 0; i < 2; ++i) if (c->walk_side[i]) cudaEventDestroy(c->walk_side[i]);
        if (c->stream_copy) cudaStreamDestroy(c->stream_copy);
        for (int i = 0; i < 2; ++i) { if (c->staged_ev[i]) cudaEventDestroy(c->staged_ev[i]); if (c->out_ev[i]) cudaEventDestroy(c->out_ev[i]); if (c->out_ready[i]) cudaEventDestroy(c->out_ready[i]); }
        if (c->ready) cudaEventDestroy(c->ready);
        if (c->mbox_h) cudaFreeHost(c->mbox_h);
        for (int i = 0; i < 2; ++i) { if (c->fill_done[i]) cudaEventDestroy(c->fill_done[i]); if (c->walk_done[i]) cudaEventDestroy(c->walk_done[i]); }
        delete c;
        if (prev >= 0) cudaSetDevice(prev);
        return CRGPU_E_CUDA;
    };
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail();
    if (prop.major != 10) return bail();                       // sm_100a SASS only: no other target, no fallback
    c->num_sms = prop.multiProcessorCount;
    // scratch of ONE traceback batch (there are two sets): 24 GiB holds half of a 2^20-read call (band flags + boundary rows
    // of 2^18 read pairs vs a 250-bp amplicon + HDR amplicon: ~15 GB), bounded by an eighth of the device's memory
    c->tb_budget = std::min<size_t>((size_t)24 << 30, prop.totalGlobalMem / 8);
    // The traceback walk reads one byte per visited cell from scattered sectors: CRGPU_L2_HINT=1 asks L2 not to over-fetch
    // neighbouring sectors from HBM.  A device-wide limit (it measured neutral, profiles/r01_notes.md), so it is opt-in
    // and crgpu_destroy puts the previous value back.
    c->trace_on = getenv("CRGPU_TRACE") != nullptr;
    if (getenv("CRGPU_L2_HINT") && cudaDeviceGetLimit(&c->l2_gran_prev, cudaLimitMaxL2FetchGranularity) == cudaSuccess) {
        if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32) == cudaSuccess) c->l2_gran_set = true;
        else cudaGetLastError();
    }
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return bail();
    if (cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking) != cudaSuccess) return bail();
    if (cudaStreamCreateWithFlags(&c->stream3, cudaStreamNonBlocking) != cudaSuccess) return bail();
    if (cudaStreamCreateWithFlags(&c->stream4, cudaStreamNonBlocking) != cudaSuccess) return bail();
    for (int i = 0; i < 2; ++i) if (cudaEventCreateWithFlags(&c->walk_side[i], cudaEventDisableTiming) != cudaSuccess) return bail();
    if (cudaStreamCreateWithFlags(&c->stream_copy, cudaStreamNonBlocking) != cudaSuccess) return bail();
    if (cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming) != cudaSuccess) return bail();
    for (int i = 0; i < 2; ++i) {
        if (cudaEventCreateWithFlags(&c->fill_done[i], cudaEventDisableTiming) != cudaSuccess) return bail();
        if (cudaEventCreateWithFlags(&c->walk_done[i], cudaEventDisableTiming) != cudaSuccess) return bail();
        if (cudaEventCreateWithFlags(&c->staged_ev[i], cudaEventDisableTiming) != cudaSuccess) return bail();
        if (cudaEventCreateWithFlags(&c->out_ev[i], cudaEventDisableTiming) != cudaSuccess) return bail();
        if (cudaEventCreateWithFlags(&c->out_ready[i], cudaEventDisableTiming) != cudaSuccess) return bail();
    }
    c->mbox_bytes = (size_t)4 << 20;
    if (cudaHostAlloc(reinterpret_cast<void **>(&c->mbox_h), c->mbox_bytes, cudaHostAllocMapped) != cudaSuccess) { c->mbox_h = nullptr; return bail(); }
    if (cudaHostGetDevicePointer(reinterpret_cast<void **>(&c->mbox_d), c->mbox_h, 0) != cudaSuccess) return bail();
    if (prev >= 0 && prev != device) cudaSetDevice(prev);      // the caller's current device is not ours to change
    *out = c;
    return CRGPU_OK;
}

void crgpu_destroy(crgpu_ctx *c)
{
    if (!c) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
    if (c->l2_gran_set) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, c->l2_gran_prev);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->stream2);
    cudaStreamSynchronize(c->stream3);
    cudaStreamSynchronize(c->stream4);
    cudaStreamSynchronize(c->stream_copy);
    DBuf *all[] = {&c->reads, &c->offsets, &c->amp, &c->prof, &c->pc, &c->pc_off, &c->plen, &c->pair_lo, &c->pair_hi,
                   &c->order, &c->plan_hist, &c->plan_tab, &c->tb2, &c->lastrow2, &c->lastcol2, &c->tb, &c->lastrow, &c->lastcol, &c->errflag, &c->recs, &c->sref, &c->smark, &c->sqry, &c->ops, &c->ops_rc, &c->alleles, &c->prof_h, &c->amp_h, &c->tbh, &c->tbh2,
                   &c->top, &c->top2, &c->lastrow_h, &c->lastrow_h2, &c->lastcol_h, &c->lastcol_h2,
                   &c->prof_s, &c->prof_hs, &c->join, &c->btops[0], &c->btops[1], &c->bleft[0], &c->bleft[1], &c->btops_h[0], &c->btops_h[1], &c->bleft_h[0], &c->bleft_h[1], &c->escaped,
                   &c->joinb, &c->fastflags, &c->need[0], &c->need[1], &c->plist[0], &c->plist[1], &c->plist2[0], &c->plist2[1], &c->selscratch[0], &c->selscratch[1], &c->need_cnt,
                   &c->rowvals[0], &c->rowvals[1], &c->rowvals_h[0], &c->rowvals_h[1], &c->badbase,
                   &c->need_read[0], &c->need_read[1], &c->rlist[0], &c->rlist[1],
                   &c->exact_go, &c->exact_sel, &c->stage_reads[0], &c->stage_reads[1], &c->stage_off[0], &c->stage_off[1], &c->stage_pack[0], &c->stage_pack[1]};
    for (DBuf *b : all) b->release();
    for (auto &b : c->q_in) b.release();
    for (auto &b : c->q_out) b.release();
    for (auto &b : c->aux) b.release();
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) { cudaEventDestroy(c->fill_done[i]); cudaEventDestroy(c->walk_done[i]); }
    cudaEventDestroy(c->ready);
    for (int i = 0; i < 2; ++i) { cudaEventDestroy(c->staged_ev[i]); cudaEventDestroy(c->out_ev[i]); cudaEventDestroy(c->out_ready[i]); }
    for (auto &slot : c->stage_out) for (auto &b : slot) b.release();
    if (c->mbox_h) cudaFreeHost(c->mbox_h);
    cudaStreamDestroy(c->stream_copy);
    cudaStreamDestroy(c->stream4);
    for (int i = 0; i < 2; ++i) { cudaEventDestroy(c->walk_side[i]); c->rlist_side[i].release(); }
    cudaStreamDestroy(c->stream3);
    cudaStreamDestroy(c->stream2);
    cudaStreamDestroy(c->stream);
    const int dev = c->device;
    delete c;
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
}

const char *crgpu_last_error(const crgpu_ctx *c) { return c ? c->err.c_str() : "null context"; }

int crgpu_set_traceback_budget(crgpu_ctx *c, size_t bytes)
{
    if (!c || bytes < ((size_t)16 << 20)) return CRGPU_E_ARG;
    c->tb_budget = bytes;
    return CRGPU_OK;
}

int crgpu_set_exact_shortcut(crgpu_ctx *c, int on)
{
    if (!c) return CRGPU_E_ARG;
    c->exact_shortcut = on != 0;
    return CRGPU_OK;
}

int64_t crgpu_last_exact(const crgpu_ctx *c) { return c ? c->n_exact : -1; }

int crgpu_set_deferred_outputs(crgpu_ctx *c, int on)
{
    if (!c) return CRGPU_E_ARG;
    c->deferred_out = on != 0;
    return CRGPU_OK;
}

int crgpu_set_overlap(crgpu_ctx *c, int on)
{
    if (!c) return CRGPU_E_ARG;
    c->overlap = on != 0;
    return CRGPU_OK;
}

int crgpu_set_share_prefix(crgpu_ctx *c, int on)
{
    if (!c) return CRGPU_E_ARG;
    c->share_prefix = on != 0;
    return CRGPU_OK;
}

int crgpu_set_band(crgpu_ctx *c, int half_width)
{
    if (!c || half_width < 0 || half_width > 512) return CRGPU_E_ARG;
    c->band_B = half_width;
    c->band_holdoff = 0;
    return CRGPU_OK;
}

int crgpu_get_band(const crgpu_ctx *c) { return c ? c->band_B : -1; }

int crgpu_set_diag_shortcut(crgpu_ctx *c, int on)
{
    if (!c) return CRGPU_E_ARG;
    c->diag = on != 0;
    return CRGPU_OK;
}

int crgpu_last_diag(const crgpu_ctx *c, int64_t out[2])
{
    if (!c || !out) return CRGPU_E_ARG;
    out[0] = c->n_diag_pairs[0]; out[1] = c->n_diag_pairs[1];
    return CRGPU_OK;
}

int crgpu_last_escaped(const crgpu_ctx *c, int out[2])
{
    if (!c || !out) return CRGPU_E_ARG;
    out[0] = c->n_escaped[0]; out[1] = c->n_escaped[1];
    return CRGPU_OK;
}

int crgpu_last_timing(const crgpu_ctx *c, float out_ms[6], int64_t out_launches[6])
{
    if (!c) return CRGPU_E_ARG;
    for (int i = 0; i <= T_OTHER; ++i) {
        if (out_ms) out_ms[i] = c->ms[i];
        if (out_launches) out_launches[i] = c->launches[i];
    }
    if (out_ms) out_ms[T_FILL] += c->ms[T_SCORE] + c->ms[T_BAND];
    if (out_launches) out_launches[T_FILL] += c->launches[T_SCORE] + c->launches[T_BAND];
    return CRGPU_OK;
}

int crgpu_last_fill_breakdown(const crgpu_ctx *c, double out_ms[3], int64_t out_launches[3], int64_t out_cells[3])
{
    if (!c) return CRGPU_E_ARG;
    const int fam[3] = {T_FILL, T_SCORE, T_BAND};
    for (int i = 0; i < 3; ++i) {
        if (out_ms) out_ms[i] = c->ms[fam[i]];
        if (out_launches) out_launches[i] = c->launches[fam[i]];
        if (out_cells) out_cells[i] = c->cells_kind[i];
    }
    return CRGPU_OK;
}

static int sync_impl(crgpu_ctx *ctx)
{
    if (!ctx) return CRGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    { const int rc = flush_stages(ctx); if (rc) return rc; }
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream_copy));          // staged inputs / deferred outputs
    return CRGPU_OK;
}

void *crgpu_stream(crgpu_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

static int qualfilter_impl(crgpu_ctx *ctx, int mem, const uint8_t *qual, const int64_t *offsets, int64_t n,
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
// the reference upper-cases the amplicon before anything else (CORE:1288); the walkers copy these bytes into the text rows
static char host_upper(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }

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
// ---- mailbox transfers (crgpu_internal.h) -------------------------------------------------------------------------------
__global__ void k_copy_small(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, size_t bytes)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    if ((((uintptr_t)dst | (uintptr_t)src) & 3) == 0) {
        const size_t w = bytes >> 2;
        for (size_t i = tid; i < w; i += nth) reinterpret_cast<uint32_t *>(dst)[i] = reinterpret_cast<const uint32_t *>(src)[i];
        for (size_t i = (w << 2) + tid; i < bytes; i += nth) dst[i] = src[i];
    } else {
        for (size_t i = tid; i < bytes; i += nth) dst[i] = src[i];
    }
}

static cudaError_t copy_small(uint8_t *dst, const uint8_t *src, size_t bytes, cudaStream_t s)
{
    const unsigned blocks = (unsigned)std::min<size_t>(64, (bytes / 4 + 255) / 256 + 1);
    k_copy_small<<<blocks, 256, 0, s>>>(dst, src, bytes);
    return cudaGetLastError();
}

cudaError_t fetch_small(crgpu_ctx *ctx, void *h_dst, const void *d_src, size_t bytes, cudaStream_t s)
{
    if (bytes == 0) return cudaSuccess;
    const size_t off = (ctx->mbox_used + 15) & ~(size_t)15;
    if (!ctx->mbox_h || off + bytes > ctx->mbox_bytes) return cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, s);
    ctx->mbox_used = off + bytes;
    ctx->mbox_pending.push_back({h_dst, off, bytes});
    ctx->launches[T_OTHER] += 1;                                  // (k_copy_small: counted, not timed)
    return copy_small(ctx->mbox_d + off, reinterpret_cast<const uint8_t *>(d_src), bytes, s);
}

cudaError_t fetch_wait(crgpu_ctx *ctx, cudaStream_t s)
{
    const cudaError_t e = cudaStreamSynchronize(s);
    if (e == cudaSuccess)
        for (const auto &f : ctx->mbox_pending) memcpy(f.dst, ctx->mbox_h + f.off, f.bytes);
    ctx->mbox_pending.clear();
    ctx->mbox_used = 0;
    return e;
}

cudaError_t push_small(crgpu_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, cudaStream_t s)
{
    if (bytes == 0) return cudaSuccess;
    const size_t off = (ctx->mbox_used + 15) & ~(size_t)15;
    if (!ctx->mbox_h || off + bytes > ctx->mbox_bytes) return cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, s);
    ctx->mbox_used = off + bytes;
    memcpy(ctx->mbox_h + off, h_src, bytes);
    ctx->launches[T_OTHER] += 1;
    return copy_small(reinterpret_cast<uint8_t *>(d_dst), ctx->mbox_d + off, bytes, s);
}



// ---------------------------------------------------------------------------------------------
// build_plan: bucket the reads (all n, or `d_subset`) by length on the device, pair consecutive
// reads of a bucket, encode the pair codes.  The plan depends only on the reads, so the amplicon
// pass and the HDR-amplicon pass share it.  Only the length histogram (8 KB) visits the host.
// ---------------------------------------------------------------------------------------------
int build_plan(crgpu_ctx *ctx, const uint8_t *d_reads, const int64_t *d_offsets, const int32_t *d_subset, int64_t nsub)
{
    PairPlan &pl = ctx->plan;
    pl = PairPlan();
    pl.nsub = nsub;
    if (nsub <= 0) return CRGPU_OK;
    cudaStream_t s = ctx->stream;
    const int NB = CRGPU_MAX_READ + 1;
    CK(ctx->plan_hist.reserve((size_t)NB * 4 * 2 + 16));          // hist + cursor + err
    int *d_hist = ctx->plan_hist.as<int>(), *d_cursor = d_hist + NB, *d_err = d_cursor + NB;
    CK(cudaMemsetAsync(d_hist, 0, (size_t)NB * 4 * 2 + 16, s));
    span_begin(ctx, T_ENCODE);
    CK(launch_len_hist(d_offsets, d_subset, nsub, CRGPU_MIN_LEN, CRGPU_MAX_READ, d_hist, d_err, s));
    span_end(ctx);
    std::vector<int> hist((size_t)NB + 0);
    int h_err = 0;
    CK(fetch_small(ctx, hist.data(), d_hist, (size_t)NB * 4, s));
    CK(fetch_small(ctx, &h_err, d_err, 4, s));
    CK(fetch_wait(ctx, s));
    if (h_err & 2)
        return fail(ctx, CRGPU_E_ALIGN, "a read has a length outside [%d, %d]", CRGPU_MIN_LEN, CRGPU_MAX_READ);
    std::vector<int64_t> read_start((size_t)NB, 0);
    int64_t rs = 0, ps = 0, pcs = 0;
    for (int l = 0; l < NB; ++l) {
        read_start[(size_t)l] = rs;
        const int c = hist[(size_t)l];
        if (c > 0) {
            HostSeg sg;
            sg.len = l; sg.cnt = c; sg.read_start = rs; sg.pair_start = ps; sg.pc_start = pcs;
            pl.segs.push_back(sg);
            const int64_t npair = ((int64_t)c + 1) / 2;
            rs += c; ps += npair; pcs += npair * l;
            pl.maxlen = l;
            if (pl.minlen == 0) pl.minlen = l;
            pl.sum_len += (int64_t)c * l;
        }
    }
    if (ps >= ((int64_t)1 << 31)) return fail(ctx, CRGPU_E_ARG, "too many read pairs");
    pl.np = (int)ps;
    pl.total_pc = pcs;
    const int nseg = (int)pl.segs.size();
    CK(ctx->plan_tab.reserve((size_t)NB * 8 + (size_t)nseg * sizeof(HostSeg)));
    int64_t *d_read_start = ctx->plan_tab.as<int64_t>();
    HostSeg *d_segs = reinterpret_cast<HostSeg *>(d_read_start + NB);
    CK(push_small(ctx, d_read_start, read_start.data(), (size_t)NB * 8, s));
    CK(push_small(ctx, d_segs, pl.segs.data(), (size_t)nseg * sizeof(HostSeg), s));
    CK(ctx->order.reserve((size_t)nsub * 4));
    CK(ctx->pc.reserve((size_t)std::max<int64_t>(pcs, 1) + 16));      // (+16: k_gotoh_score2 prefetches one code past an odd-length read)
    CK(ctx->pc_off.reserve((size_t)(pl.np + 1) * 8));
    CK(ctx->plen.reserve((size_t)pl.np * 4));
    CK(ctx->pair_lo.reserve((size_t)pl.np * 4));
    CK(ctx->pair_hi.reserve((size_t)pl.np * 4));
    span_begin(ctx, T_ENCODE);
    CK(launch_scatter_order(d_offsets, d_subset, nsub, d_read_start, d_cursor, ctx->order.as<int32_t>(), s));
    CK(launch_build_pairs(d_segs, nseg, pl.np, ctx->order.as<int32_t>(), ctx->pair_lo.as<int32_t>(), ctx->pair_hi.as<int32_t>(),
                          ctx->plen.as<int32_t>(), ctx->pc_off.as<int64_t>(), pcs, s));
    CK(launch_encode(d_reads, d_offsets, ctx->pair_lo.as<int32_t>(), ctx->pair_hi.as<int32_t>(), ctx->pc_off.as<int64_t>(), pl.np,
                     ctx->pc.as<uint8_t>(), d_err, ctx->d_bad, ctx->num_sms, s));
    span_end(ctx, 3);
    if (ctx->d_bad) return CRGPU_OK;            // bases outside ACGTN(U) are flagged per read: nothing to wait for
    CK(fetch_small(ctx, &h_err, d_err, 4, s));
    CK(fetch_wait(ctx, s));
    if (h_err & 1) return fail(ctx, CRGPU_E_ALIGN, "a read contains a base outside ACGTN(U)");
    return CRGPU_OK;
}

// pc_off of pair p, from the per-length segments (host)
static int64_t plan_pc_off(const PairPlan &pl, int64_t p)
{
    if (p >= pl.np) return pl.total_pc;
    size_t lo = 0, hi = pl.segs.size() - 1;
    while (lo < hi) {
        const size_t mid = (lo + hi + 1) / 2;
        if (pl.segs[mid].pair_start <= p) lo = mid; else hi = mid - 1;
    }
    return pl.segs[lo].pc_start + (p - pl.segs[lo].pair_start) * pl.segs[lo].len;
}

// ---------------------------------------------------------------------------------------------
// run_plan: align every pair of the current plan to `amplicon`.  recs / rows are indexed by read
// (or by d_out_index[read]).  Batches are bounded by the traceback budget.
// ---------------------------------------------------------------------------------------------
int run_plan(crgpu_ctx *ctx, const char *amplicon, int La, const uint8_t *d_reads, const int64_t *d_offsets,
             const int32_t *d_out_index, int rc_out, double gapopen, double gapextend, crgpu_aln_rec *d_recs,
             uint8_t *d_ref, uint8_t *d_mark, uint8_t *d_qry, int64_t slot, int64_t *n_cells, uint32_t *d_ops,
             int64_t ops_stride, int lane)
{
    const PairPlan &pl = ctx->plan;
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
        amp_up[i] = host_upper(amplicon[i]);
    }
    if (pl.np == 0) return CRGPU_OK;
    int G, K;
    if (!choose_tile(La, &G, &K)) return fail(ctx, CRGPU_E_ALIGN, "no kernel tile for amplicon length %d", La);
    // A plan of a few thousand pairs (band escapes, RC rescue) is a single wave whatever the tile: it lasts as long as one
    // pair's systolic sweep, (read length + G) steps of K rows.  Take the flattest compiled tile that covers the amplicon.
    if (pl.np <= 4096 && !getenv("CRGPU_TILE")) {
        static const int flat[][2] = {{16, 16}, {16, 24}, {32, 24}};
        for (const auto &t : flat)
            if (t[1] < K && t[0] * t[1] >= La && tile_available(t[0], t[1])) { G = t[0]; K = t[1]; break; }
    }
    const int GK = G * K, P = GK - La;
    const int maxlen = pl.maxlen;
    if ((int64_t)scale * 5 * std::min(La, maxlen) + 64 >= MAX_ABS_SCORE ||
        (int64_t)2 * open_s + (int64_t)ext_s * (La + maxlen) + 8 * scale + 64 >= MAX_ABS_SCORE)
        return fail(ctx, CRGPU_E_ALIGN, "scores would leave the exact int16 range (scale %d, lengths %d/%d)", scale, La, maxlen);
    if (slot < (int64_t)La + maxlen && d_ref) return fail(ctx, CRGPU_E_ARG, "slot %lld < amplicon + longest read %d", (long long)slot, La + maxlen);
    if (n_cells) *n_cells += (int64_t)La * pl.sum_len;
    ctx->cells_kind[0] += (int64_t)La * pl.sum_len;

    // ---- batches bounded by the traceback budget (pairs are ordered by length) ----
    std::vector<int> batch_start(1, 0);
    int64_t max_tb_words = 0, max_lr = 0;
    int max_bp = 0;
    {
        const int64_t words_per_col = GK / 2;
        const int64_t budget_words = (int64_t)(ctx->tb_budget / 4);
        int64_t acc = 0, acc_lr = 0;
        int p_begin = 0;
        auto close = [&](int p_end) {
            max_tb_words = std::max(max_tb_words, acc);
            max_lr = std::max(max_lr, acc_lr);
            max_bp = std::max(max_bp, p_end - p_begin);
            batch_start.push_back(p_end);
            p_begin = p_end; acc = 0; acc_lr = 0;
        };
        for (const HostSeg &sg : pl.segs) {
            const int64_t w = (int64_t)sg.len * words_per_col;
            int64_t left = ((int64_t)sg.cnt + 1) / 2;
            int64_t p = sg.pair_start;
            while (left > 0) {
                int64_t fit = (budget_words - acc) / w;
                if (fit <= 0) {
                    if (acc > 0) { close((int)p); continue; }
                    fit = 1;                                   // a single pair always runs
                }
                const int64_t take = std::min(fit, left);
                acc += take * w; acc_lr += take * sg.len; p += take; left -= take;
                if (left > 0) close((int)p);
            }
        }
        if (acc > 0) close(pl.np);
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
    cudaStream_t s = ctx->stream;
    // lane 1 (a second, small pass that runs BESIDE a lane-0 pass over the same plan: the HDR re-alignment of the reads that
    // left the band): its own tables, the second scratch set, the third stream; the caller joins ctx->walk_done[1]
    DBuf &ampb = lane ? ctx->amp_h : ctx->amp, &profb = lane ? ctx->prof_h : ctx->prof;
    CK(ampb.reserve((size_t)La));
    CK(profb.reserve(prof.size() * 4));
    const bool two = !lane && batch_start.size() > 2 && ctx->overlap && !getenv("CRGPU_NO_OVERLAP");     // double-buffer only when there is a second batch to overlap with
    if (!lane) {
        CK(ctx->tb.reserve((size_t)max_tb_words * 4));
        CK(ctx->lastrow.reserve((size_t)max_bp * 12));
        CK(ctx->lastcol.reserve((size_t)max_bp * G * 12));
    }
    if (two || lane) {
        CK(ctx->tb2.reserve((size_t)max_tb_words * 4));
        CK(ctx->lastrow2.reserve((size_t)max_bp * 12));
        CK(ctx->lastcol2.reserve((size_t)max_bp * G * 12));
    }
    CK(push_small(ctx, ampb.p, amp_up.data(), (size_t)La, s));          // (always on the main stream: the mailbox is its)
    CK(push_small(ctx, profb.p, prof.data(), prof.size() * 4, s));

    // Even fill batches run on the main stream, odd ones on a third stream, so that the persistent CTAs
    // of batch b+1 back-fill the SMs that the tail of batch b vacates (a fill launch is ~9 waves; its
    // last partial wave would otherwise leave most of the chip idle for one pair-time).  The traceback
    // walk of batch b runs on the second stream, overlapped with the fill of batch b+1 (the walk is
    // DRAM-latency bound, the fill integer-issue bound).  Two sets of traceback scratch alternate.
    cudaStream_t s2 = two ? ctx->stream2 : lane ? ctx->stream3 : ctx->stream;
    cudaStream_t sf[2] = {s, (two || lane) ? ctx->stream3 : s};
    if (two || lane) {
        CK(cudaEventRecord(ctx->ready, s));                   // plan, profile, amplicon are in place
        CK(cudaStreamWaitEvent(sf[1], ctx->ready, 0));
    }
    bool used[2] = {false, false};
    for (size_t b = 0; b + 1 < batch_start.size(); ++b) {
        const int cur = two ? (int)(b & 1) : lane;
        FillArgs fa;
        fa.prof = profb.as<int32_t>();
        fa.pc = ctx->pc.as<uint8_t>();
        fa.pc_off = ctx->pc_off.as<int64_t>();
        fa.plen = ctx->plen.as<int32_t>();
        fa.tb = (cur ? ctx->tb2 : ctx->tb).as<uint32_t>();
        fa.lastrow = (cur ? ctx->lastrow2 : ctx->lastrow).as<uint32_t>();
        fa.lastcol = (cur ? ctx->lastcol2 : ctx->lastcol).as<uint32_t>();
        fa.p0 = batch_start[b]; fa.p1 = batch_start[b + 1];
        fa.open = open_s; fa.ext = ext_s; fa.La = La; fa.one = 1;
        fa.top_out = nullptr; fa.top_out_lane = -1; fa.top_in = nullptr;
        fa.band_B = 0; fa.band_W = 0; fa.band_K = 0; fa.d_e = fa.d_eK = fa.d_eKb = fa.d_copen = 0; fa.band_row0 = 0; fa.band_tops = nullptr; fa.band_left = nullptr; fa.band_tb = nullptr;
        if (used[cur]) CK(cudaStreamWaitEvent(sf[cur], ctx->walk_done[cur], 0));     // scratch `cur` is free again
        span_begin(ctx, T_FILL, sf[cur]);
        CK(launch_fill(G, K, fa, ctx->num_sms, sf[cur]));
        span_end(ctx);
        CK(cudaEventRecord(ctx->fill_done[cur], sf[cur]));

        WalkArgs wa;
        wa.tb = fa.tb; wa.lastrow = fa.lastrow; wa.lastcol = fa.lastcol;
        wa.tb_upper = nullptr; wa.lastcol_upper = nullptr; wa.G_upper = 0; wa.split_row = 0;
        wa.join_row = 0; wa.join_out = nullptr; wa.join_in = nullptr;
        wa.band_B = 0; wa.band_W = 0; wa.band_K = 0; wa.kdiv_magic = 0; wa.escaped = nullptr; wa.escape_bit = 0;
        wa.pc_off = fa.pc_off; wa.plen = fa.plen;
        wa.pair_lo = ctx->pair_lo.as<int32_t>(); wa.pair_hi = ctx->pair_hi.as<int32_t>();
        wa.reads = d_reads; wa.offsets = d_offsets; wa.amplicon = ampb.as<uint8_t>();
        wa.La = La; wa.GK = GK; wa.P = P; wa.G = G; wa.K = K; wa.p0 = fa.p0; wa.p1 = fa.p1;
        wa.open = open_s; wa.ext = ext_s; wa.scale = scale;
        wa.recs = d_recs; wa.ref_out = d_ref; wa.mark_out = d_mark; wa.qry_out = d_qry; wa.slot = slot;
        wa.out_index = d_out_index; wa.rc_out = rc_out; wa.ops_out = d_ops; wa.ops_stride = ops_stride;
        CK(cudaStreamWaitEvent(s2, ctx->fill_done[cur], 0));
        span_begin(ctx, T_WALK, s2);
        CK(launch_walk(wa, s2));
        span_end(ctx);
        CK(cudaEventRecord(ctx->walk_done[cur], s2));
        used[cur] = true;
    }
    // everything queued later on the main stream (next pass, quantification, copies) sees the walks' results
    if (!lane) for (int i = 0; i < 2; ++i) if (used[i]) CK(cudaStreamWaitEvent(s, ctx->walk_done[i], 0));
    (void)plan_pc_off;
    return CRGPU_OK;
}


// ---------------------------------------------------------------------------------------------
// run_plan_dual: the amplicon pass and the HDR-amplicon pass of CORE:1791-1828 in one sweep.
// DP row y depends only on amplicon rows <= y, so all rows above the first base where the two
// amplicons differ are IDENTICAL in both matrices (values, flags, last-column summaries).  The
// amplicon pass runs the full tile (G lanes) and saves what lane t0-1 hands down; the HDR pass runs
// only the bottom Gh = G - t0 lanes (a compiled tile: Gh in {4,8,16,32}) with that saved row as its
// top boundary, and its traceback walks into the amplicon pass's flags above the split.  Results are
// bit-identical to two full passes (tests/test_gpu_hotpath.py, scripts/gpu_fuzz_hotpath.py).
// ---------------------------------------------------------------------------------------------
int run_plan_dual(crgpu_ctx *ctx, const char *amplicon, const char *hdr_amplicon, int La, const uint8_t *d_reads,
                  const int64_t *d_offsets, double gapopen, double gapextend, crgpu_aln_rec *d_recs, crgpu_aln_rec *d_recs_hdr,
                  uint8_t *d_ref, uint8_t *d_mark, uint8_t *d_qry, int64_t slot, int64_t *n_cells, int64_t *n_cells_computed,
                  uint32_t *d_ops, int64_t ops_stride, bool *done)
{
    *done = false;
    const PairPlan &pl = ctx->plan;
    if (!ctx->share_prefix || getenv("CRGPU_NO_SHARE") || pl.np == 0) return CRGPU_OK;
    if (La < CRGPU_MIN_LEN || La > CRGPU_MAX_AMPLICON) return CRGPU_OK;      // let run_plan report it
    int scale, open_s, ext_s;
    if (!scale_penalties(gapopen, gapextend, &scale, &open_s, &ext_s) || gapopen < 4.0 || gapextend < 0.0 || gapextend > gapopen)
        return CRGPU_OK;
    std::vector<int> acode(La), hcode(La);
    std::string amp_up(La, 'N'), hdr_up(La, 'N');
    int d = La;                                                            // first differing base
    for (int i = 0; i < La; ++i) {
        acode[i] = host_code(amplicon[i]); hcode[i] = host_code(hdr_amplicon[i]);
        if (acode[i] < 0 || hcode[i] < 0) return CRGPU_OK;
        amp_up[i] = host_upper(amplicon[i]); hdr_up[i] = host_upper(hdr_amplicon[i]);
        if (d == La && acode[i] != hcode[i]) d = i;
    }
    int G, K;
    if (!choose_tile(La, &G, &K)) return CRGPU_OK;
    const int GK = G * K, P = GK - La;
    // smallest compiled lane count Gh < G whose rows all lie at or below the first difference
    int Gh = 0;
    for (int g = 4; g < G; g *= 2)
        if (tile_available(g, K) && (G - g) * K - P <= d) { Gh = g; break; }
    if (Gh == 0) return CRGPU_OK;
    const int t0 = G - Gh, split = t0 * K, GKh = Gh * K;
    const int maxlen = pl.maxlen;
    // the score pass adds a drift of up to ext * (rows + columns) to every value (gotoh_score2.cu)
    if ((int64_t)scale * 5 * std::min(La, maxlen) + (int64_t)ext_s * (GK + maxlen + 2) + 64 >= MAX_ABS_SCORE ||
        (int64_t)2 * open_s + (int64_t)ext_s * (La + maxlen + 2) + 8 * scale + 64 >= MAX_ABS_SCORE)
        return CRGPU_OK;
    if (slot < (int64_t)La + maxlen && d_ref) return CRGPU_OK;
    if (n_cells) *n_cells += 2 * (int64_t)La * pl.sum_len;
    if (n_cells_computed) *n_cells_computed += ((int64_t)La + (int64_t)(La - (split - P))) * pl.sum_len;
    ctx->cells_kind[0] += ((int64_t)La + (int64_t)(La - (split - P))) * pl.sum_len;

    // ---- batches: a pair needs Lb * (GK + GKh) / 2 traceback words ----
    std::vector<int> batch_start(1, 0);
    int64_t max_cols = 0;
    int max_bp = 0;
    {
        const int64_t words_per_col = (GK + GKh) / 2;
        const int64_t budget_words = (int64_t)(ctx->tb_budget / 4);
        int64_t acc = 0, acc_cols = 0;
        int p_begin = 0;
        auto close = [&](int p_end) {
            max_cols = std::max(max_cols, acc_cols);
            max_bp = std::max(max_bp, p_end - p_begin);
            batch_start.push_back(p_end);
            p_begin = p_end; acc = 0; acc_cols = 0;
        };
        for (const HostSeg &sg : pl.segs) {
            const int64_t w = (int64_t)sg.len * words_per_col;
            int64_t left = ((int64_t)sg.cnt + 1) / 2;
            int64_t p = sg.pair_start;
            while (left > 0) {
                int64_t fit = (budget_words - acc) / w;
                if (fit <= 0) {
                    if (acc > 0) { close((int)p); continue; }
                    fit = 1;
                }
                const int64_t take = std::min(fit, left);
                acc += take * w; acc_cols += take * sg.len; p += take; left -= take;
                if (left > 0) close((int)p);
            }
        }
        if (acc > 0) close(pl.np);
    }

    // ---- profiles: full frame for the amplicon, bottom Gh lanes for the HDR amplicon ----
    auto build_prof = [&](const std::vector<int> &code, int g, int row0, std::vector<int32_t> &prof) {
        const int PS = prof_stride(g, K), SS = strip_stride(K);
        prof.assign((size_t)NPAIR * PS, 0);
        for (int cp = 0; cp < NPAIR; ++cp) {
            const int lo = cp % NCODE, hi = cp / NCODE;
            for (int r = std::max(P, row0); r < GK; ++r) {
                const int a = code[r - P];
                const int32_t slo = scale * host_ednafull(a, lo), shi = scale * host_ednafull(a, hi);
                const int rr = r - row0;
                prof[(size_t)cp * PS + (rr / K) * SS + (rr % K)] = shi * 65536 + slo;
            }
        }
    };
    std::vector<int32_t> prof_a, prof_h;
    build_prof(acode, G, 0, prof_a);
    build_prof(hcode, Gh, split, prof_h);
    cudaStream_t s = ctx->stream;
    const bool two = batch_start.size() > 2 && ctx->overlap && !getenv("CRGPU_NO_OVERLAP");
    CK(ctx->amp.reserve((size_t)La)); CK(ctx->amp_h.reserve((size_t)La));
    CK(ctx->prof.reserve(prof_a.size() * 4)); CK(ctx->prof_h.reserve(prof_h.size() * 4));
    DBuf *tbA[2] = {&ctx->tb, &ctx->tb2}, *tbH[2] = {&ctx->tbh, &ctx->tbh2}, *top[2] = {&ctx->top, &ctx->top2};
    DBuf *lrA[2] = {&ctx->lastrow, &ctx->lastrow2}, *lcA[2] = {&ctx->lastcol, &ctx->lastcol2};
    DBuf *lrH[2] = {&ctx->lastrow_h, &ctx->lastrow_h2}, *lcH[2] = {&ctx->lastcol_h, &ctx->lastcol_h2};
    for (int i = 0; i < (two ? 2 : 1); ++i) {
        CK(tbA[i]->reserve((size_t)max_cols * (GK / 2) * 4));
        CK(tbH[i]->reserve((size_t)max_cols * (GKh / 2) * 4));
        CK(top[i]->reserve((size_t)(max_cols + max_bp + 2) * 16));
        CK(lrA[i]->reserve((size_t)max_bp * 12)); CK(lcA[i]->reserve((size_t)max_bp * G * 12));
        CK(lrH[i]->reserve((size_t)max_bp * 12)); CK(lcH[i]->reserve((size_t)max_bp * Gh * 12));
    }
    CK(push_small(ctx, ctx->amp.p, amp_up.data(), (size_t)La, s));
    CK(push_small(ctx, ctx->amp_h.p, hdr_up.data(), (size_t)La, s));
    CK(push_small(ctx, ctx->prof.p, prof_a.data(), prof_a.size() * 4, s));
    CK(push_small(ctx, ctx->prof_h.p, prof_h.data(), prof_h.size() * 4, s));

    cudaStream_t s2 = two ? ctx->stream2 : ctx->stream;
    cudaStream_t sf[2] = {s, two ? ctx->stream3 : s};
    if (two) {
        CK(cudaEventRecord(ctx->ready, s));
        CK(cudaStreamWaitEvent(sf[1], ctx->ready, 0));
    }
    bool used[2] = {false, false};
    for (size_t b = 0; b + 1 < batch_start.size(); ++b) {
        const int cur = two ? (int)(b & 1) : 0;
        FillArgs fa;
        fa.prof = ctx->prof.as<int32_t>();
        fa.pc = ctx->pc.as<uint8_t>(); fa.pc_off = ctx->pc_off.as<int64_t>(); fa.plen = ctx->plen.as<int32_t>();
        fa.tb = tbA[cur]->as<uint32_t>(); fa.lastrow = lrA[cur]->as<uint32_t>(); fa.lastcol = lcA[cur]->as<uint32_t>();
        fa.p0 = batch_start[b]; fa.p1 = batch_start[b + 1];
        fa.open = open_s; fa.ext = ext_s; fa.La = La; fa.one = 1;
        fa.top_out = top[cur]->as<uint32_t>(); fa.top_out_lane = t0 - 1; fa.top_in = nullptr;
        fa.band_B = 0; fa.band_W = 0; fa.band_K = 0; fa.d_e = fa.d_eK = fa.d_eKb = fa.d_copen = 0; fa.band_row0 = 0; fa.band_tops = nullptr; fa.band_left = nullptr; fa.band_tb = nullptr;
        FillArgs fh = fa;                                               // HDR pass: bottom Gh lanes
        fh.prof = ctx->prof_h.as<int32_t>();
        fh.tb = tbH[cur]->as<uint32_t>(); fh.lastrow = lrH[cur]->as<uint32_t>(); fh.lastcol = lcH[cur]->as<uint32_t>();
        fh.La = GKh;                                                    // no padding rows inside the sub-tile
        fh.top_out = nullptr; fh.top_out_lane = -1; fh.top_in = top[cur]->as<uint32_t>();
        if (used[cur]) CK(cudaStreamWaitEvent(sf[cur], ctx->walk_done[cur], 0));
        span_begin(ctx, T_FILL, sf[cur]);
        CK(launch_fill(G, K, fa, ctx->num_sms, sf[cur]));
        CK(launch_fill(Gh, K, fh, ctx->num_sms, sf[cur]));
        span_end(ctx, 2);
        CK(cudaEventRecord(ctx->fill_done[cur], sf[cur]));

        WalkArgs wa;
        wa.tb = fa.tb; wa.lastrow = fa.lastrow; wa.lastcol = fa.lastcol;
        wa.tb_upper = nullptr; wa.lastcol_upper = nullptr; wa.G_upper = 0; wa.split_row = 0;
        wa.join_row = 0; wa.join_out = nullptr; wa.join_in = nullptr;
        wa.band_B = 0; wa.band_W = 0; wa.band_K = 0; wa.kdiv_magic = 0; wa.escaped = nullptr; wa.escape_bit = 0;
        wa.pc_off = fa.pc_off; wa.plen = fa.plen;
        wa.pair_lo = ctx->pair_lo.as<int32_t>(); wa.pair_hi = ctx->pair_hi.as<int32_t>();
        wa.reads = d_reads; wa.offsets = d_offsets; wa.amplicon = ctx->amp.as<uint8_t>();
        wa.La = La; wa.GK = GK; wa.P = P; wa.G = G; wa.K = K; wa.p0 = fa.p0; wa.p1 = fa.p1;
        wa.open = open_s; wa.ext = ext_s; wa.scale = scale;
        wa.recs = d_recs; wa.ref_out = d_ref; wa.mark_out = d_mark; wa.qry_out = d_qry; wa.slot = slot;
        wa.out_index = nullptr; wa.rc_out = 0; wa.ops_out = d_ops; wa.ops_stride = ops_stride;
        WalkArgs wh = wa;                                               // HDR alignment: identity only
        wh.tb = fh.tb; wh.lastrow = fh.lastrow; wh.lastcol = fh.lastcol;
        wh.tb_upper = fa.tb; wh.lastcol_upper = fa.lastcol; wh.G_upper = G; wh.split_row = split;
        wh.amplicon = ctx->amp_h.as<uint8_t>();
        wh.GK = GKh; wh.G = Gh;
        wh.recs = d_recs_hdr; wh.ref_out = wh.mark_out = wh.qry_out = nullptr; wh.ops_out = nullptr;
        CK(cudaStreamWaitEvent(s2, ctx->fill_done[cur], 0));
        span_begin(ctx, T_WALK, s2);
        CK(launch_walk(wa, s2));
        CK(launch_walk(wh, s2));
        span_end(ctx, 2);
        CK(cudaEventRecord(ctx->walk_done[cur], s2));
        used[cur] = true;
    }
    for (int i = 0; i < 2; ++i) if (used[i]) CK(cudaStreamWaitEvent(s, ctx->walk_done[i], 0));
    *done = true;
    return CRGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// run_plan_band: the forward alignment pass(es) with the banded two-pass fill (gotoh_fill.cu).
//   hdr_amplicon == nullptr : one amplicon.
//   hdr_amplicon != nullptr : amplicon + HDR amplicon sharing the DP prefix (as run_plan_dual); returns
//                             *done = false when the two cannot share rows (the caller then makes two
//                             single-amplicon calls).
// Reads whose traceback leaves the band get d_escaped[read] |= escape bit (1: amplicon pass, 2: HDR pass)
// and are re-aligned by the caller with the full fill.  Returns *done = false (nothing launched) when the
// band would not pay (reads not much longer than the band).
// ---------------------------------------------------------------------------------------------
int run_plan_band(crgpu_ctx *ctx, const char *amplicon, const char *hdr_amplicon, int La, const uint8_t *d_reads,
                  const int64_t *d_offsets, double gapopen, double gapextend, crgpu_aln_rec *d_recs, crgpu_aln_rec *d_recs_hdr,
                  uint8_t *d_ref, uint8_t *d_mark, uint8_t *d_qry, int64_t slot, int64_t *n_cells, int64_t *n_cells_computed,
                  uint32_t *d_ops, int64_t ops_stride, uint8_t *d_escaped, int escape_bit, uint8_t *d_fast, bool *done)
{
    *done = false;
    const PairPlan &pl = ctx->plan;
    int B = ctx->band_B;
    if (B <= 0 || getenv("CRGPU_NO_BAND") || pl.np == 0 || pl.nsub == 0) return CRGPU_OK;
    if (La < CRGPU_MIN_LEN || La > CRGPU_MAX_AMPLICON) return CRGPU_OK;      // let run_plan report it
    int scale, open_s, ext_s;
    if (!scale_penalties(gapopen, gapextend, &scale, &open_s, &ext_s) || gapopen < 4.0 || gapextend < 0.0 || gapextend > gapopen)
        return CRGPU_OK;
    const bool dual = hdr_amplicon != nullptr;
    std::vector<int> acode(La), hcode(La);
    std::string amp_up(La, 'N'), hdr_up(La, 'N');
    int d = La;                                                            // first differing base
    for (int i = 0; i < La; ++i) {
        acode[i] = host_code(amplicon[i]);
        if (acode[i] < 0) return CRGPU_OK;
        amp_up[i] = host_upper(amplicon[i]);
        if (dual) {
            hcode[i] = host_code(hdr_amplicon[i]);
            if (hcode[i] < 0) return CRGPU_OK;
            hdr_up[i] = host_upper(hdr_amplicon[i]);
            if (d == La && acode[i] != hcode[i]) d = i;
        }
    }
    int G, K;
    if (!choose_tile(La, &G, &K)) return CRGPU_OK;
    const int GK = G * K, P = GK - La;
    // k_gotoh_score2 takes two read columns per step and saves a sub-strip's left edge after the SECOND one: every
    // sub-strip's first band column xlo = u*Kb - P - B has to be even (Kb is), so an odd P + B widens the band by one
    if ((P + B) & 1) ++B;
    // the band pass works on sub-strips of Kb rows: half a lane's strip when that keeps 16-byte flag stores
    const int nsub = (K % 16 == 0 && !getenv("CRGPU_NO_SUBSTRIP")) ? 2 : 1;
    const int Kb = K / nsub;
    const int W = Kb + 2 * B + 1;
    // the band pays when it is well below the read length (score pass ~0.4x + band pass ~W/Lb of a full fill)
    if ((int64_t)W * 5 * pl.nsub > 3 * pl.sum_len) return CRGPU_OK;
    int Gh = 0;
    if (dual) {
        if (!ctx->share_prefix || getenv("CRGPU_NO_SHARE")) return CRGPU_OK;
        for (int g = 4; g < G; g *= 2)
            if (tile_available(g, K) && (G - g) * K - P <= d) { Gh = g; break; }
        if (Gh == 0) return CRGPU_OK;
    }
    const int t0 = G - Gh, split = dual ? t0 * K : 0, GKh = Gh * K;
    const int maxlen = pl.maxlen;
    // the score pass adds a drift of up to ext * (rows + columns) to every value (gotoh_score2.cu)
    if ((int64_t)scale * 5 * std::min(La, maxlen) + (int64_t)ext_s * (GK + maxlen + 2) + 64 >= MAX_ABS_SCORE ||
        (int64_t)2 * open_s + (int64_t)ext_s * (La + maxlen + 2) + 8 * scale + 64 >= MAX_ABS_SCORE)
        return CRGPU_OK;
    if (slot < (int64_t)La + maxlen && d_ref) return CRGPU_OK;

    // ---- batches: fixed scratch per pair and lane (tops, left edge, band flags) ----
    const int TOPW = band_topw(W), LEFTW = band_leftw(Kb);
    const size_t sub_bytes = (size_t)TOPW * 16 + (size_t)LEFTW * 4 + (size_t)W * Kb * 2;      // per sub-strip
    const size_t pair_bytes = sub_bytes * (size_t)(G + Gh) * nsub + (dual ? (size_t)(maxlen + 2) * 16 : 0);
    // ~8 batches per call (the walks of one batch overlap the fills of the next), each a whole number of WAVES of the
    // persistent score kernels: a warp of k_gotoh_score2 owns pairs w, w + wave, ..., so a batch of 9.2 waves costs 10
    int64_t bp = (int64_t)(ctx->tb_budget / pair_bytes);
    {
        const int64_t wave = getenv("CRGPU_NO_WAVE_ALIGN") ? 1 : std::max<int64_t>(1, score2_wave_pairs(dual ? Gh : G, K, nsub, ctx->num_sms));
        // two batches per call (CRGPU_BATCHES): enough for the walks of one to run beside the fills of the other, and every
        // extra batch pays the walk's latency floor -- ~0.4 ms for 40 k alignments as for 140 k -- again (2^20 reads + HDR,
        // ms per call: 8 batches 33.8, 4: 31.7, 2: 30.5, 1: 31.6; profiles/r02_notes.md)
        static const int nbatch = getenv("CRGPU_BATCHES") ? std::max(1, atoi(getenv("CRGPU_BATCHES"))) : 2;
        int64_t want = std::max<int64_t>(((int64_t)pl.np + nbatch - 1) / nbatch, 16384);
        if (want >= 4 * wave) {                   // (small calls, e.g. the chunks of a pipelined run, measured better unaligned)
            want = (want + wave - 1) / wave * wave;
            if (bp >= wave) bp = bp / wave * wave;
        }
        bp = std::max<int64_t>(1, std::min<int64_t>(bp, want));
    }
    std::vector<int> batch_start;
    for (int64_t p = 0; p < pl.np; p += bp) batch_start.push_back((int)p);
    batch_start.push_back(pl.np);
    const int max_bp = (int)std::min<int64_t>(bp, pl.np);
    int64_t max_cols = 0;
    for (size_t b = 0; b + 1 < batch_start.size(); ++b)
        max_cols = std::max(max_cols, plan_pc_off(pl, batch_start[b + 1]) - plan_pc_off(pl, batch_start[b]));
    if (n_cells) *n_cells += (dual ? 2 : 1) * (int64_t)La * pl.sum_len;
    int64_t band_cells_all = 0, band_cells_hdr = 0;
    {
        // score pass: every cell (HDR: the rows below the split); band pass: at most W columns per lane
        const int64_t band_cols = std::min<int64_t>((int64_t)W * pl.nsub, pl.sum_len);
        const int64_t score_cells = ((int64_t)La + (dual ? (int64_t)(La - (split - P)) : 0)) * pl.sum_len;
        const int64_t band_cells = ((int64_t)GK + (dual ? (int64_t)GKh : 0)) * band_cols;
        band_cells_all = band_cells;
        band_cells_hdr = dual ? (int64_t)GKh * band_cols : 0;        // (the shortcut only thins out the amplicon pass)
        if (n_cells_computed) *n_cells_computed += score_cells;
        ctx->cells_kind[1] += score_cells;
    }

    // ---- profiles ----
    // `add` = 2 ext for the score pass, which works in drift coordinates (gotoh_score2.cu); padding rows score 0 (+ add)
    auto build_prof = [&](const std::vector<int> &code, int g, int row0, int add, std::vector<int32_t> &prof) {
        const int PS = prof_stride(g, K), SS = strip_stride(K);
        prof.assign((size_t)NPAIR * PS, 0);
        for (int cp = 0; cp < NPAIR; ++cp) {
            const int lo = cp % NCODE, hi = cp / NCODE;
            for (int r = row0; r < GK; ++r) {
                int32_t slo = add, shi = add;
                if (r >= P) {
                    const int a = code[r - P];
                    slo += scale * host_ednafull(a, lo); shi += scale * host_ednafull(a, hi);
                }
                const int rr = r - row0;
                prof[(size_t)cp * PS + (rr / K) * SS + (rr % K)] = shi * 65536 + slo;
            }
        }
    };
    std::vector<int32_t> prof_a, prof_h, prof_as, prof_hs;
    build_prof(acode, G, 0, 0, prof_a);
    build_prof(acode, G, 0, 2 * ext_s, prof_as);
    if (dual) { build_prof(hcode, Gh, split, 0, prof_h); build_prof(hcode, Gh, split, 2 * ext_s, prof_hs); }
    cudaStream_t s = ctx->stream;
    const bool two = batch_start.size() > 2 && ctx->overlap && !getenv("CRGPU_NO_OVERLAP");
    CK(ctx->amp.reserve((size_t)La)); CK(ctx->prof.reserve(prof_a.size() * 4));
    if (dual) { CK(ctx->amp_h.reserve((size_t)La)); CK(ctx->prof_h.reserve(prof_h.size() * 4)); }
    CK(ctx->prof_s.reserve(prof_as.size() * 4));
    if (dual) CK(ctx->prof_hs.reserve(prof_hs.size() * 4));
    DBuf *tbA[2] = {&ctx->tb, &ctx->tb2}, *tbH[2] = {&ctx->tbh, &ctx->tbh2}, *top[2] = {&ctx->top, &ctx->top2};
    DBuf *lrA[2] = {&ctx->lastrow, &ctx->lastrow2}, *lcA[2] = {&ctx->lastcol, &ctx->lastcol2};
    DBuf *lrH[2] = {&ctx->lastrow_h, &ctx->lastrow_h2}, *lcH[2] = {&ctx->lastcol_h, &ctx->lastcol_h2};
    if (dual) {                                                                 // one set of join records per scratch set
        CK(ctx->join.reserve((size_t)max_bp * 2 * JOIN_STRIDE * 4));
        CK(ctx->joinb.reserve((size_t)max_bp * 2 * JOIN_STRIDE * 4));
    }
    // (with an HDR amplicon the shortcut relies on the walk join: the HDR walk of a shortcut read must end at a checkpoint)
    const bool diag = ctx->diag && !getenv("CRGPU_NO_DIAG") && !(dual && getenv("CRGPU_NO_JOIN"));
    const size_t nbatches = batch_start.size() - 1;
    const size_t selbytes = select_scratch_bytes((int64_t)max_bp * 2);
    if (diag) {
        if (!d_fast) return fail(ctx, CRGPU_E_ARG, "run_plan_band: the diagonal shortcut needs the per-read flag array");
        CK(ctx->need_cnt.reserve(nbatches * 16));
        CK(cudaMemsetAsync(ctx->need_cnt.p, 0, nbatches * 16, ctx->stream));
    }
    for (int i = 0; i < (two ? 2 : 1); ++i) {
        CK(tbA[i]->reserve((size_t)max_bp * G * W * K * 2));
        CK(ctx->btops[i].reserve((size_t)max_bp * G * nsub * TOPW * 16));
        CK(ctx->bleft[i].reserve((size_t)max_bp * G * nsub * LEFTW * 4));
        CK(lrA[i]->reserve((size_t)max_bp * 12)); CK(lcA[i]->reserve((size_t)max_bp * G * 12));
        CK(ctx->rowvals[i].reserve((size_t)(max_cols + max_bp + 2) * 4));
        if (dual) CK(ctx->rowvals_h[i].reserve((size_t)(max_cols + max_bp + 2) * 4));
        if (diag) {
            CK(ctx->need[i].reserve((size_t)max_bp)); CK(ctx->plist[i].reserve((size_t)max_bp * 4));
            CK(ctx->need_read[i].reserve((size_t)max_bp * 2)); CK(ctx->rlist[i].reserve((size_t)max_bp * 8));
            if (dual) { CK(ctx->plist2[i].reserve((size_t)max_bp * 4)); CK(ctx->rlist_side[i].reserve((size_t)max_bp * 8)); }
            CK(ctx->selscratch[i].reserve(selbytes));
        }
        if (dual) {
            CK(tbH[i]->reserve((size_t)max_bp * Gh * W * K * 2));
            CK(ctx->btops_h[i].reserve((size_t)max_bp * Gh * nsub * TOPW * 16));
            CK(ctx->bleft_h[i].reserve((size_t)max_bp * Gh * nsub * LEFTW * 4));
            CK(top[i]->reserve((size_t)(max_cols + max_bp + 2) * 16));
            CK(lrH[i]->reserve((size_t)max_bp * 12)); CK(lcH[i]->reserve((size_t)max_bp * Gh * 12));
        }
    }
    CK(push_small(ctx, ctx->amp.p, amp_up.data(), (size_t)La, s));
    CK(push_small(ctx, ctx->prof.p, prof_a.data(), prof_a.size() * 4, s));
    if (dual) {
        CK(push_small(ctx, ctx->amp_h.p, hdr_up.data(), (size_t)La, s));
        CK(push_small(ctx, ctx->prof_h.p, prof_h.data(), prof_h.size() * 4, s));
        CK(push_small(ctx, ctx->prof_hs.p, prof_hs.data(), prof_hs.size() * 4, s));
    }
    CK(push_small(ctx, ctx->prof_s.p, prof_as.data(), prof_as.size() * 4, s));

    cudaStream_t s2 = two ? ctx->stream2 : ctx->stream;
    cudaStream_t sf[2] = {s, two ? ctx->stream3 : s};
    if (two) {
        CK(cudaEventRecord(ctx->ready, s));
        CK(cudaStreamWaitEvent(sf[1], ctx->ready, 0));
    }
    const uint32_t magic = (uint32_t)((((uint64_t)1 << 32) + Kb - 1) / Kb);
    bool used[2] = {false, false};
    for (size_t b = 0; b + 1 < batch_start.size(); ++b) {
        const int cur = two ? (int)(b & 1) : 0;
        FillArgs fa;
        fa.prof = ctx->prof.as<int32_t>();
        fa.pc = ctx->pc.as<uint8_t>(); fa.pc_off = ctx->pc_off.as<int64_t>(); fa.plen = ctx->plen.as<int32_t>();
        fa.tb = nullptr; fa.lastrow = lrA[cur]->as<uint32_t>(); fa.lastcol = lcA[cur]->as<uint32_t>();
        fa.p0 = batch_start[b]; fa.p1 = batch_start[b + 1];
        fa.open = open_s; fa.ext = ext_s; fa.La = La; fa.one = 1;
        fa.top_out = dual ? top[cur]->as<uint32_t>() : nullptr; fa.top_out_lane = dual ? t0 - 1 : -1; fa.top_in = nullptr;
        fa.band_B = B; fa.band_W = W; fa.band_K = Kb; fa.band_row0 = -P;
        fa.d_e = (uint32_t)ext_s * 0x10001u; fa.d_eK = fa.d_e * (uint32_t)K; fa.d_eKb = fa.d_e * (uint32_t)Kb;
        fa.d_copen = ((uint32_t)(ext_s - open_s) & 0xffffu) * 0x10001u;
        fa.band_tops = ctx->btops[cur].as<uint32_t>(); fa.band_left = ctx->bleft[cur].as<uint32_t>();
        fa.band_tb = tbA[cur]->as<uint32_t>();
        fa.lastrow_vals = ctx->rowvals[cur].as<uint32_t>();
        FillArgs fh = fa;                                               // HDR pass: bottom Gh lanes
        if (dual) {
            fh.prof = ctx->prof_h.as<int32_t>();
            fh.lastrow = lrH[cur]->as<uint32_t>(); fh.lastcol = lcH[cur]->as<uint32_t>();
            fh.La = GKh;                                                // no padding rows inside the sub-tile
            fh.top_out = nullptr; fh.top_out_lane = -1; fh.top_in = top[cur]->as<uint32_t>();
            fh.top_in_row = split - 1;                                  // the amplicon pass saved its tile row split - 1
            fh.lastrow_vals = ctx->rowvals_h[cur].as<uint32_t>();
            fh.band_row0 = split - P;
            fh.band_tops = ctx->btops_h[cur].as<uint32_t>(); fh.band_left = ctx->bleft_h[cur].as<uint32_t>();
            fh.band_tb = tbH[cur]->as<uint32_t>();
        }
        if (used[cur]) CK(cudaStreamWaitEvent(sf[cur], ctx->walk_done[cur], 0));
        FillArgs fas = fa, fhs = fh;                                    // score pass: drifted profiles
        fas.prof = ctx->prof_s.as<int32_t>(); fhs.prof = ctx->prof_hs.as<int32_t>();
        span_begin(ctx, T_SCORE, sf[cur]);
        CK(launch_fill(G, K, fas, ctx->num_sms, sf[cur], 1));
        if (dual) CK(launch_fill(Gh, K, fhs, ctx->num_sms, sf[cur], 1));
        // the GPU has milliseconds of work queued: the moment to start the copy of the NEXT staged batch (a big DMA slows the
        // launches and the small transfers that share PCIe with it; it costs nothing while the host only waits)
        if (b == 0) { const int frc = flush_stages(ctx); if (frc) return frc; }
        span_end(ctx, dual ? 2 : 1);

        WalkArgs wa;
        wa.tb = fa.band_tb; wa.lastrow = fa.lastrow; wa.lastcol = fa.lastcol;
        wa.tb_upper = nullptr; wa.lastcol_upper = nullptr; wa.G_upper = 0; wa.split_row = 0;
        wa.band_B = B; wa.band_W = W; wa.band_K = Kb; wa.kdiv_magic = magic; wa.escaped = d_escaped; wa.escape_bit = escape_bit;
        wa.join_row = 0; wa.join_out = nullptr; wa.join_in = nullptr;
        if (dual && !getenv("CRGPU_NO_JOIN")) { wa.join_row = split; wa.join_out = (cur ? ctx->joinb : ctx->join).as<int32_t>(); }
        wa.pc_off = fa.pc_off; wa.plen = fa.plen;
        wa.pair_lo = ctx->pair_lo.as<int32_t>(); wa.pair_hi = ctx->pair_hi.as<int32_t>();
        wa.reads = d_reads; wa.offsets = d_offsets; wa.amplicon = ctx->amp.as<uint8_t>();
        wa.La = La; wa.GK = GK; wa.P = P; wa.G = G; wa.K = K; wa.p0 = fa.p0; wa.p1 = fa.p1;
        wa.open = open_s; wa.ext = ext_s; wa.scale = scale;
        wa.recs = d_recs; wa.ref_out = d_ref; wa.mark_out = d_mark; wa.qry_out = d_qry; wa.slot = slot;
        wa.out_index = nullptr; wa.rc_out = 0; wa.ops_out = d_ops; wa.ops_stride = ops_stride;
        WalkArgs wh = wa;                                               // HDR alignment: identity only
        if (dual) {
            wh.tb = fh.band_tb; wh.lastrow = fh.lastrow; wh.lastcol = fh.lastcol;
            wh.tb_upper = fa.band_tb; wh.lastcol_upper = fa.lastcol; wh.G_upper = G; wh.split_row = split;
            wh.amplicon = ctx->amp_h.as<uint8_t>();
            wh.GK = GKh; wh.G = Gh;
            wh.escape_bit = 2;
            wh.join_out = nullptr; wh.join_in = wa.join_out;      // (join_row: the same split row)
            wh.recs = d_recs_hdr; wh.ref_out = wh.mark_out = wh.qry_out = nullptr; wh.ops_out = nullptr;
        }
        FillArgs fa_up;                                                 // (amplicon band of the shortcut pairs: rows above the split)
        const int32_t *side_list = nullptr; const int *side_n = nullptr;
        if (wa.join_out) CK(cudaMemsetAsync(wa.join_out, 0, (size_t)(fa.p1 - fa.p0) * 2 * JOIN_STRIDE * 4, sf[cur]));
        if (diag) {
            // diagonal shortcut: amplicon alignments whose traceback is provably the diagonal through the start cell are
            // emitted here, right after the score pass; the amplicon band pass and walk below only visit the remaining
            // pairs.  (The HDR alignments of a run with an HDR amplicon keep the full path: an unedited read is not
            // diagonal against the HDR amplicon, and its walk ends at the first join checkpoint above the split anyway.)
            int *d_cnt = ctx->need_cnt.as<int>() + b;
            WalkArgs wd = wa;
            wd.fast = d_fast; wd.fast_bit = escape_bit; wd.need = ctx->need[cur].as<uint8_t>();
            CK(cudaMemsetAsync(wd.need, 2, (size_t)(fa.p1 - fa.p0), sf[cur]));      // 2: both alignments took the shortcut (so far)
            wd.need_read = ctx->need_read[cur].as<uint8_t>();
            CK(cudaMemsetAsync(wd.need_read, 0, (size_t)2 * (fa.p1 - fa.p0), sf[cur]));
            span_begin(ctx, T_WALK, sf[cur]);
            CK(launch_diag_emit(wd, sf[cur]));
            span_end(ctx, 1);
            CK(select_flagged(wd.need, fa.p1 - fa.p0, 1, ctx->plist[cur].as<int32_t>(), d_cnt, ctx->selscratch[cur].p, selbytes, sf[cur]));
            fa.pair_list = ctx->plist[cur].as<int32_t>();
            fa.pair_list_n = d_cnt;
            // ... and the walk only the alignments that did not take it
            int *d_cnt3 = ctx->need_cnt.as<int>() + 2 * nbatches + b;
            CK(select_flagged(wd.need_read, 2 * (fa.p1 - fa.p0), 1, ctx->rlist[cur].as<int32_t>(), d_cnt3, ctx->selscratch[cur].p, selbytes, sf[cur]));
            wa.read_list = ctx->rlist[cur].as<int32_t>(); wa.read_list_n = d_cnt3;
            wa.fast = d_fast; wa.fast_bit = escape_bit;
            if (dual) {
                // the HDR walks of the reads whose amplicon alignment was emitted above (need_read == 2) wait for nothing but
                // the band pass: they run beside the amplicon walks of the other reads, on a stream of their own
                int *d_cnt4 = ctx->need_cnt.as<int>() + 3 * nbatches + b;
                CK(select_flagged(wd.need_read, 2 * (fa.p1 - fa.p0), 2, ctx->rlist_side[cur].as<int32_t>(), d_cnt4, ctx->selscratch[cur].p, selbytes, sf[cur]));
                side_list = ctx->rlist_side[cur].as<int32_t>(); side_n = d_cnt4;
                // the HDR walk of such a pair still needs the amplicon pass's flags for the rows right above the split, until it
                // meets the amplicon alignment's diagonal at a join checkpoint: ~32 rows of sub-strips for those pairs
                int *d_cnt2 = ctx->need_cnt.as<int>() + nbatches + b;
                CK(select_flagged(wd.need, fa.p1 - fa.p0, 2, ctx->plist2[cur].as<int32_t>(), d_cnt2, ctx->selscratch[cur].p, selbytes, sf[cur]));
                const int nup = std::min(split / Kb, (32 + Kb - 1) / Kb);
                fa_up = fa;
                fa_up.pair_list = ctx->plist2[cur].as<int32_t>(); fa_up.pair_list_n = d_cnt2;
                fa_up.sub_lo = split / Kb - nup; fa_up.sub_n = nup;
                wh.upper_need = wd.need; wh.upper_lo = fa_up.sub_lo * Kb;
            }
        }
        span_begin(ctx, T_BAND, sf[cur]);
        CK(launch_fill(G, K, fa, ctx->num_sms, sf[cur], 2));
        if (dual) CK(launch_fill(Gh, K, fh, ctx->num_sms, sf[cur], 2));
        if (fa_up.sub_n > 0) CK(launch_fill(G, K, fa_up, ctx->num_sms, sf[cur], 2));
        span_end(ctx, (dual ? 2 : 1) + (fa_up.sub_n > 0 ? 1 : 0));
        CK(cudaEventRecord(ctx->fill_done[cur], sf[cur]));

        CK(cudaStreamWaitEvent(s2, ctx->fill_done[cur], 0));
        if (side_list) {
            cudaStream_t s4 = two ? ctx->stream4 : s2;
            WalkArgs ws = wh;
            ws.read_list = side_list; ws.read_list_n = side_n;
            if (two) CK(cudaStreamWaitEvent(s4, ctx->fill_done[cur], 0));
            span_begin(ctx, T_WALK, s4);
            CK(launch_walk(ws, s4));
            span_end(ctx, 1);
            if (two) CK(cudaEventRecord(ctx->walk_side[cur], s4));
            wh.read_list = wa.read_list; wh.read_list_n = wa.read_list_n;      // ... and behind the amplicon walks, the HDR walks of their reads
        }
        span_begin(ctx, T_WALK, s2);
        CK(launch_walk(wa, s2));
        if (dual) CK(launch_walk(wh, s2));
        span_end(ctx, dual ? 2 : 1);
        if (side_list && two) CK(cudaStreamWaitEvent(s2, ctx->walk_side[cur], 0));
        CK(cudaEventRecord(ctx->walk_done[cur], s2));
        used[cur] = true;
    }
    for (int i = 0; i < 2; ++i) if (used[i]) CK(cudaStreamWaitEvent(s, ctx->walk_done[i], 0));
    // cells the band pass evaluated: every pair, or the pairs the diagonal shortcut left over
    int64_t band_cells = band_cells_all;
    ctx->n_diag_pairs[0] += pl.np;
    if (diag) {
        std::vector<int> hc(nbatches, 0);
        CK(fetch_small(ctx, hc.data(), ctx->need_cnt.p, nbatches * 4, s));
        CK(fetch_wait(ctx, s));
        int64_t left = 0;
        for (int c : hc) left += c;
        ctx->n_diag_pairs[1] += left;
        band_cells = band_cells_hdr + (int64_t)((double)(band_cells_all - band_cells_hdr) * (double)left / (double)std::max(pl.np, 1));
    } else ctx->n_diag_pairs[1] += pl.np;
    if (n_cells_computed) *n_cells_computed += band_cells;
    ctx->cells_kind[2] += band_cells;
    *done = true;
    return CRGPU_OK;
}

}  // namespace crgpu

extern "C" {

static int align_impl(crgpu_ctx *ctx, int mem, const char *amplicon, int amplicon_len,
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

    const uint8_t *d_reads = reads; const int64_t *d_off = offsets;
    crgpu_aln_rec *d_recs = recs; uint8_t *d_ref = ref_out, *d_mark = mark_out, *d_qry = qry_out;
    if (mem == CRGPU_MEM_HOST) {
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
    CK(ctx->badbase.reserve((size_t)n));
    CK(cudaMemsetAsync(ctx->badbase.p, 0, (size_t)n, s));
    ctx->d_bad = ctx->badbase.as<uint8_t>();
    int rc = build_plan(ctx, d_reads, d_off, nullptr, n);
    if (rc == CRGPU_OK)
        rc = run_plan(ctx, amplicon, amplicon_len, d_reads, d_off, nullptr, 0, gapopen, gapextend, d_recs, d_ref, d_mark, d_qry,
                      slot, nullptr);
    ctx->d_bad = nullptr;
    if (rc != CRGPU_OK) { cudaStreamSynchronize(s); return rc; }
    CK(launch_clear_bad_recs(d_recs, ctx->badbase.as<uint8_t>(), n, slot, s));
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

// BAM's 4-bit base codes, two per byte, high nibble first -> one base per byte.  16 bases per thread.
__global__ void k_unpack_bam4(const uint8_t *__restrict__ packed, int64_t nbases, uint8_t *__restrict__ out)
{
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (i >= nbases) return;
    const char *lut = "=ACMGRSVTWYHKDBN";
    const int64_t left = nbases - i;
    if (left >= 16) {
        const uint2 v = *reinterpret_cast<const uint2 *>(packed + i / 2);
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t two = ((q < 2 ? v.x : v.y) >> (16 * (q & 1))) & 0xffffu;     // bytes 2q, 2q+1 of the eight
            w[q] = (uint32_t)(uint8_t)lut[(two >> 4) & 15] | ((uint32_t)(uint8_t)lut[two & 15] << 8) |
                   ((uint32_t)(uint8_t)lut[(two >> 12) & 15] << 16) | ((uint32_t)(uint8_t)lut[(two >> 8) & 15] << 24);
        }
        *reinterpret_cast<uint4 *>(out + i) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
        for (int64_t j = i; j < nbases; ++j) {
            const uint8_t bb = packed[j >> 1];
            out[j] = (uint8_t)lut[(j & 1) ? (bb & 15) : (bb >> 4)];
        }
    }
}

// the copy of a staged batch (crgpu_stage_reads recorded it): H2D on the copy stream + unpack
static int issue_stage(crgpu_ctx *ctx, int slot)
{
    crgpu_ctx::StagePending &sp = ctx->stage_pend[slot];
    if (!sp.on) return CRGPU_OK;
    sp.on = false;
    cudaStream_t cs = ctx->stream_copy;
    CK(cudaMemcpyAsync(ctx->stage_off[slot].p, sp.offsets, (size_t)(sp.n + 1) * 8, cudaMemcpyHostToDevice, cs));
    if (sp.format == CRGPU_READS_BYTES) {
        CK(cudaMemcpyAsync(ctx->stage_reads[slot].p, sp.reads, (size_t)sp.total, cudaMemcpyHostToDevice, cs));
    } else {
        const size_t pb = (size_t)((sp.total + 1) / 2);
        CK(cudaMemcpyAsync(ctx->stage_pack[slot].p, sp.reads, pb, cudaMemcpyHostToDevice, cs));
        const int64_t nthreads = (sp.total + 15) / 16;
        k_unpack_bam4<<<(unsigned)((nthreads + 255) / 256), 256, 0, cs>>>(ctx->stage_pack[slot].as<uint8_t>(), sp.total,
                                                                         ctx->stage_reads[slot].as<uint8_t>());
        CK(cudaGetLastError());
        ctx->launches[T_OTHER] += 1;                              // (k_unpack_bam4)
    }
    CK(cudaEventRecord(ctx->staged_ev[slot], cs));
    return CRGPU_OK;
}

extern "C++" {
namespace crgpu {
int flush_stages(crgpu_ctx *ctx)
{
    for (int slot = 0; slot < 2; ++slot) {
        const int rc = issue_stage(ctx, slot);
        if (rc) return rc;
    }
    // ... and the deferred per-read outputs of the previous staged call
    for (int slot = 0; slot < 2; ++slot) {
        crgpu_ctx::OutPending &op = ctx->out_pend[slot];
        if (!op.on) continue;
        op.on = false;
        CK(cudaStreamWaitEvent(ctx->stream_copy, ctx->out_ready[slot], 0));
        for (int i = 0; i < op.n; ++i) CK(cudaMemcpyAsync(op.h[i], op.d[i], op.bytes[i], cudaMemcpyDeviceToHost, ctx->stream_copy));
        CK(cudaEventRecord(ctx->out_ev[slot], ctx->stream_copy));
    }
    return CRGPU_OK;
}
}  // namespace crgpu
}  // extern "C++"

// crgpu_stage_reads: checks, buffers, and a note of what to copy.  The copy itself starts when it disturbs least: inside the
// next crgpu_align_quantify* call, right after its first score-pass launches (flush_stages) -- or, when the batch is run /
// the context synchronised before that, right there.  CRGPU_EAGER_STAGE=1: start it at once.
static int stage_reads_impl(crgpu_ctx *ctx, int slot, int format, const uint8_t *reads, const int64_t *offsets, int64_t n)
{
    if (slot < 0 || slot > 1) return fail(ctx, CRGPU_E_ARG, "crgpu_stage_reads: slot must be 0 or 1");
    if (format != CRGPU_READS_BYTES && format != CRGPU_READS_BAM4) return fail(ctx, CRGPU_E_ARG, "crgpu_stage_reads: unknown format");
    if (n < 0 || (n > 0 && (!reads || !offsets)) || n >= (int64_t)1 << 31) return fail(ctx, CRGPU_E_ARG, "crgpu_stage_reads: bad argument");
    ctx->stage_n[slot] = -1;
    ctx->stage_pend[slot].on = false;
    if (n == 0) { ctx->stage_n[slot] = 0; return CRGPU_OK; }
    const int64_t total = offsets[n] - offsets[0];
    if (offsets[0] != 0 || total < 0) return fail(ctx, CRGPU_E_ARG, "crgpu_stage_reads: offsets must start at 0");
    // (16 bytes of slack: the unpack kernel stores whole 16-byte groups)
    CK(ctx->stage_reads[slot].reserve((size_t)std::max<int64_t>(total, 1) + 32));
    CK(ctx->stage_off[slot].reserve((size_t)(n + 1) * 8));
    if (format == CRGPU_READS_BAM4) CK(ctx->stage_pack[slot].reserve((size_t)((total + 1) / 2) + 16));
    ctx->stage_pend[slot] = {true, format, reads, offsets, n, total};
    ctx->stage_n[slot] = n;
    static const bool eager = getenv("CRGPU_EAGER_STAGE") != nullptr;
    if (eager) return issue_stage(ctx, slot);
    return CRGPU_OK;
}

static int int_peak_impl(crgpu_ctx *ctx, int which, double *lane_ops_per_s)
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

// ---- the exported entry points: device guard + "no work of a failed call is left running" (ApiGuard, crgpu_internal.h)
int crgpu_stage_reads(crgpu_ctx *ctx, int slot, int format, const uint8_t *reads, const int64_t *offsets, int64_t n)
{
    if (!ctx) return CRGPU_E_ARG;
    ApiGuard guard(ctx);
    const int rc = stage_reads_impl(ctx, slot, format, reads, offsets, n);
    if (rc != CRGPU_OK) cudaStreamSynchronize(ctx->stream_copy);
    return guard.done(rc);
}

int crgpu_sync(crgpu_ctx *ctx)
{
    if (!ctx) return CRGPU_E_ARG;
    ApiGuard guard(ctx);
    return guard.done(sync_impl(ctx));
}

int crgpu_qualfilter(crgpu_ctx *ctx, int mem, const uint8_t *qual, const int64_t *offsets, int64_t n,
                     int min_mean_q, int min_single_q, uint8_t *keep)
{
    if (!ctx) return CRGPU_E_ARG;
    ApiGuard guard(ctx);
    return guard.done(qualfilter_impl(ctx, mem, qual, offsets, n, min_mean_q, min_single_q, keep));
}

int crgpu_align(crgpu_ctx *ctx, int mem, const char *amplicon, int amplicon_len,
                const uint8_t *reads, const int64_t *offsets, int64_t n, double gapopen, double gapextend,
                crgpu_aln_rec *recs, uint8_t *ref_out, uint8_t *mark_out, uint8_t *qry_out, int64_t slot)
{
    if (!ctx) return CRGPU_E_ARG;
    ApiGuard guard(ctx);
    return guard.done(align_impl(ctx, mem, amplicon, amplicon_len, reads, offsets, n, gapopen, gapextend, recs, ref_out, mark_out, qry_out, slot));
}

int crgpu_int_peak(crgpu_ctx *ctx, int which, double *lane_ops_per_s)
{
    if (!ctx) return CRGPU_E_ARG;
    ApiGuard guard(ctx);
    return guard.done(int_peak_impl(ctx, which, lane_ops_per_s));
}

}  // extern "C"
