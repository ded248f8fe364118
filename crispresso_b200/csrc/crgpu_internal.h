// crgpu_internal.h -- context layout and helpers shared by the host-side translation units.
#pragma once
#include "../../include/crgpu.h"
#include "crgpu_common.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

// T_SCORE / T_BAND: the two passes of the banded fill; crgpu_last_timing folds them into the fill slot and
// crgpu_last_fill_breakdown reports them (with the DP cells each kind evaluated) separately
enum { T_ENCODE = 0, T_FILL, T_WALK, T_QUANT, T_QUAL, T_OTHER, T_SCORE, T_BAND, T_N };

struct DBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct TimedSpan { int family; cudaEvent_t a, b; };

// One length bucket of the pairing plan (same layout as crgpu::LenSeg in traceback_walk.cu).
struct HostSeg { int len; int cnt; int64_t read_start; int64_t pair_start; int64_t pc_start; };

// Pairing plan of the current read set (device arrays live in the context's scratch buffers).
struct PairPlan {
    int64_t nsub = 0;         // reads covered
    int np = 0;               // pairs
    int maxlen = 0, minlen = 0;
    int64_t sum_len = 0;      // sum of read lengths (DP cells = La * sum_len)
    int64_t total_pc = 0;     // sum of pair lengths
    std::vector<HostSeg> segs;
};

struct crgpu_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;      // traceback walks run here, overlapped with the next batch's fill
    cudaStream_t stream4 = nullptr;      // HDR walks of the reads whose amplicon alignment took the diagonal shortcut (beside the amplicon walks)
    cudaEvent_t walk_side[2] = {nullptr, nullptr};
    DBuf rlist_side[2];                  // those reads, per batch
    cudaStream_t stream3 = nullptr;      // odd fill batches: their CTAs back-fill the SMs the previous fill's tail vacates
    cudaEvent_t ready = nullptr;
    size_t l2_gran_prev = 0;             // CRGPU_L2_HINT: the device's L2 fetch granularity before crgpu_create changed it
    bool l2_gran_set = false;
    cudaStream_t span_stream = nullptr;
    cudaEvent_t fill_done[2] = {nullptr, nullptr}, walk_done[2] = {nullptr, nullptr};
    size_t tb_budget = (size_t)8 << 30;
    bool overlap = true;
    std::string err;
    // device scratch
    DBuf reads, offsets, amp, prof, pc, pc_off, plen, pair_lo, pair_hi, order, plan_hist, plan_tab, tb, lastrow, lastcol, tb2, lastrow2, lastcol2, errflag;
    PairPlan plan;
    DBuf recs, sref, smark, sqry, ops, ops_rc, alleles;
    DBuf prof_h, amp_h, tbh, tbh2, top, top2, lastrow_h, lastrow_h2, lastcol_h, lastcol_h2;   // HDR pass of run_plan_dual
    bool share_prefix = true;
    DBuf join, joinb;                                                      // amplicon walk -> HDR walk join records (WalkArgs.join_out)
    DBuf prof_s, prof_hs;                                           // drifted profiles of the score pass
    DBuf btops[2], bleft[2], btops_h[2], bleft_h[2], escaped;      // banded two-pass fill (run_plan_band)
    DBuf rowvals[2], rowvals_h[2];
    // reads with a base outside ACGTN(U) are not aligned but reported per read: one byte per read of the CALL in flight
    // (crgpu_align, crgpu_align_quantify), null otherwise (build_plan then fails the call)
    DBuf badbase;
    uint8_t *d_bad = nullptr;
    // staged inputs (crgpu_stage_reads): two slots filled on the copy stream while the main stream computes
    cudaStream_t stream_copy = nullptr;
    cudaEvent_t staged_ev[2] = {nullptr, nullptr};
    DBuf stage_reads[2], stage_off[2], stage_pack[2];
    int64_t stage_n[2] = {-1, -1};
    // deferred outputs of the staged calls: per-slot device buffers (kept, aln, recs, tenths_rep) + "they have left" events
    bool deferred_out = false;
    DBuf stage_out[2][4];
    cudaEvent_t out_ev[2] = {nullptr, nullptr};                                 // score pass: last-row values per column (k_lastrow_scan)
    int n_escaped[2] = {0, 0};                                     // reads re-aligned after the last banded call (amplicon, HDR)
    int band_holdoff = 0;                                          // calls left that skip the band (set when > 25 % of a call's reads escaped)
    int band_B = 16;                                               // band half-width in read columns; 0 = single-pass fill
    // diagonal shortcut of the banded fill (run_plan_band): alignments whose traceback is provably the diagonal through
    // the start cell are emitted right after the score pass; only the other pairs go through the band pass and the walk
    bool diag = true;
    // exact-read shortcut (hotpath.cu): reads identical to the amplicon skip the DP
    bool exact_shortcut = true;
    // staged batches whose copy has not been started yet (crgpu_stage_reads; flush_stages)
    struct StagePending { bool on; int format; const uint8_t *reads; const int64_t *offsets; int64_t n, total; };
    StagePending stage_pend[2] = {{false, 0, nullptr, nullptr, 0, 0}, {false, 0, nullptr, nullptr, 0, 0}};
    // deferred per-read outputs of a staged call whose copies have not been started yet (flush_stages)
    struct OutPending { bool on; int n; void *h[4]; const void *d[4]; size_t bytes[4]; };
    OutPending out_pend[2] = {};
    cudaEvent_t out_ready[2] = {nullptr, nullptr};
    // mailbox: pinned host memory mapped into the device's address space.  The counts, histograms and accumulator blocks the
    // host reads between launches (and the small tables it writes) travel through it with a copy KERNEL: a cudaMemcpyAsync of
    // 4 bytes would queue on the copy engine behind the staged reads / deferred outputs of the neighbouring batches
    uint8_t *mbox_h = nullptr, *mbox_d = nullptr;
    size_t mbox_bytes = 0, mbox_used = 0;
    struct Fetch { void *dst; size_t off, bytes; };
    std::vector<Fetch> mbox_pending;
    // CRGPU_TRACE=1: host wall-clock marks of the phases of crgpu_align_quantify, printed to stderr at the end of the call
    bool trace_on = false;
    std::vector<std::pair<const char *, double>> trace;
    int64_t n_exact = 0;                                           // last call: reads that skipped it
    DBuf exact_go, exact_sel;
    DBuf fastflags, need[2], plist[2], plist2[2], need_read[2], rlist[2], selscratch[2], need_cnt;
    int64_t n_diag_pairs[2] = {0, 0};                              // last call: pairs of the batches / pairs that needed the band pass
    DBuf q_in[8], q_out[4];
    DBuf aux[8];
    // timing
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    std::vector<TimedSpan> spans;
    float ms[T_N] = {0};
    int64_t launches[T_N] = {0};
    int64_t cells_kind[3] = {0, 0, 0};                             // DP cells evaluated: single-pass fill, score pass, band pass
};

inline int crgpu_fail(crgpu_ctx *c, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return crgpu_fail(ctx, e_ == cudaErrorMemoryAllocation ? CRGPU_E_NOMEM : CRGPU_E_CUDA, "%s: %s (%s:%d)", #call, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                     \
    } while (0)

inline cudaEvent_t next_event(crgpu_ctx *c)
{
    if (c->ev_used == c->ev_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        c->ev_pool.push_back(e);
    }
    return c->ev_pool[c->ev_used++];
}
inline void span_begin(crgpu_ctx *c, int family, cudaStream_t st = nullptr)
{
    if (!st) st = c->stream;
    c->span_stream = st;
    TimedSpan s{family, next_event(c), next_event(c)};
    cudaEventRecord(s.a, st);
    c->spans.push_back(s);
}
inline void span_end(crgpu_ctx *c, int kernels = 1)   // kernels launched inside the span
{
    cudaEventRecord(c->spans.back().b, c->span_stream);
    c->launches[c->spans.back().family] += kernels;
}
inline void timing_reset(crgpu_ctx *c)
{
    c->ev_used = 0;
    c->spans.clear();
    for (int i = 0; i < T_N; ++i) { c->ms[i] = 0; c->launches[i] = 0; }
    for (int i = 0; i < 3; ++i) c->cells_kind[i] = 0;
}
inline void timing_collect(crgpu_ctx *c)
{
    for (auto &s : c->spans) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) c->ms[s.family] += ms;
    }
    c->spans.clear();
    c->ev_used = 0;
}


#define fail crgpu_fail

// Every exported entry point runs inside one of these: the context's device is made current for the call and the caller's
// is restored afterwards; after a FAILED call everything it queued on the context's three streams is waited for, so that
// nothing is still writing caller buffers (CRGPU_MEM_DEVICE) once the error has been returned.
inline void trace_mark(crgpu_ctx *ctx, const char *what)
{
    if (!ctx->trace_on) return;
    const double t = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
    ctx->trace.emplace_back(what, t);
}

inline void trace_dump(crgpu_ctx *ctx)
{
    if (!ctx->trace_on || ctx->trace.empty()) return;
    std::string line = "[crgpu trace]";
    char buf[96];
    for (size_t i = 1; i < ctx->trace.size(); ++i) {
        snprintf(buf, sizeof buf, " %s %.3f", ctx->trace[i].first, ctx->trace[i].second - ctx->trace[i - 1].second);
        line += buf;
    }
    snprintf(buf, sizeof buf, " | total %.3f ms\n", ctx->trace.back().second - ctx->trace.front().second);
    line += buf;
    fputs(line.c_str(), stderr);
    ctx->trace.clear();
}

struct ApiGuard {
    crgpu_ctx *c;
    int prev = -1;
    explicit ApiGuard(crgpu_ctx *ctx) : c(ctx)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != c->device) cudaSetDevice(c->device);
    }
    int done(int rc)
    {
        if (rc != CRGPU_OK) {
            cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->stream2); cudaStreamSynchronize(c->stream3); cudaStreamSynchronize(c->stream4);
            c->mbox_pending.clear(); c->mbox_used = 0;       // (their destinations were locals of the failed call)
            cudaGetLastError();                       // (a sticky launch error has been reported through rc already)
        }
        return rc;
    }
    ~ApiGuard() { if (prev >= 0 && prev != c->device) cudaSetDevice(prev); }
};

namespace crgpu {
// start the copies of the staged batches that are still waiting for a good moment (crgpu_stage_reads)
int flush_stages(crgpu_ctx *ctx);
// device -> host through the mailbox, asynchronous on s: h_dst is valid after fetch_wait(ctx, s)
cudaError_t fetch_small(crgpu_ctx *ctx, void *h_dst, const void *d_src, size_t bytes, cudaStream_t s);
// cudaStreamSynchronize(s) + delivery of the pending fetches; the mailbox is empty afterwards
cudaError_t fetch_wait(crgpu_ctx *ctx, cudaStream_t s);
// host -> device through the mailbox, asynchronous on s; h_src may be reused as soon as the call returns
cudaError_t push_small(crgpu_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, cudaStream_t s);

bool choose_tile(int La, int *G, int *K);
bool tile_available(int G, int K);
cudaError_t launch_fill(int G, int K, const FillArgs &a, int num_sms, cudaStream_t stream, int kind = 0);
// pairs one resident wave of k_gotoh_score2<G,K,nsub> covers (0: no such kernel)
int64_t score2_wave_pairs(int G, int K, int nsub, int num_sms);
cudaError_t launch_encode(const uint8_t *reads, const int64_t *offsets, const int32_t *pair_lo, const int32_t *pair_hi,
                          const int64_t *pc_off, int npairs, uint8_t *pc, int *err, uint8_t *bad, int num_sms, cudaStream_t s);
cudaError_t launch_clear_bad_recs(crgpu_aln_rec *recs, const uint8_t *bad, int64_t n, int64_t slot, cudaStream_t s);
cudaError_t launch_walk(const WalkArgs &a, cudaStream_t s);
cudaError_t launch_diag_emit(const WalkArgs &a, cudaStream_t s);
cudaError_t launch_qualfilter(const uint8_t *qual, const int64_t *offsets, int64_t n, int q, int sq, uint8_t *keep,
                              int num_sms, cudaStream_t s);
cudaError_t launch_int_peak(int which, int num_sms, int iters, unsigned *sink, cudaStream_t s, double *lane_ops);
cudaError_t launch_len_hist(const int64_t *offsets, const int32_t *subset, int64_t n, int min_len, int max_len, int *hist,
                            int *err, cudaStream_t s);
cudaError_t launch_scatter_order(const int64_t *offsets, const int32_t *subset, int64_t n, const int64_t *read_start,
                                 int *cursor, int32_t *order, cudaStream_t s);
cudaError_t launch_build_pairs(const void *segs, int nseg, int np, const int32_t *order, int32_t *pair_lo, int32_t *pair_hi,
                               int32_t *plen, int64_t *pc_off, int64_t total_pc, cudaStream_t s);
size_t allele_scratch_bytes(int64_t m);
size_t select_scratch_bytes(int64_t n);
cudaError_t select_flagged(const uint8_t *flags, int64_t n, int bit, int32_t *out_idx, int *d_count, void *tmp, size_t tmp_bytes,
                           cudaStream_t s);
cudaError_t allele_groups(const uint8_t *reads, const int64_t *offsets, int64_t n, const uint8_t *kept, const int32_t *rc_read,
                          int64_t nrc, const uint32_t *ops_fw, const uint32_t *ops_rc, int64_t ops_stride,
                          const crgpu_aln_rec *aln_fw, const crgpu_aln_rec *aln_rc, const crgpu_read_rec *rec_fw,
                          const crgpu_read_rec *rec_rc, int reads_upper, void *scratch, size_t scratch_bytes, cudaStream_t s,
                          int32_t **d_rep_sorted, int32_t **d_count_sorted, int **d_nruns, int **d_err, uint64_t **d_key_pairs);
// alignment core (crgpu_api.cu): device pointers only
int build_plan(crgpu_ctx *ctx, const uint8_t *d_reads, const int64_t *d_offsets, const int32_t *d_subset, int64_t nsub);
int run_plan(crgpu_ctx *ctx, const char *amplicon, int La, const uint8_t *d_reads, const int64_t *d_offsets,
             const int32_t *d_out_index, int rc_out, double gapopen, double gapextend, crgpu_aln_rec *d_recs,
             uint8_t *d_ref, uint8_t *d_mark, uint8_t *d_qry, int64_t slot, int64_t *n_cells,
             uint32_t *d_ops = nullptr, int64_t ops_stride = 0, int lane = 0);
// amplicon + HDR amplicon in one sweep over the batches; the HDR pass reuses the DP rows it shares with
// the amplicon pass.  Returns CRGPU_OK and *done = false when the two amplicons cannot share a prefix
// (the caller then runs two plain passes).
int run_plan_dual(crgpu_ctx *ctx, const char *amplicon, const char *hdr_amplicon, int La, const uint8_t *d_reads,
                  const int64_t *d_offsets, double gapopen, double gapextend, crgpu_aln_rec *d_recs, crgpu_aln_rec *d_recs_hdr,
                  uint8_t *d_ref, uint8_t *d_mark, uint8_t *d_qry, int64_t slot, int64_t *n_cells, int64_t *n_cells_computed,
                  uint32_t *d_ops, int64_t ops_stride, bool *done);
int run_plan_band(crgpu_ctx *ctx, const char *amplicon, const char *hdr_amplicon, int La, const uint8_t *d_reads,
                  const int64_t *d_offsets, double gapopen, double gapextend, crgpu_aln_rec *d_recs, crgpu_aln_rec *d_recs_hdr,
                  uint8_t *d_ref, uint8_t *d_mark, uint8_t *d_qry, int64_t slot, int64_t *n_cells, int64_t *n_cells_computed,
                  uint32_t *d_ops, int64_t ops_stride, uint8_t *d_escaped, int escape_bit, uint8_t *d_fast, bool *done);
}  // namespace crgpu
